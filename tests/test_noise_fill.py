"""Device-side disturbance draws: the Philox4x32-10 restatement against the published known-answer
vectors (CPU), the CUDA generator against the restatement (GPU)."""
import numpy as np
import pytest

import mpc_arpo_project_b200 as M
from oracle.gen_golden import make_params
from oracle.philox_ref import philox4x32_10, noise_fill


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(v) for v in philox4x32_10(ctr, key)] == want


def test_noise_model_moments_and_stream_separation():
    n = noise_fill(100000, 4, 0.75, 0.5, seed=7)
    assert abs(n.mean()) < 5e-3
    np.testing.assert_allclose(n[:, 0].std(), 0.75, rtol=1e-2)
    np.testing.assert_allclose(n[:, 1].std(), 0.5, rtol=1e-2)
    assert abs(np.corrcoef(n[:, 0].ravel(), n[:, 1].ravel())[0, 1]) < 1e-2
    # lane_offset continues the same stream: ranks of a multi-GPU run draw disjoint lanes
    a = noise_fill(64, 3, 1.0, 1.0, seed=9, lane_offset=0)
    b = noise_fill(32, 3, 1.0, 1.0, seed=9, lane_offset=32)
    np.testing.assert_array_equal(a[:, :, 32:], b)
    assert not np.array_equal(noise_fill(8, 1, 1, 1, seed=1), noise_fill(8, 1, 1, 1, seed=2))


@pytest.mark.gpu
def test_device_generator_matches_the_restatement():
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.75, noise_length=50))
    eng = M.Engine(M.build_problem(sc, mp, fp, None))
    B, R, seed, off = 1000, 7, 0x1234_5678_9abc_def0, 2 ** 33 + 5
    noise, words = eng.noise_fill(B, R, seed, off, raw=True)
    lane = np.arange(B, dtype=np.uint64) + np.uint64(off)
    ctr = np.zeros((R, B, 4), dtype=np.uint32)
    ctr[..., 0] = (lane & np.uint64(0xFFFFFFFF)).astype(np.uint32)[None]
    ctr[..., 1] = (lane >> np.uint64(32)).astype(np.uint32)[None]
    ctr[..., 2] = np.arange(R, dtype=np.uint32)[:, None]
    want = philox4x32_10(ctr, [seed & 0xFFFFFFFF, seed >> 32])
    np.testing.assert_array_equal(words.transpose(0, 2, 1), want)                    # bit-exact integer stream
    np.testing.assert_allclose(noise, noise_fill(B, R, 0.75, 0.75, seed, off), rtol=1e-13, atol=1e-15)
    import torch
    nd = eng.noise_fill(B, R, seed, off, on_device=True)
    np.testing.assert_array_equal(nd.cpu().numpy(), noise)
    eng.close()


@pytest.mark.gpu
def test_batch_entry_point_with_device_noise():
    """trajectorySimulateBatch(noise_rng='philox') == the same call fed the restated draws."""
    case = dict(Nx=10, sigma=0.4, noise_length=6, T_final=10)
    sc, mp, fp, _ = make_params(M, case)
    B = 16
    rng = np.random.default_rng(1)
    x0 = np.array([100., 10., 0, 0])[None, :] + np.concatenate([rng.uniform(-5, 5, (B, 2)), np.zeros((B, 2))], axis=1)
    a = M.trajectorySimulateBatch(sc, mp, fp, None, x0, seed=11, noise_rng="philox", lane_offset=100)
    nb = noise_fill(B, 20 // 6 + 1, 0.4, 0.4, 11, 100)
    b = M.trajectorySimulateBatch(sc, mp, fp, None, x0, nb)
    np.testing.assert_array_equal(a.i_term, b.i_term)
    np.testing.assert_allclose(a.x_true, b.x_true, rtol=1e-9, atol=1e-9, equal_nan=True)


@pytest.mark.gpu
def test_monte_carlo_drivers_with_device_noise():
    """success_rate / final_distance_ratio_sweep with noise_rng='philox': rejection and no-rejection lanes share their draws
    (same seed, same lanes) and the counts are consistent."""
    case = dict(Nx=10, sigma=0.5, noise_length=10, T_final=15)
    sc, mp, fp, _ = make_params(M, case)
    r = M.success_rate(sc, mp, fp, None, mc_num=64, seed=3, noise_rng="philox")
    assert r["runs"] == 64 and 0 <= r["success_count"] <= 64
    r2 = M.success_rate(sc, mp, fp, None, mc_num=64, seed=3, noise_rng="philox")
    assert r2 == r
    sw = M.final_distance_ratio_sweep(sc, mp, fp, None, (0.5, 0.5), [5, 10], mc_num=32, seed=1, noise_rng="philox")
    assert sw["dist_ratios"].shape == (2,) and np.all(np.isfinite(sw["dist_ratios"])) and np.all(sw["dist_ratios"] > 0)
