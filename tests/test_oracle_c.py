"""The compiled oracle twin (oracle/c/mpc_ref.c, dense Cholesky of the reduced KKT system) against

* the lane-batched numpy oracle (spectral form) on the same seeded lanes: identical discrete record, controls to 1e-8;
* the fixtures captured from the reference's own driver code (tests/golden/ref_*.npz), through the same noise draws the
  reference's legacy RNG produces (seed 123, trajectorySimulate.py:28, 268, 352).

Three float64 implementations with independent linear algebra (numpy eigen-decomposition, C Cholesky, the CUDA kernels)
agreeing here is what the oracle's status "faithful, parity unpinned at the third-party boundary" rests on.
"""
import os

import numpy as np
import pytest

import mpc_arpo_project_b200 as M
from mpc_arpo_project_b200.presets import make_params
from oracle import c_ref
from oracle.batched_ref import simulate_discrete_batch
from oracle.gen_golden import CASES as GOLDEN_CASES
from conftest import GOLDEN

CASES = {
    "radial_nx10_sigma0.1": (dict(Nx=10, sigma=0.1, noise_length=5, T_final=20), 24),
    "radial_nx10_sigma0.75": (dict(Nx=10, sigma=0.75, noise_length=50, T_final=20), 24),
    "intrack_dv_nx20": (dict(Nx=20, inTrack=True, isDeltaV=True, isReject=False, sigma=None, T_final=12), 12),
    "radial_nx30_norej": (dict(Nx=30, sigma=0.7, noise_length=10, isReject=False, T_final=10), 8),
}


@pytest.fixture(scope="module", autouse=True)
def _built():
    c_ref.build()


@pytest.mark.parametrize("name", list(CASES))
def test_c_twin_matches_batched_oracle(name):
    case, B = CASES[name]
    sc, mp, fp, _ = make_params(case)
    rng = np.random.default_rng(21)
    base = np.array([-10., 100., 0, 0]) if case.get('inTrack') else np.array([100., 10., 0, 0])
    x0 = base[None, :] + np.concatenate([rng.uniform(-5, 5, (B, 2)), np.zeros((B, 2))], axis=1)
    nsim = int(case['T_final'] / 0.5)
    sig = case.get('sigma') or 0.0
    noise = sig * rng.standard_normal((nsim // case.get('noise_length', 50) + 1, 2, B)) if sig else None
    prob = M.build_problem(sc, mp, fp, None)
    got = c_ref.simulate_discrete(prob, np.ascontiguousarray(x0.T), noise, nsim, nthreads=4)
    ref = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp')
    assert np.array_equal(got["i_term"], ref["i_term"])
    assert np.array_equal(got["iters"], ref["iters"])
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["ctrlr_seq"], ref["ctrlr_seq"])
    assert got["qp_solves"] == int(ref["i_term"].sum()) and got["admm_iterations"] == int(ref["iters"].sum())
    for b in range(B):
        T = int(ref["i_term"][b])
        np.testing.assert_allclose(got["ctrl_hist"][:T + 1, b], ref["ctrl_hist"][:T + 1, b], rtol=0, atol=1e-7)
        np.testing.assert_allclose(got["x_true"][:T + 1, b], ref["x_true"][:T + 1, b], rtol=1e-8, atol=1e-7)
        np.testing.assert_allclose(got["x_est"][:T + 1, b], ref["x_est"][:T + 1, b], rtol=1e-8, atol=1e-7)


@pytest.mark.parametrize("name", [k for k, v in GOLDEN_CASES.items() if v[0] == 'D' and 'debris' not in k])
def test_c_twin_matches_reference_driver_fixture(name):
    g = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"), allow_pickle=False)
    case = GOLDEN_CASES[name][1]
    sc, mp, fp, _ = make_params(case)
    nsim = int(sc.T_final / sc.time_stp)
    np.random.seed(123)
    if sc.noise is not None:
        nl = int(sc.noise.noise_length)
        sigm = sc.noise.constructSigMat()
        draws = np.stack([sigm @ np.random.normal(0, 1, 4) for _ in range(nsim // nl + 1)])
        noise = np.ascontiguousarray(draws[:, :2, None])
    else:
        noise = None
    prob = M.build_problem(sc, mp, fp, None)
    r = c_ref.simulate_discrete(prob, np.asarray(sc.x0, float).reshape(4, 1), noise, nsim, nthreads=1)
    it = int(g["i_term"])
    assert int(r["i_term"][0]) == it
    assert list(r["iters"][:it, 0]) == list(g["solve_iter"][:it])
    np.testing.assert_allclose(r["ctrl_hist"][:it + 1, 0].T, g["ctrl_hist"][:, :it + 1], rtol=0, atol=1e-6)
    np.testing.assert_allclose(r["x_true"][:it, 0].T, g["x_true_pcw"], rtol=1e-7, atol=1e-6)
    np.testing.assert_allclose(r["x_est"][:it + 1, 0].T, g["x_est"][:, :it + 1], rtol=1e-7, atol=1e-6)
    np.testing.assert_array_equal(r["ctrlr_seq"][:it, 0].astype(float), g["ctrlr_seq"])
    assert bool(r["isSuccess"][0]) == bool(g["isSuccess"])


def test_c_twin_full_horizon_against_the_batched_oracle():
    """All 300 control steps of config 2's lanes: two CPU implementations with independent linear algebra (numpy
    eigen-decomposition of M(rho) vs the twin's dense Cholesky).  The closed loop is chaotic in OSQP's discrete decisions
    (tests/test_full_horizon_parity.py), so the two agree on a FRACTION of the lanes to the end -- this is the floor any
    float64 implementation, the CUDA engine included, is measured against -- and to 1e-6 in the controls until a lane's first
    differing decision."""
    from mpc_arpo_project_b200.presets import WORKLOADS, make_inputs
    from oracle.parity import as_engine_layout, full_horizon_report
    wl = WORKLOADS["config2"]
    sc, mp, fp, _ = make_params(wl["case"])
    B = 96
    x0, noise = make_inputs(wl, B, 4321)
    prob = M.build_problem(sc, mp, fp, None)
    twin = c_ref.simulate_discrete(prob, x0, noise, 300, nthreads=0)
    ref = simulate_discrete_batch(sc, mp, fp, np.ascontiguousarray(x0.T), noise, chol_fail='clamp')
    rep = full_horizon_report(as_engine_layout(twin), ref)
    print("C twin vs numpy oracle, config 2, 300 steps:", rep["exact_lanes"], "/", B, "max du on prefix", rep["max_du_prefix"])
    assert 0.45 <= rep["exact_frac"] < 1.0, rep           # measured 0.62 on 256 lanes; 1.0 would mean no sensitivity to show
    assert rep["max_du_prefix"] <= 1e-6, rep             # while rho agrees to 1e-6 as well (oracle/parity.py); measured 1e-7
    assert rep["solves_exact_prefix"] >= 0.75 * rep["solves_compared"], rep
    assert rep["i_term_equal_frac"] >= 0.9


def test_c_twin_models_row_retyping():
    """Lanes parked next to the target (tests/test_batched_ref.py::retype_lanes): the twin refactors with rho_vec = 1e3 rho on
    the re-typed rows exactly as OSQP's update_rho_vec does, and agrees with the Woodbury form of the numpy oracle."""
    from test_batched_ref import retype_lanes
    case, x0, noise = retype_lanes()
    sc, mp, fp, _ = make_params(case)
    prob = M.build_problem(sc, mp, fp, None)
    nsim = int(case["T_final"] / 0.5)
    got = c_ref.simulate_discrete(prob, np.ascontiguousarray(x0.T), noise, nsim, nthreads=4)
    ref = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp')
    assert ref["flip_flag"].sum() >= 6
    assert np.array_equal(got["i_term"], ref["i_term"]) and np.array_equal(got["iters"], ref["iters"])
    assert np.array_equal(got["status"], ref["status"])
    for b in range(x0.shape[0]):
        T = int(ref["i_term"][b])
        np.testing.assert_allclose(got["ctrl_hist"][:T + 1, b], ref["ctrl_hist"][:T + 1, b], rtol=0, atol=1e-5)
