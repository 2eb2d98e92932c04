"""The lane-batched numpy restatement (the algebraic form the CUDA engine uses: reduced KKT,
spectral per-lane-rho operator, per-lane bound patches) against the scalar oracle
(OSQP-shaped KKT solve, ``oracle/sim_ref.py``) on identical inputs."""
import numpy as np
import pytest

import mpc_arpo_project_b200.mpcsim as M
from oracle.gen_golden import make_params
from oracle.sim_ref import trajectory_simulate
from oracle.batched_ref import simulate_discrete_batch

CASES = {
    "radial_nx10_sigma0.1": (dict(Nx=10, sigma=0.1, noise_length=5, T_final=20), 3),
    "radial_nx10_sigma0.75": (dict(Nx=10, sigma=0.75, noise_length=50, T_final=25), 3),
    "radial_nx10_nonoise": (dict(Nx=10, sigma=None, T_final=20), 2),
    "intrack_dv_nx20": (dict(Nx=20, inTrack=True, isDeltaV=True, isReject=False, sigma=None, T_final=15), 2),
    "radial_nx30_norej": (dict(Nx=30, sigma=0.7, noise_length=10, isReject=False, T_final=12), 2),
}


def lanes(case, B, seed):
    rng = np.random.default_rng(seed)
    base = np.array([-10., 100., 0, 0]) if case.get('inTrack') else np.array([100., 10., 0, 0])
    x0 = base[None, :] + np.concatenate([rng.uniform(-5, 5, (B, 2)), np.zeros((B, 2))], axis=1)
    nsim = int(case['T_final'] / 0.5)
    draws = rng.standard_normal((B, nsim + 2, 4))
    return x0, draws


@pytest.mark.parametrize("name", list(CASES))
def test_batched_form_equals_scalar_oracle(name):
    case, B = CASES[name]
    x0, draws = lanes(case, B, 7)
    sc, mp, fp, _ = make_params(M, case)
    sig = case.get('sigma') or 0.0
    nl = case.get('noise_length', 50)
    nb = np.ascontiguousarray((sig * draws[:, :, :2]).transpose(1, 2, 0))
    out = simulate_discrete_batch(sc, mp, fp, x0, nb)
    for b in range(B):
        sc.x0 = x0[b].copy()
        it = iter(draws[b])
        r = trajectory_simulate(sc, mp, fp, None, draw=lambda: next(it))
        assert out['i_term'][b] == r.i_term
        T = r.i_term
        assert list(out['iters'][:T, b]) == list(r.iters)
        assert list(out['status'][:T, b]) == list(r.status_val)
        # controls: tolerance of the task statement (1e-4 abs); observed ~1e-9
        np.testing.assert_allclose(out['ctrl_hist'][:T + 1, b].T, r.ctrl_hist[:, :T + 1], rtol=0, atol=1e-6)
        np.testing.assert_allclose(out['x_true'][:T + 1, b].T, r.x_true[:, :T + 1], rtol=1e-7, atol=1e-6)
        np.testing.assert_allclose(out['x_est'][:T + 1, b].T, r.x_est[:, :T + 1], rtol=1e-7, atol=1e-6)


def retype_lanes(B=12, seed=5, nsim=40, sigma=0.002, Nx=10):
    """Lanes that start within centimetres of the target: the velocity 1-norm bound ``|p^ - r|_1`` of the LOS rows then
    comes within RHO_TOL of its lower bound and OSQP re-types those rows as equalities."""
    rng = np.random.default_rng(seed)
    x0 = np.array([2.5, 0., 0, 0])[None, :] + np.concatenate(
        [rng.uniform(0.03, 0.4, (B, 1)), rng.uniform(-0.02, 0.02, (B, 1)), np.zeros((B, 2))], axis=1)
    noise = sigma * rng.standard_normal((nsim // 5 + 1, 2, B))
    return dict(Nx=Nx, sigma=sigma, noise_length=5, T_final=nsim * 0.5), x0, noise


def test_row_retyping_matches_the_kkt_oracle():
    """``prob.update(l, u)`` re-classifies a row whose scaled bounds come within 1e-4 of each other as an equality
    (rho_vec = 1e3 rho, KKT refactored; osqp auxil.c update_rho_vec, restated in oracle/osqp_ref.py::_set_rho_vec).
    The spectral form models it as a Woodbury correction of M(rho)^-1; on lanes that park next to the target it must
    give the scalar KKT oracle's iteration counts -- and NOT give them when the correction is switched off."""
    case, x0, noise = retype_lanes()
    B = x0.shape[0]
    sc, mp, fp, _ = make_params(M, case)
    out = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp')
    off = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp', retype=False)
    assert out["flip_flag"].sum() >= 6
    differ = 0
    for b in range(B):
        sc.x0 = x0[b].copy()
        draws = np.concatenate([noise[:, :, b] / case["sigma"], np.zeros((noise.shape[0], 2))], axis=1)
        it = iter(draws)
        r = trajectory_simulate(sc, mp, fp, None, draw=lambda: next(it, np.zeros(4)), chol_fail='clamp')
        T = r.i_term
        assert out['i_term'][b] == T
        assert list(out['iters'][:T, b]) == list(r.iters)
        assert list(out['status'][:T, b]) == list(r.status_val)
        np.testing.assert_allclose(out['ctrl_hist'][:T + 1, b].T, r.ctrl_hist[:, :T + 1], rtol=0, atol=1e-5)
        differ += list(off['iters'][:T, b]) != list(r.iters)
    assert differ >= 3, "the scenario no longer exercises re-typing"
