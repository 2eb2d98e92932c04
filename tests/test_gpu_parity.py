"""GPU parity: the CUDA engine, called through the C ABI, against the CPU oracle on identical
inputs, and against the fixtures captured from the reference's own driver code.

Tolerances.  The ADMM iterates are float64 on both sides; the engine and the oracle differ only
in summation order, so discrete decisions (iteration counts, OSQP status, controller choice,
termination step) must agree exactly on these seeds, controls to 1e-6 abs (task bound: 1e-4)
and closed-loop states to 1e-6 abs / 1e-7 rel.
"""
import os

import numpy as np
import pytest

import mpc_arpo_project_b200 as M
from oracle.gen_golden import CASES as GOLDEN_CASES, make_params
from oracle.batched_ref import BatchedQP, simulate_discrete_batch
from oracle.sim_ref import build_setup, trajectory_simulate, trajectory_simulate_c, state_eqn_n
from oracle.ukf_ref import UKFRef, MerweScaledSigmaPointsRef
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

U_ATOL = 1e-6
X_ATOL, X_RTOL = 1e-6, 1e-7


def lanes(case, B, seed):
    rng = np.random.default_rng(seed)
    base = np.array([-10., 100., 0, 0]) if case.get('inTrack') else np.array([100., 10., 0, 0])
    x0 = base[None, :] + np.concatenate([rng.uniform(-5, 5, (B, 2)), np.zeros((B, 2))], axis=1)
    return x0, rng


# ------------------------------------------------------------------------------------ QP seam
@pytest.mark.parametrize("Nx", [10, 20, 30, 40])
def test_qp_solve_cold_and_warm(Nx):
    """mpcb_qp_solve == prob.update(l,u); prob.update(Ax,l,u); prob.solve() lane by lane."""
    _qp_seam(Nx, 48)


@pytest.mark.parametrize("solver,B", [("tile", 48), ("tile", 77), ("block", 48), ("team", 48), ("wave", 48), ("wave", 77)])
def test_qp_solve_every_solver_block(solver, B, monkeypatch):
    """The same seam with each of the solver blocks forced (MPCB_SOLVER): the multi-RHS wave kernel and the older DMMA tile
    kernel (8 lanes per warp; B = 77 leaves ragged tiles in every sign variant), the warp-per-lane block
    kernel and the persistent team kernel all reproduce the oracle's iterates."""
    monkeypatch.setenv("MPCB_SOLVER", solver)
    _qp_seam(10, B)


def _qp_seam(Nx, B):
    case = dict(Nx=Nx, sigma=0.1)
    sc, mp, fp, _ = make_params(M, case)
    rng = np.random.default_rng(Nx)
    s = build_setup(sc, mp, fp, None)
    qp = BatchedQP(s, B)
    eng = M.Engine(M.build_problem(sc, mp, fp, None))
    eng.batch_alloc(B)
    xh = np.zeros((B, 6))
    xh[:, 0] = 100 + rng.uniform(-10, 10, B)
    xh[:, 1] = 10 + rng.uniform(-5, 5, B)
    for rnd in range(3):
        if rnd:     # a plausible next-step estimate: small move, both velocity signs, a disturbance estimate
            xh[:, :2] += rng.normal(0, 0.05, (B, 2))
            xh[:, 2:4] = rng.normal(0, 0.1, (B, 2))
            xh[:, 4:6] = rng.normal(0, 0.01, (B, 2))
        val = np.abs(xh[:, 0] - s.xr[0]) + np.abs(xh[:, 1] - s.xr[1])
        var = (xh[:, 2] < 0).astype(int) + 2 * (xh[:, 3] < 0).astype(int)
        idx = np.arange(B)
        qp.set_params(idx, xh[:, :4], val, xh[:, 4:6], var)
        st_ref, it_ref = qp.solve(idx)
        u_ref = qp.x[:, s.Nx * 4 + 4:s.Nx * 4 + 6] * qp.D[s.Nx * 4 + 4:s.Nx * 4 + 6][None, :]
        u0, st, it = eng.qp_solve(np.ascontiguousarray(xh.T))
        assert np.array_equal(it, it_ref), f"round {rnd}"
        assert np.array_equal(st, st_ref), f"round {rnd}"
        np.testing.assert_allclose(u0.T, u_ref, rtol=0, atol=U_ATOL)
        x, z, y, rho = eng.qp_state(B // 2)
        np.testing.assert_allclose(x, qp.x[B // 2], rtol=1e-7, atol=1e-6)      # scaled iterates, O(1e2) entries
        np.testing.assert_allclose(y, qp.y[B // 2], rtol=1e-6, atol=1e-6)
        assert rho == pytest.approx(qp.rho[B // 2], rel=1e-6)
    eng.close()


# ------------------------------------------------------------------------------------ unit seams
def test_ukf_step_matches_oracle():
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.75))
    prob = M.build_problem(sc, mp, fp, None)
    eng = M.Engine(prob)
    B = 257
    rng = np.random.default_rng(3)
    x = np.zeros((6, B))
    x[0] = 100 + rng.uniform(-10, 10, B)
    x[1] = 10 + rng.uniform(-5, 5, B)
    x[2:4] = rng.normal(0, 0.2, (2, B))
    P = np.zeros((36, B))
    for i in range(6):
        P[i * 7] = 1e-20 if i < 4 else 1.0
    u = rng.uniform(-0.2, 0.2, (2, B))
    pts = MerweScaledSigmaPointsRef(6, alpha=0.1, beta=2., kappa=-1)
    xg, Pg = x, P
    xr_, Pr_ = x.copy(), P.copy()
    for step in range(4):
        xt = prob.Ad @ xg[:4] + prob.Bd @ u + rng.normal(0, 0.3, (4, B)) * np.array([1, 1, 0, 0])[:, None]
        z = np.stack([np.hypot(xt[0], xt[1]), np.arctan2(xt[1], xt[0])])
        xg, Pg = eng.ukf_step(xg, Pg, u, z)
        for b in range(0, B, 16):
            kf = UKFRef(6, 2, lambda s_, uu: prob.Ao @ s_ + prob.Bou @ uu,
                        lambda s_: np.array([np.hypot(s_[0], s_[1]), np.arctan2(s_[1], s_[0])]), pts)
            kf.x, kf.P, kf.Q, kf.R = xr_[:, b].copy(), Pr_[:, b].reshape(6, 6).copy(), prob.Qw, np.zeros((2, 2))
            kf.predict(u[:, b])
            kf.update(z[:, b])
            np.testing.assert_allclose(xg[:, b], kf.x, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(Pg[:, b].reshape(6, 6), kf.P, rtol=1e-7, atol=1e-12)
        xr_, Pr_ = xg.copy(), Pg.copy()
    eng.close()


def test_plant_steps_match_oracle():
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.1))
    prob = M.build_problem(sc, mp, fp, None)
    eng = M.Engine(prob)
    B = 300
    rng = np.random.default_rng(4)
    x = np.stack([100 + rng.uniform(-10, 10, B), 10 + rng.uniform(-5, 5, B), rng.normal(0, .2, B), rng.normal(0, .2, B)])
    u = rng.uniform(-0.2, 0.2, (2, B))
    w = rng.normal(0, 0.1, (2, B))
    got = eng.plant_lin_step(x, u, w)
    ref = prob.Ad @ x + prob.Bd @ u + np.vstack([w, np.zeros((2, B))])
    np.testing.assert_allclose(got, ref, rtol=1e-14, atol=1e-13)
    # RK4 substeps of the nonlinear plant (trajectorySimulateC.py:64-79)
    nsub, h = 50, 1e-3
    got = eng.plant_rk4(x, u, w * 1e-3, nsub, h)
    ref = x.copy()
    for b in range(0, B, 25):
        xb = x[:, b].copy()
        for _ in range(nsub):
            k1 = state_eqn_n(xb, u[:, b], prob.mean_mtn)
            k2 = state_eqn_n(xb + 0.5 * h * k1, u[:, b], prob.mean_mtn)
            k3 = state_eqn_n(xb + 0.5 * h * k2, u[:, b], prob.mean_mtn)
            k4 = state_eqn_n(xb + h * k3, u[:, b], prob.mean_mtn)
            xb = xb + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
            xb[:2] += w[:, b] * 1e-3
        np.testing.assert_allclose(got[:, b], xb, rtol=1e-12, atol=1e-11)
    eng.close()


# ------------------------------------------------------------------------------------ closed loops
DISCRETE = {
    "radial_nx10_sigma0.1": (dict(Nx=10, sigma=0.1, noise_length=5, T_final=20), 96),
    "radial_nx10_sigma0.75": (dict(Nx=10, sigma=0.75, noise_length=50, T_final=25), 96),
    "radial_nx10_nonoise": (dict(Nx=10, sigma=None, T_final=20), 40),
    "intrack_dv_nx20": (dict(Nx=20, inTrack=True, isDeltaV=True, isReject=False, sigma=None, T_final=15), 40),
    "radial_nx30_norej": (dict(Nx=30, sigma=0.7, noise_length=10, isReject=False, T_final=12), 40),
    "radial_nx40_sigma0.3": (dict(Nx=40, sigma=0.3, noise_length=7, T_final=8), 33),
}


def _compare_batch(got, ref, B):
    assert np.array_equal(got.i_term, ref['i_term'])
    assert np.array_equal(got.iters.astype(int), ref['iters'])
    assert np.array_equal(got.status.astype(int), ref['status'])
    assert np.array_equal(got.ctrlr_seq, ref['ctrlr_seq'])
    for b in range(B):
        T = int(ref['i_term'][b])
        np.testing.assert_allclose(got.ctrl_hist[:, :T + 1, b].T, ref['ctrl_hist'][:T + 1, b], rtol=0, atol=U_ATOL)
        np.testing.assert_allclose(got.u_raw[:, :T, b].T, ref['u_raw'][:T, b], rtol=0, atol=U_ATOL)
        np.testing.assert_allclose(got.x_true[:, :T + 1, b].T, ref['x_true'][:T + 1, b], rtol=X_RTOL, atol=X_ATOL)
        np.testing.assert_allclose(got.x_est[:, :T + 1, b].T, ref['x_est'][:T + 1, b], rtol=X_RTOL, atol=X_ATOL)


@pytest.mark.parametrize("name", list(DISCRETE))
def test_discrete_closed_loop_matches_batched_oracle(name):
    _closed_loop(name)


@pytest.mark.parametrize("solver", ["tile", "block", "wave"])
def test_discrete_closed_loop_every_solver_block(solver, monkeypatch):
    monkeypatch.setenv("MPCB_SOLVER", solver)
    monkeypatch.setenv("MPCB_RESUME_BELOW", "0")      # wave: stay on the rounds to the end instead of handing the lanes to the team kernel
    for name, (case, _) in DISCRETE.items():
        if case.get('Nx', 10) == 10:
            _closed_loop(name)


def _closed_loop(name):
    case, B = DISCRETE[name]
    x0, rng = lanes(case, B, 11)
    sc, mp, fp, _ = make_params(M, case)
    nsim = int(case['T_final'] / 0.5)
    sig = case.get('sigma') or 0.0
    nl = case.get('noise_length', 50)
    noise = sig * rng.standard_normal((nsim // nl + 1, 2, B))
    got = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise if sig else None)
    ref = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp')
    _compare_batch(got, ref, B)
    # Lanes where the reference's UKF would have raised (singular posterior since R = 0; a pivot of
    # -1e-18 vs +1e-18 is a coin flip of summation order, so the SET differs between engine and
    # oracle): both continue on the semi-definite factor and still agree to the tolerances above.
    assert got.ukf_clamped.sum() <= max(2, B // 8) and ref['ukf_clamped'].sum() <= max(2, B // 8)
    assert int(got.stats['qp_solves']) == int(ref['i_term'].sum())
    assert int(got.stats['admm_iterations']) == int(ref['iters'].sum())


def test_discrete_closed_loop_matches_scalar_oracle():
    """Three lanes against oracle/sim_ref.trajectory_simulate (OSQP-shaped KKT solve, per-step
    re-scaling and refactorisation like the real package)."""
    case = dict(Nx=10, sigma=0.4, noise_length=6, T_final=15)
    B = 3
    x0, rng = lanes(case, B, 5)
    sc, mp, fp, _ = make_params(M, case)
    nsim = 30
    draws = rng.standard_normal((B, nsim + 2, 4))
    noise = np.ascontiguousarray((0.4 * draws[:, :nsim // 6 + 1, :2]).transpose(1, 2, 0))
    got = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise)
    for b in range(B):
        sc.x0 = x0[b].copy()
        it = iter(draws[b])
        r = trajectory_simulate(sc, mp, fp, None, draw=lambda: next(it), chol_fail='clamp')
        T = r.i_term
        assert got.i_term[b] == T
        assert list(got.iters[:T, b]) == list(r.iters)
        assert list(got.status[:T, b]) == list(r.status_val)
        np.testing.assert_allclose(got.ctrl_hist[:, :T + 1, b], r.ctrl_hist[:, :T + 1], rtol=0, atol=U_ATOL)
        np.testing.assert_allclose(got.x_true[:, :T + 1, b], r.x_true[:, :T + 1], rtol=X_RTOL, atol=X_ATOL)
        np.testing.assert_allclose(got.x_est[:, :T + 1, b], r.x_est[:, :T + 1], rtol=X_RTOL, atol=X_ATOL)
        assert bool(got.isSuccess[b]) == bool(r.isSuccess)


@pytest.mark.parametrize("name", [k for k, v in GOLDEN_CASES.items() if v[0] == 'D'])
def test_drop_in_matches_reference_driver_fixture(name):
    """trajectorySimulate(...) -> SimRun against the fixture captured from the reference's
    src/trajectorySimulate.py (same legacy-RNG noise, seed 123)."""
    g = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"), allow_pickle=False)
    sc, mp, fp, debris = make_params(M, GOLDEN_CASES[name][1])
    r = M.trajectorySimulate(sc, mp, fp, debris)
    it = int(g["i_term"])
    assert r.i_term == it
    assert bool(r.isSuccess) == bool(g["isSuccess"])
    assert r.x_true_pcw.shape == g["x_true_pcw"].shape and r.x_est.shape == g["x_est"].shape
    assert r.ctrl_hist.shape == g["ctrl_hist"].shape and r.ctrlr_seq.shape == g["ctrlr_seq"].shape
    np.testing.assert_allclose(r.ctrl_hist[:, :it + 1], g["ctrl_hist"][:, :it + 1], rtol=0, atol=U_ATOL)
    np.testing.assert_allclose(r.x_true_pcw, g["x_true_pcw"], rtol=X_RTOL, atol=X_ATOL)
    np.testing.assert_allclose(r.x_est[:, :it + 1], g["x_est"][:, :it + 1], rtol=X_RTOL, atol=X_ATOL)
    np.testing.assert_array_equal(r.ctrlr_seq, g["ctrlr_seq"])
    np.testing.assert_allclose(r.noise_hist[:, :it + 1], g["noise_hist"][:, :it + 1], rtol=0, atol=1e-12)


@pytest.mark.parametrize("dv", [False, True])
def test_continuous_closed_loop_matches_scalar_oracle(dv):
    _continuous_vs_scalar(dict(Nx=10, sigma=0.0012, noise_length=4, T_cont=0.001, T_final=3, isDeltaV=dv))


@pytest.mark.parametrize("env", [
    {"MPCB_VISIT_ITERS": "25"},          # every solve longer than one check period spans several visits of the list-mode kernel
    {"MPCB_SCACHE_GB": "0"},             # no operator cache: every visit rebuilds S
    {"MPCB_SOLVER": "block"},            # the warp-per-lane block kernel under the same rounds
    {"MPCB_SOLVER": "tile"},             # DMMA tile kernel
    {"MPCB_TEAM": "regs"},               # register-resident operator
])
def test_continuous_closed_loop_solver_paths(env, monkeypatch):
    """The round-based (continuous) simulator gives the same trajectory whichever kernel solves its QPs and however
    the solves are cut into visits."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _continuous_vs_scalar(dict(Nx=10, sigma=0.0012, noise_length=4, T_cont=0.001, T_final=3, isDeltaV=False))


@pytest.mark.parametrize("Nx", [20, 30])
def test_continuous_closed_loop_longer_horizons(Nx):
    """n = 121 / 161: tensor-memory team kernels (2 and 1 teams per SM) in list mode, long disturbance columns."""
    _continuous_vs_scalar(dict(Nx=Nx, sigma=0.0012, noise_length=4, T_cont=0.001, T_final=2, isDeltaV=False))


@pytest.mark.parametrize("case", [
    dict(Nx=10, sigma=0.0012, noise_length=4, T_cont=0.001, T_final=3, debris=((60., 0.), 5., 20)),
    # test/traj_eval_radialC.py's obstacle (40, 0) side 5 detect 20, shorter horizon and run
    dict(Nx=20, sigma=0.0012, noise_length=4, T_cont=0.001, T_final=2, debris=((40., 0.), 5., 20)),
], ids=["nx10", "radialC_script_debris_nx20"])
def test_continuous_closed_loop_with_debris(case):
    """trajectorySimulateC WITH a Debris object (what test/traj_eval_radialC.py calls): per-lane path, RK4 plant."""
    _continuous_vs_scalar(case)


def _continuous_vs_scalar(case):
    B = 2
    x0, rng = lanes(case, B, 9)
    sc, mp, fp, debris = make_params(M, case)
    Tf = case['T_final']
    nsimD, nsimC, ratio = int(Tf / 0.5), int(Tf / 0.001), 500
    n_refresh = np.arange(0, Tf, 0.5 * 4).size
    V = 0.0012 * rng.standard_normal((B, 2, n_refresh))
    noise = np.ascontiguousarray(V.transpose(2, 1, 0))
    got = M.trajectorySimulateCBatch(sc, mp, fp, debris, x0, noise)
    for b in range(B):
        sc.x0 = x0[b].copy()
        r = trajectory_simulate_c(sc, mp, fp, debris, V=V[b], integrator='rk4', chol_fail='clamp')
        assert got.i_term[b] == r.i_term
        ns = len(r.iters)
        assert list(got.iters[:ns, b]) == list(r.iters)
        assert list(got.status[:ns, b]) == list(r.status_val)
        np.testing.assert_allclose(got.u_raw[:, :ns, b], r.u_raw, rtol=0, atol=U_ATOL)
        np.testing.assert_allclose(got.x_est[:, :r.n_est, b], r.x_est[:, :r.n_est], rtol=X_RTOL, atol=X_ATOL)
        # column j+1 of x_true: the plant state right after the j-th solve's substep
        for j, i_sub in enumerate(r.solve_at):
            np.testing.assert_allclose(got.x_true[:, j + 1, b], r.x_true[:, i_sub + 1], rtol=X_RTOL, atol=X_ATOL)
        assert bool(got.isSuccess[b]) == bool(r.isSuccess)


# ------------------------------------------------------------------------------------ full-size properties
def test_full_size_config2_properties():
    """BASELINE config 2 (4096 lanes, Nx=10, sigma=0.75 held 50 steps, 300 steps): size-independent
    properties -- the recorded trajectory replays through the plant model from the recorded controls
    and noise; applied controls respect the (component-wise effective) clip; lanes are independent of their position in
    the batch (duplicated lanes give bit-identical results)."""
    case = dict(Nx=10, sigma=0.75, noise_length=50, T_final=150)
    sc, mp, fp, _ = make_params(M, case)
    B, nsim = 4096, 300
    rng = np.random.default_rng(1234)
    x0 = np.stack([100 + rng.uniform(-10, 10, B), 10 + rng.uniform(-5, 5, B), np.zeros(B), np.zeros(B)], axis=1)
    noise = 0.75 * rng.standard_normal((7, 2, B))
    x0[B // 2:] = x0[:B // 2][::-1]                 # second half = first half, reversed order
    noise[:, :, B // 2:] = noise[:, :, :B // 2][:, :, ::-1]
    got = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise)
    prob = M.build_problem(sc, mp, fp, None)
    it = got.i_term
    assert it.min() >= 1 and it.max() <= nsim
    assert np.array_equal(it[B // 2:], it[:B // 2][::-1])
    assert np.array_equal(got.ctrl_hist[:, :, B // 2:], got.ctrl_hist[:, :, :B // 2][:, :, ::-1], equal_nan=True)
    assert np.array_equal(got.x_true[:, :, B // 2:], got.x_true[:, :, :B // 2][:, :, ::-1], equal_nan=True)
    steps = np.arange(nsim)[:, None]
    live = steps < it[None, :]                                  # [nsim, B] step i executed
    # the reference's clip is sequential (u0 scaled by the norm, then u1 by the NEW norm, :317-319):
    # it bounds each component, not the norm
    assert np.all(np.abs(got.ctrl_hist[:, 1:][:, live]) <= 0.2 * (1 + 1e-12))
    w = noise[np.minimum(np.arange(nsim) // 50, 6)]             # [nsim, 2, B]
    pred = np.einsum('ij,jtb->itb', prob.Ad, got.x_true[:, :-1]) + np.einsum('ij,jtb->itb', prob.Bd, got.ctrl_hist[:, :-1])
    pred[:2] += w.transpose(1, 0, 2)
    err = np.abs(pred - got.x_true[:, 1:])
    assert np.nanmax(np.where(live[None], err, 0.0)) < 1e-9
    assert int(got.stats['qp_solves']) == int(it.sum())
    assert set(np.unique(got.ctrlr_seq[live])) <= {1, 2}
    # lanes that came within ~2 cm of the target: OSQP would re-type the velocity-bound rows as
    # equalities there (u - l < 1e-4 after scaling), which the shared operator does not model; they
    # are counted, not hidden
    assert int(got.stats['flip_lanes']) <= B // 500


@pytest.mark.parametrize("name", [k for k, v in GOLDEN_CASES.items() if v[0] == 'C'])
def test_drop_in_continuous_matches_reference_driver_fixture(name):
    """trajectorySimulateC(...) -> SimRun (telemetry at every T_cont substep, the reference's array
    shapes) against the fixture captured from the reference's src/trajectorySimulateC.py (legacy-RNG
    noise with seed 321; the fixture keeps every 50th substep)."""
    g = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"), allow_pickle=False)
    sc, mp, fp, debris = make_params(M, GOLDEN_CASES[name][1])
    np.random.seed(321)
    r = M.trajectorySimulateC(sc, mp, fp, debris)
    it = int(g["i_term"])
    assert r.i_term == it
    assert bool(r.isSuccess) == bool(g["isSuccess"])
    nsimC = int(sc.T_final / sc.T_cont)
    assert r.x_true_pcw.shape == (4, it) and r.ctrl_hist.shape == (2, nsimC) and r.ctrlr_seq.shape == (it,)
    assert r.x_est.shape == g["x_est"].shape
    np.testing.assert_allclose(r.x_true_pcw[:, ::50], g["x_true_pcw"], rtol=X_RTOL, atol=X_ATOL)
    np.testing.assert_allclose(r.ctrl_hist[:, :it:50], g["ctrl_hist"], rtol=0, atol=U_ATOL)
    np.testing.assert_array_equal(r.ctrlr_seq[::50], g["ctrlr_seq"])
    nd = int(np.isfinite(g["x_est"][0]).sum())
    fin = np.isfinite(g["noise_hist"])
    np.testing.assert_allclose(r.noise_hist[fin], g["noise_hist"][fin], rtol=0, atol=1e-12)
    # the estimate columns the reference filled (one per executed solve + the initial one)
    ns = (it - 1) // int(sc.time_stp / sc.T_cont)
    np.testing.assert_allclose(r.x_est[:, :ns], g["x_est"][:, :ns], rtol=X_RTOL, atol=X_ATOL)


def test_continuous_full_rate_telemetry_matches_scalar_oracle():
    case = dict(Nx=10, sigma=0.0012, noise_length=2, T_cont=0.001, T_final=2, isDeltaV=False)
    sc, mp, fp, _ = make_params(M, case)
    rng = np.random.default_rng(21)
    x0 = np.array([[99.0, 9.0, 0., 0.], [101.5, 11.0, 0., 0.]])
    B = 2
    n_refresh = np.arange(0, 2, 0.5 * 2).size
    V = 0.0012 * rng.standard_normal((B, 2, n_refresh))
    got = M.trajectorySimulateCBatch(sc, mp, fp, None, x0, np.ascontiguousarray(V.transpose(2, 1, 0)),
                                     record=("x_true_sub", "ctrl_sub", "ctrlr_sub", "x_est"))
    for b in range(B):
        sc.x0 = x0[b].copy()
        r = trajectory_simulate_c(sc, mp, fp, None, V=V[b], integrator='rk4', chol_fail='clamp')
        it = r.i_term
        assert got.i_term[b] == it
        np.testing.assert_allclose(got.x_true_sub[:, :it, b], r.x_true_pcw, rtol=X_RTOL, atol=X_ATOL)
        np.testing.assert_allclose(got.ctrl_sub[:, :it, b], r.ctrl_hist[:, :it], rtol=0, atol=U_ATOL)
        seq = got.ctrlr_sub[:it, b].astype(float)
        seq[-1] = seq[-2]
        np.testing.assert_array_equal(seq, r.ctrlr_seq)


# ------------------------------------------------------------------------------------ edge cases
def test_edge_cases_single_lane_short_runs_and_immediate_termination():
    case = dict(Nx=10, sigma=0.1, noise_length=5, T_final=5)
    sc, mp, fp, _ = make_params(M, case)
    # lane 0 starts inside the platform radius (terminates before its first solve, i_term = 0),
    # lane 1 starts behind the platform (x < r_p - r_tol), lane 2 is a normal lane
    x0 = np.array([[1.0, 0.5, 0., 0.], [0.5, 8.0, 0., 0.], [100., 10., 0., 0.]])
    noise = 0.1 * np.random.default_rng(2).standard_normal((3, 2, 3))
    got = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise)
    ref = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp')
    assert list(got.i_term) == list(ref['i_term']) and got.i_term[0] == 0 and got.i_term[1] == 0
    assert not got.isSuccess[0] and not got.isSuccess[1]
    assert np.isnan(got.final_dist[0])                   # x_true_pcw[:, -1] of an empty trajectory
    T = int(ref['i_term'][2])
    np.testing.assert_allclose(got.ctrl_hist[:, :T + 1, 2].T, ref['ctrl_hist'][:T + 1, 2], rtol=0, atol=U_ATOL)
    # a single lane, a single control step
    one = M.trajectorySimulateBatch(sc, mp, fp, None, x0[2:3], noise[:, :, 2:3], nsteps=1)
    assert one.batch == 1 and one.i_term[0] == 1 and one.ctrl_hist.shape == (2, 2, 1)
    np.testing.assert_allclose(one.ctrl_hist[:, 1, 0], ref['ctrl_hist'][1, 2], rtol=0, atol=U_ATOL)
    # zero steps: nothing is solved, nothing terminates
    zero = M.trajectorySimulateBatch(sc, mp, fp, None, x0[2:3], noise[:, :, 2:3], nsteps=0)
    assert zero.i_term[0] == 0 and int(zero.stats['qp_solves']) == 0


def test_bad_arguments_are_reported_not_crashed():
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.1))
    eng = M.Engine(M.build_problem(sc, mp, fp, None))
    with pytest.raises(M._lib.MpcbError):                  # noisy problem without a noise tensor
        eng.simulate_discrete(np.zeros((4, 8)), None, 5)
    with pytest.raises(ValueError):                        # wrong shape
        eng.simulate_discrete(np.zeros((3, 8)), np.zeros((2, 2, 8)), 5)
    with pytest.raises(M._lib.MpcbError):
        eng.batch_alloc(0)
    eng.close()
    mp.Nx = 50                                             # a horizon the kernels are not instantiated for
    with pytest.raises(M._lib.MpcbError):
        M.Engine(M.build_problem(sc, mp, fp, None))


# ------------------------------------------------------------------------------------ debris lanes (per-lane path)
@pytest.mark.parametrize("case,tol", [
    (dict(Nx=10, sigma=0.3, noise_length=6, T_final=12, debris=((60., 0.), 5., 20)), 1.0),
    # lanes that start next to a vertex of the box: slope = dy/dx of the steering line amplifies rounding
    # differences of the estimate by 1/dx^2, so this case is held to 2e-5 instead of 1e-6
    (dict(Nx=10, sigma=None, T_final=10, debris=((95., 9.), 6., 20)), 20.0),
    (dict(Nx=20, sigma=0.2, noise_length=5, T_final=6, debris=((80., 4.), 8., 30)), 1.0),
    # test/traj_eval_in_track.py as shipped (u_lim supplied): in-track approach, debris at (0, 40), Nx = 40
    (dict(Nx=40, inTrack=True, isReject=False, sigma=None, T_final=6, debris=((0., 40.), 5., 20)), 1.0),
    (dict(Nx=10, inTrack=True, isDeltaV=True, isReject=False, sigma=0.2, noise_length=5, T_final=25, debris=((0., 70.), 5., 20)), 1.0),
], ids=["radial_noise", "debris_at_start", "nx20", "in_track_script_nx40", "in_track_dv_noise"])
def test_debris_lanes_match_scalar_oracle(case, tol):
    """Debris-avoidance lanes: per-step constraint geometry, OSQP re-scaling and refactorisation on the device
    (csrc/generic.cuh) against oracle/sim_ref.trajectory_simulate with the same debris, lane by lane.

    Parity is asserted step by step until OSQP's own arithmetic stops being well conditioned: once adaptive
    rho has run up past 1e2 (infeasible stretches drive it to RHO_MAX = 1e6, i.e. 1e9 on equality rows) the
    KKT system has condition 1e9..1e12 and the iteration path depends on the linear solver's rounding (LU of the
    KKT matrix in the oracle, explicit inverse of the reduced matrix here; real OSQP's QDLDL would differ
    from both).  Up to that point every count and status must be equal and the telemetry equal to `tol`."""
    B = 3
    x0, rng = lanes(case, B, 17)
    sc, mp, fp, debris = make_params(M, case)
    nsim = int(case['T_final'] / 0.5)
    sig = case.get('sigma') or 0.0
    nl = case.get('noise_length', 50)
    draws = rng.standard_normal((B, nsim + 2, 4))
    noise = np.ascontiguousarray((sig * draws[:, :nsim // nl + 1, :2]).transpose(1, 2, 0)) if sig else None
    got = M.trajectorySimulateBatch(sc, mp, fp, debris, x0, noise)
    for b in range(B):
        sc.x0 = x0[b].copy()
        it = iter(draws[b])
        r = trajectory_simulate(sc, mp, fp, debris, draw=lambda: next(it), chol_fail='clamp')
        T = r.i_term
        wild = np.nonzero(np.asarray(r.rho) > 1e2)[0]          # rho AFTER the solve of that step
        Tc = int(wild[0]) + 1 if wild.size else T               # steps 0..Tc-1 are compared
        assert Tc >= min(T, 8), "the well-conditioned prefix is too short to mean anything"
        if Tc == T:
            assert got.i_term[b] == T
        assert list(got.iters[:Tc, b]) == list(r.iters[:Tc])
        assert list(got.status[:Tc, b]) == list(r.status_val[:Tc])
        np.testing.assert_array_equal(got.ctrlr_seq[:Tc, b].astype(float), r.ctrlr_seq[:Tc])
        np.testing.assert_allclose(got.ctrl_hist[:, :Tc + 1, b], r.ctrl_hist[:, :Tc + 1], rtol=0, atol=tol * U_ATOL)
        np.testing.assert_allclose(got.x_true[:, :Tc + 1, b], r.x_true[:, :Tc + 1], rtol=X_RTOL, atol=tol * X_ATOL)
        np.testing.assert_allclose(got.x_est[:, :Tc + 1, b], r.x_est[:, :Tc + 1], rtol=X_RTOL, atol=tol * X_ATOL)


def test_results_are_reproducible_run_to_run():
    """Same inputs, same outputs, bit for bit: lanes are scheduled dynamically (atomic queue, two teams per SM) but no
    floating-point result may depend on it.  Debris lanes included (their M = P + sigma I + A'RA was once accumulated
    with atomics)."""
    for case, B in ((dict(Nx=10, sigma=0.75, noise_length=50, T_final=40), 600),
                    (dict(Nx=10, sigma=0.5, noise_length=6, T_final=15, debris=((60., 0.), 5., 20)), 40)):
        x0, rng = lanes(case, B, 23)
        sc, mp, fp, debris = make_params(M, case)
        nsim = int(case['T_final'] / 0.5)
        noise = case['sigma'] * rng.standard_normal((nsim // case['noise_length'] + 1, 2, B))
        runs = [M.trajectorySimulateBatch(sc, mp, fp, debris, x0, noise) for _ in range(3)]
        for r in runs[1:]:
            np.testing.assert_array_equal(r.i_term, runs[0].i_term)
            np.testing.assert_array_equal(r.iters, runs[0].iters)
            np.testing.assert_array_equal(r.status, runs[0].status)
            assert np.array_equal(r.x_true, runs[0].x_true, equal_nan=True)
            assert np.array_equal(r.x_est, runs[0].x_est, equal_nan=True)
            assert np.array_equal(r.ctrl_hist, runs[0].ctrl_hist, equal_nan=True)


def test_operator_cache_does_not_change_results(monkeypatch):
    """In-track delta-v lanes flip their velocity-sign variant at unchanged rho all the time (config 4): the Nx = 20 team
    kernel then reloads parked operators from its L2 cache instead of rebuilding them.  A reloaded operator is the rebuilt
    one bit for bit, so the run with the cache disabled must give identical telemetry."""
    case = dict(Nx=20, inTrack=True, isDeltaV=True, isReject=False, sigma=None, T_final=40)
    B = 64
    x0, rng = lanes(case, B, 31)
    sc, mp, fp, _ = make_params(M, case)
    a = M.trajectorySimulateBatch(sc, mp, fp, None, x0, None)
    monkeypatch.setenv("MPCB_NO_OCACHE", "1")
    b = M.trajectorySimulateBatch(sc, mp, fp, None, x0, None)
    assert a.stats["qp_solves"] == b.stats["qp_solves"] and a.stats["admm_iterations"] == b.stats["admm_iterations"]
    np.testing.assert_array_equal(a.iters, b.iters)
    assert np.array_equal(a.x_true, b.x_true, equal_nan=True) and np.array_equal(a.ctrl_hist, b.ctrl_hist, equal_nan=True)


@pytest.mark.parametrize("Nx,B,env", [(10, 48, {}), (20, 24, {}), (30, 16, {}),
                                      (10, 200, {"MPCB_SOLVER": "wave", "MPCB_RESUME_BELOW": "0"})],
                         ids=["nx10", "nx20", "nx30", "nx10_wave_rounds_defer_to_team"])
def test_row_retyping_on_the_team_kernel(Nx, B, env, monkeypatch):
    """lanes parked next to the target make OSQP re-type the velocity-bound rows as equalities
    (rho_vec = 1e3 rho + refactor).  The team kernel folds the re-typed rows into its tensor-memory operator with one
    Sherman-Morrison step per row; iteration counts, statuses and controls must follow the oracle that models the
    re-typing (checked against the scalar KKT oracle in tests/test_batched_ref.py) -- and differ from the one that does not."""
    from test_batched_ref import retype_lanes
    for k, v in env.items():         # last case: rounds of the multi-RHS wave kernel, which hands re-typing lanes to the team kernel
        monkeypatch.setenv(k, v)
    case, x0, noise = retype_lanes(B=B, seed=11, Nx=Nx)
    sc, mp, fp, _ = make_params(M, case)
    got = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise)
    ref = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp')
    off = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp', retype=False)
    # (the oracle also flags a lane whose LAST parameter refresh -- after its final solve -- would re-type a row)
    assert B // 3 <= int(got.stats["flip_lanes"]) <= int(ref["flip_flag"].sum())
    gi, gs = got.iters.astype(int), got.status.astype(int)
    same = np.array([np.array_equal(gi[:, b], ref['iters'][:, b]) and np.array_equal(gs[:, b], ref['status'][:, b]) and
                     np.array_equal(got.ctrlr_seq[:, b], ref['ctrlr_seq'][:, b]) and got.i_term[b] == ref['i_term'][b] for b in range(B)])
    same_off = np.array([np.array_equal(gi[:, b], off['iters'][:, b]) for b in range(B)])
    # lanes parked at the target run 800..4000-iteration solves at rho_vec = 1e3 rho: 2 of 200 lanes take a 25-iteration decision
    # the other way (on the team kernel and on the wave + team path alike); without the re-typing model a third of them would
    assert same.sum() >= B - max(1, B // 50) and same_off.sum() < same.sum(), (same.mean(), same_off.mean())
    for b in np.nonzero(same)[0]:      # rows scaled by E ~ 1e-3 carrying rho_vec = 1e3 rho: the task bar 1e-4 here; measured 1.3e-5 on the
        T = int(ref['i_term'][b])      # team kernel, 7.8e-5 (1 entry of 200 lanes) on wave rounds, which iterate in unscaled variables
        np.testing.assert_allclose(got.ctrl_hist[:, :T + 1, b].T, ref['ctrl_hist'][:T + 1, b], rtol=0, atol=1e-4)
        # states accumulate the control differences over up to 40 steps: 1e-4 (measured 1.2e-5)
        np.testing.assert_allclose(got.x_true[:, :T + 1, b].T, ref['x_true'][:T + 1, b], rtol=0, atol=1e-4)
        np.testing.assert_allclose(got.x_est[:, :T + 1, b].T, ref['x_est'][:T + 1, b], rtol=0, atol=1e-4)
    assert not np.array_equal(off["iters"], ref["iters"]), "the scenario no longer exercises re-typing"


# ------------------------------------------------------------------------------------ Monte-Carlo drivers
def test_monte_carlo_reductions_match_per_lane_results():
    """disturbRejComp / success_rates_test as batched calls: the device-side statistics equal what the
    per-lane outputs give, and rejection / no-rejection lanes share their noise realisation."""
    case = dict(Nx=10, sigma=0.5, noise_length=10, T_final=20)
    sc, mp, fp, _ = make_params(M, case)
    out = M.final_distance_ratio_sweep(sc, mp, fp, None, (0.5, 0.5), [5., 20.], mc_num=24, seed=3)
    assert out["dist_ratios"].shape == (2,) and np.all(np.isfinite(out["dist_ratios"]))
    # recompute the first setting by hand through the batch API
    rng = np.random.default_rng(3)
    R = 40 // 5 + 1
    noise = rng.standard_normal((R, 2, 24)) * 0.5
    x0 = np.tile(np.asarray(sc.x0, float)[None, :], (24, 1))
    means = []
    for rej in (False, True):
        sc.isReject = rej
        sc.noise = M.Noise((0.5, 0.5), 5)
        run = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise)
        means.append(run.final_dist.mean())
        assert run.stats["sum_final_dist"] == pytest.approx(run.final_dist.sum(), rel=1e-12)
    assert out["dist_ratios"][0] == pytest.approx(means[1] / means[0], rel=1e-12)
    sc.isReject = True
    sr = M.success_rate(sc, mp, fp, None, mc_num=16, seed=1)
    assert sr["runs"] == 16 and 0 <= sr["success_count"] <= 16


def test_batch_pipeline_equals_single_engine():
    """pipeline.BatchPipeline keeps several engines (handles, streams) of one problem family in flight, each driven by its own
    host thread.  Lanes are independent and every engine is deterministic: results must equal a single engine's, batch by batch."""
    from mpc_arpo_project_b200.pipeline import BatchPipeline
    case = dict(Nx=10, sigma=0.75, noise_length=10, T_final=15)
    sc, mp, fp, _ = make_params(M, case)
    prob = M.build_problem(sc, mp, fp, None)
    nsim = 30
    batches = []
    for k in range(5):
        x0, rng = lanes(case, 300 + 10 * k, 100 + k)
        noise = 0.75 * rng.standard_normal((nsim // 10 + 1, 2, x0.shape[0]))
        batches.append((np.ascontiguousarray(x0.T), noise))
    with M.Engine(prob) as eng:
        ref = [eng.simulate_discrete(x0, nz, nsim, ("x_true", "ctrl", "iters")) for x0, nz in batches]
    with BatchPipeline(prob, depth=3) as pipe:
        got = pipe.map_discrete(batches, nsim, ("x_true", "ctrl", "iters"))
    for a, b in zip(got, ref):
        assert a.stats["qp_solves"] == b.stats["qp_solves"] and a.stats["sum_final_dist"] == b.stats["sum_final_dist"]
        np.testing.assert_array_equal(a.iters, b.iters)
        assert np.array_equal(a.x_true, b.x_true, equal_nan=True) and np.array_equal(a.ctrl_hist, b.ctrl_hist, equal_nan=True)
