"""Linear-KF estimator variant: the oracle restatement on CPU, the engine against it on the GPU."""
import numpy as np
import pytest

import mpc_arpo_project_b200 as M
from oracle.gen_golden import make_params
from oracle.kf_ref import LinearKFRef
from oracle.sim_ref import build_setup, trajectory_simulate, trajectory_simulate_c


def test_kf_update_reproduces_the_measured_position_exactly():
    """R = 0: after update(z) the position estimate IS the measurement and its covariance block vanishes
    (misc/MPCrendezKALMANdisturb.py:263-266 with Co = [I2 0])."""
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.3))
    s = build_setup(sc, mp, fp, None)
    kf = LinearKFRef(s.Ao, s.Bou)
    rng = np.random.default_rng(3)
    kf.x = rng.normal(size=6)
    A = rng.normal(size=(6, 6))
    kf.P = A @ A.T + np.eye(6)
    kf.Q = 1e-3 * np.eye(6)
    for _ in range(4):
        kf.predict(rng.normal(size=2))
        z = rng.normal(size=2)
        kf.update(z)
        np.testing.assert_allclose(kf.x[:2], z, rtol=0, atol=1e-12)
        np.testing.assert_allclose(kf.P[:2, :], 0, atol=1e-12)
        np.testing.assert_allclose(kf.P, kf.P.T, atol=1e-12)


def test_kf_closed_loop_tracks_the_true_state():
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.3, noise_length=5, T_final=10))
    rng = np.random.default_rng(0)
    r = trajectory_simulate(sc, mp, fp, None, draw=lambda: rng.standard_normal(4), estimator='kf')
    T = r.i_term
    assert T == 20
    np.testing.assert_allclose(r.x_est[:2, 1:T + 1], r.x_true[:2, 1:T + 1], rtol=0, atol=1e-9)      # positions: measured exactly
    assert np.abs(r.x_est[2:4, 3:T + 1] - r.x_true[2:4, 3:T + 1]).max() < 0.7                       # velocities: within the disturbance


def test_unknown_estimator_is_rejected():
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.3))
    with pytest.raises(ValueError):
        M.build_problem(sc, mp, fp, None, M.SolverSettings(estimator="ekf"))


@pytest.mark.gpu
@pytest.mark.parametrize("case", [dict(Nx=10, sigma=0.4, noise_length=6, T_final=15),
                                  dict(Nx=20, sigma=0.3, noise_length=4, T_final=8),
                                  dict(Nx=40, sigma=0.3, noise_length=4, T_final=5)])
def test_discrete_closed_loop_with_linear_kf_matches_scalar_oracle(case):
    """Team kernel (Nx = 10, 20) and the round-based block-kernel path (Nx = 40) with estimator = 'kf'."""
    B = 3
    rng = np.random.default_rng(5)
    x0 = np.array([100., 10., 0, 0])[None, :] + np.concatenate([rng.uniform(-5, 5, (B, 2)), np.zeros((B, 2))], axis=1)
    sc, mp, fp, _ = make_params(M, case)
    nsim, nl, sig = int(case['T_final'] / 0.5), case['noise_length'], case['sigma']
    draws = rng.standard_normal((B, nsim + 2, 4))
    noise = np.ascontiguousarray((sig * draws[:, :nsim // nl + 1, :2]).transpose(1, 2, 0))
    got = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise, settings=M.SolverSettings(estimator="kf"))
    assert got.ukf_clamped.sum() == 0
    for b in range(B):
        sc.x0 = x0[b].copy()
        it = iter(draws[b])
        r = trajectory_simulate(sc, mp, fp, None, draw=lambda: next(it), estimator='kf')
        T = r.i_term
        assert got.i_term[b] == T
        assert list(got.iters[:T, b]) == list(r.iters)
        assert list(got.status[:T, b]) == list(r.status_val)
        np.testing.assert_allclose(got.ctrl_hist[:, :T + 1, b], r.ctrl_hist[:, :T + 1], rtol=0, atol=1e-6)
        np.testing.assert_allclose(got.x_true[:, :T + 1, b], r.x_true[:, :T + 1], rtol=1e-7, atol=1e-6)
        np.testing.assert_allclose(got.x_est[:, :T + 1, b], r.x_est[:, :T + 1], rtol=1e-7, atol=1e-6)


@pytest.mark.gpu
def test_continuous_closed_loop_with_linear_kf_matches_scalar_oracle():
    case = dict(Nx=10, sigma=0.0012, noise_length=4, T_cont=0.001, T_final=3, isDeltaV=False)
    B = 2
    rng = np.random.default_rng(9)
    x0 = np.array([100., 10., 0, 0])[None, :] + np.concatenate([rng.uniform(-5, 5, (B, 2)), np.zeros((B, 2))], axis=1)
    sc, mp, fp, _ = make_params(M, case)
    n_refresh = np.arange(0, 3, 0.5 * 4).size
    V = 0.0012 * rng.standard_normal((B, 2, n_refresh))
    got = M.trajectorySimulateCBatch(sc, mp, fp, None, x0, np.ascontiguousarray(V.transpose(2, 1, 0)),
                                     settings=M.SolverSettings(estimator="kf"))
    for b in range(B):
        sc.x0 = x0[b].copy()
        r = trajectory_simulate_c(sc, mp, fp, None, V=V[b], integrator='rk4', estimator='kf')
        ns = len(r.iters)
        assert got.i_term[b] == r.i_term
        assert list(got.iters[:ns, b]) == list(r.iters)
        assert list(got.status[:ns, b]) == list(r.status_val)
        np.testing.assert_allclose(got.u_raw[:, :ns, b], r.u_raw, rtol=0, atol=1e-6)
        np.testing.assert_allclose(got.x_est[:, :r.n_est, b], r.x_est[:, :r.n_est], rtol=1e-7, atol=1e-6)


@pytest.mark.gpu
def test_estimator_seam_with_linear_kf():
    """mpcb_ukf_step with estimator = 'kf': z is the measured position."""
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.3))
    s = build_setup(sc, mp, fp, None)
    p = M.build_problem(sc, mp, fp, None, M.SolverSettings(estimator="kf"))
    eng = M.Engine(p)
    B = 5
    rng = np.random.default_rng(2)
    x = rng.normal(size=(6, B)) + np.array([100, 10, 0, 0, 0, 0])[:, None]
    P = np.empty((36, B))
    for b in range(B):
        A = rng.normal(size=(6, 6))
        P[:, b] = (A @ A.T * 1e-2 + 1e-3 * np.eye(6)).ravel()
    u = rng.normal(size=(2, B)) * 0.1
    z = x[:2] + rng.normal(size=(2, B)) * 0.1
    x1, P1 = eng.ukf_step(x.copy(), P.copy(), u, z)
    for b in range(B):
        kf = LinearKFRef(s.Ao, s.Bou)
        kf.x, kf.P, kf.Q = x[:, b].copy(), P[:, b].reshape(6, 6).copy(), np.asarray(p.Qw)
        kf.predict(u[:, b])
        kf.update(z[:, b])
        np.testing.assert_allclose(x1[:, b], kf.x, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(P1[:, b].reshape(6, 6), kf.P, rtol=1e-8, atol=1e-10)
    eng.close()
