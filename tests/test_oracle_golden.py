"""Oracle restatement vs fixtures captured from the reference's own driver code.

``tests/golden/ref_*.npz`` were produced by ``oracle/gen_golden.py``: the reference's
unmodified ``src/trajectorySimulate.py`` / ``trajectorySimulateC.py`` / ``simhelpers.py``
running in the build container.  This pins the oracle's QP assembly (reference
``trajectorySimulate.py:216-236``, ``simhelpers.py:11-172``) and closed-loop driver logic
(``:285-387``; ``trajectorySimulateC.py:325-445``).
"""
import os

import numpy as np
import pytest

import mpc_arpo_project_b200.mpcsim as M
from oracle.gen_golden import CASES, make_params
from oracle.sim_ref import trajectory_simulate, trajectory_simulate_c
from conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, f"ref_{name}.npz"), allow_pickle=False)


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v[0] == 'D'])
def test_discrete_matches_reference_driver(name):
    g = _load(name)
    sc, mp, fp, debris = make_params(M, CASES[name][1])
    r = trajectory_simulate(sc, mp, fp, debris, use_sympy=True)   # literal sympy+quad discretisation (:101-109)
    s = r.setup
    # setup payload of osqp.setup (trajectorySimulate.py:245)
    np.testing.assert_allclose(s.P, g["P"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(s.q, g["q"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(s.A, g["A"], rtol=1e-10, atol=1e-12)
    for a, b in ((s.l, g["l"]), (s.u, g["u"])):
        assert np.array_equal(np.isinf(a), np.isinf(b))
        fin = np.isfinite(a)
        np.testing.assert_allclose(a[fin], b[fin], rtol=1e-12, atol=1e-12)
    # closed loop
    assert r.i_term == int(g["i_term"])
    assert bool(r.isSuccess) == bool(g["isSuccess"])
    assert list(r.iters) == list(g["solve_iter"])
    it = r.i_term
    np.testing.assert_allclose(r.ctrl_hist[:, :it + 1], g["ctrl_hist"][:, :it + 1], rtol=0, atol=1e-9)
    np.testing.assert_allclose(r.x_est[:, :it + 1], g["x_est"][:, :it + 1], rtol=0, atol=1e-8)
    np.testing.assert_allclose(r.x_true_pcw, g["x_true_pcw"], rtol=0, atol=1e-8)
    np.testing.assert_array_equal(r.ctrlr_seq, g["ctrlr_seq"])
    np.testing.assert_allclose(r.noise_hist[:, :it + 1], g["noise_hist"][:, :it + 1], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v[0] == 'C'])
def test_continuous_matches_reference_driver(name):
    g = _load(name)
    sc, mp, fp, debris = make_params(M, CASES[name][1])
    np.random.seed(321)
    r = trajectory_simulate_c(sc, mp, fp, debris, integrator="rk45", use_sympy=True)
    assert r.i_term == int(g["i_term"])
    assert list(r.iters) == list(g["solve_iter"])
    it = r.i_term
    np.testing.assert_allclose(r.x_true_pcw[:, :it:50], g["x_true_pcw"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(r.ctrl_hist[:, :it:50], g["ctrl_hist"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(r.ctrlr_seq[:it:50], g["ctrlr_seq"])
    nd = r.n_est
    np.testing.assert_allclose(r.x_est[:, :nd], g["x_est"][:, :nd], rtol=0, atol=1e-8)
    fin = np.isfinite(g["noise_hist"])
    np.testing.assert_allclose(r.noise_hist[fin], g["noise_hist"][fin], rtol=0, atol=1e-12)


def test_rk4_equals_rk45_on_short_run():
    """fixed-step RK4 at h=T_cont reproduces solve_ivp's RK45."""
    sc, mp, fp, debris = make_params(M, dict(CASES['contC_nx10_accel'][1], T_final=2))
    np.random.seed(5)
    a = trajectory_simulate_c(sc, mp, fp, debris, integrator='rk45')
    np.random.seed(5)
    b = trajectory_simulate_c(sc, mp, fp, debris, integrator='rk4')
    assert a.i_term == b.i_term
    np.testing.assert_allclose(a.x_true_pcw, b.x_true_pcw, rtol=0, atol=1e-9)
