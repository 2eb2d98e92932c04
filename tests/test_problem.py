"""Host setup of the product (mpc_arpo_project_b200/problem.py) against the oracle's restatement
of the reference setup (oracle/sim_ref.build_setup, itself pinned to the reference's own assembly by
tests/test_oracle_golden.py) and against OSQP's scaling / KKT solve as restated in oracle/osqp_ref.py."""
import numpy as np
import pytest

import mpc_arpo_project_b200 as M
from mpc_arpo_project_b200.problem import RHO_EQ_FACTOR, RHO_MIN
from oracle.gen_golden import make_params
from oracle.osqp_ref import ruiz_scale
from oracle.sim_ref import build_setup

CASES = [dict(Nx=10, sigma=0.75), dict(Nx=40, sigma=None), dict(Nx=20, inTrack=True, isDeltaV=True, isReject=False, sigma=None),
         dict(Nx=30, sigma=0.7, isReject=False)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"Nx{c['Nx']}{'_it' if c.get('inTrack') else ''}")
def test_tables_match_oracle_setup(case):
    sc, mp, fp, _ = make_params(M, case)
    s = build_setup(sc, mp, fp, None)
    p = M.build_problem(sc, mp, fp, None)
    np.testing.assert_allclose(p.Ad, s.Ad, rtol=0, atol=1e-15)
    np.testing.assert_allclose(p.Bd, s.Bd, rtol=0, atol=1e-15)
    np.testing.assert_allclose(p.P, s.P, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(p.q, s.q, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(p.A, s.A, rtol=1e-12, atol=1e-14)
    for a, b in ((p.l, s.l), (p.u, s.u)):
        assert np.array_equal(np.isinf(a), np.isinf(b))
        np.testing.assert_allclose(a[np.isfinite(a)], b[np.isfinite(b)], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(p.Kpf, s.Kpf, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(p.Kif, s.Kif, rtol=1e-9, atol=1e-12)
    Ps, qs, As, D, E, c = ruiz_scale(s.P, s.q, s.A, 10)
    np.testing.assert_allclose(p.D, D, rtol=1e-13)
    np.testing.assert_allclose(p.E, E, rtol=1e-13)
    assert p.c == pytest.approx(c, rel=1e-13)
    np.testing.assert_allclose(p.A_s, As, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("rho", [0.1, 0.687, 3.3e-3])
def test_spectral_operator_is_the_kkt_inverse(rho):
    """V diag(1/(1+rho*lam)) V' == (P + sigma I + A' diag(rho_vec) A)^-1 for every sign variant."""
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.1))
    p = M.build_problem(sc, mp, fp, None)
    rv = np.where(p.ctype == -1, RHO_MIN, np.where(p.ctype == 1, RHO_EQ_FACTOR * rho, rho))
    for v in range(4):
        A = p.A_variant(v)
        Mk = p.P_s + p.settings.sigma * np.eye(p.n) + A.T @ (rv[:, None] * A)
        Minv = p.V[v] @ np.diag(1.0 / (1.0 + rho * p.lam[v])) @ p.V[v].T
        np.testing.assert_allclose(Minv @ Mk, np.eye(p.n), rtol=0, atol=2e-9)


def test_debris_problem_carries_unscaled_data_and_deadbeat_gains():
    """Debris problems run on the per-lane path: no shared scaling / spectral tables, unscaled QP with the
    slope entries left at zero (the device fills them per lane and step), deadbeat gains as in the oracle."""
    case = dict(Nx=40, sigma=0.75, debris=((40., 0.), 5., 20))
    sc, mp, fp, debris = make_params(M, case)
    p = M.build_problem(sc, mp, fp, debris)
    s = build_setup(sc, mp, fp, debris)
    assert p.has_debris and p.A_s is None and p.V is None
    np.testing.assert_allclose(p.K_dead, s.K_total, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(p.Ki_dead, s.K_i.ravel(), rtol=1e-12, atol=1e-14)
    diff = np.argwhere(np.abs(s.A - p.A) > 1e-12)
    nX = 4 * (p.Nx + 1)
    assert len(diff) == p.Nx + 1 and all(r == nX + 5 * k + 4 and c == 4 * k for k, (r, c) in enumerate(diff))
    np.testing.assert_allclose(p.debris_verts, debris.constructVertArr())


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"Nx{c['Nx']}{'_it' if c.get('inTrack') else ''}")
def test_dual_infeasibility_certificate_cannot_fire(case):
    """OSQP's ``check_termination`` also tests a dual-infeasibility certificate (``is_dual_infeasible``, restated in
    oracle/osqp_ref.py): ``q'dx < -eps |dx|`` with ``|P dx| < eps |dx|`` (eps_dual_inf = 1e-4, norms unscaled).  The kernels do not
    carry it (they would need x of the previous iteration), and need not: every problem family of this path has P > 0 --
    state, input, slack AND disturbance variables are all weighted -- so ``|P dx|_inf >= lambda_min |dx|_inf / sqrt(n)``,
    three orders of magnitude above the threshold.  A status of DUAL_INFEASIBLE is unreachable for the reference as well."""
    sc, mp, fp, _ = make_params(M, case)
    p = M.build_problem(sc, mp, fp, None)
    P = np.asarray(p.P)
    lam_min = np.linalg.eigvalsh(0.5 * (P + P.T)).min()
    assert lam_min / np.sqrt(P.shape[0]) > 100 * 1e-4, lam_min
