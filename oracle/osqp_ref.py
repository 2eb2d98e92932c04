"""Restatement of OSQP *as the reference uses it* (ORACLE -- test infrastructure).

[3P: ``osqp`` 0.6.x C core (``osqp_setup``, ``osqp_solve``, ``osqp_update_bounds``,
``osqp_update_A``, ``scale_data``, ``check_termination``, ``adapt_rho``) -- the
package is NOT vendored in /root/reference and cannot be installed here; this
file restates its published algorithm (Stellato et al., "OSQP: an operator
splitting solver for quadratic programs", 2020) and its 0.6.x control flow.
PARITY UNPINNED at this boundary: there are no golden vectors to check against.]

Reference call sites this replaces:
  ``src/trajectorySimulate.py:242-245`` (setup, ``warm_start=True, verbose=False``),
  ``:296`` (solve), ``:342`` (update l,u), ``:348`` (update Ax,l,u);
  ``src/trajectorySimulateC.py:269-272,338,399,405``.

Deliberate, documented deviations from OSQP 0.6.x:
  * dense storage (m<=406, n<=201), KKT solved by LU instead of QDLDL;
  * ``adaptive_rho_interval`` is a fixed setting (default 50) -- in 0.6.x the
    default 0 means "pick it from wall-clock time on the first solve", which is
    not reproducible; 50 is OSQP 1.x's fixed default (unverified offline);
  * no polishing (off by default in OSQP too), no timing.
"""
from types import SimpleNamespace

import numpy as np
import scipy.linalg as sla

RHO_MIN = 1e-6
RHO_MAX = 1e6
RHO_EQ_OVER_RHO_INEQ = 1e3
RHO_TOL = 1e-4
OSQP_INFTY = 1e30
MIN_SCALING = 1e-4
MAX_SCALING = 1e4
OSQP_DIVISION_TOL = 1.0 / OSQP_INFTY

# status_val codes (osqp/include/constants.h)
OSQP_SOLVED = 1
OSQP_SOLVED_INACCURATE = 2
OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3
OSQP_DUAL_INFEASIBLE_INACCURATE = 4
OSQP_MAX_ITER_REACHED = -2
OSQP_PRIMAL_INFEASIBLE = -3
OSQP_DUAL_INFEASIBLE = -4
OSQP_NON_CVX = -7
OSQP_UNSOLVED = -10

STATUS_STR = {
    OSQP_SOLVED: "solved",
    OSQP_SOLVED_INACCURATE: "solved inaccurate",
    OSQP_PRIMAL_INFEASIBLE_INACCURATE: "primal infeasible inaccurate",
    OSQP_DUAL_INFEASIBLE_INACCURATE: "dual infeasible inaccurate",
    OSQP_MAX_ITER_REACHED: "maximum iterations reached",
    OSQP_PRIMAL_INFEASIBLE: "primal infeasible",
    OSQP_DUAL_INFEASIBLE: "dual infeasible",
    OSQP_NON_CVX: "problem non convex",
    OSQP_UNSOLVED: "unsolved",
}

DEFAULT_SETTINGS = dict(
    rho=0.1, sigma=1e-6, alpha=1.6, max_iter=4000, eps_abs=1e-3, eps_rel=1e-3,
    eps_prim_inf=1e-4, eps_dual_inf=1e-4, scaling=10, adaptive_rho=True,
    adaptive_rho_interval=50, adaptive_rho_tolerance=5.0, check_termination=25,
    scaled_termination=False, warm_start=True, verbose=False,
)


def _limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


def ruiz_scale(P, q, A, iters):
    """``scale_data`` (osqp/src/scaling.c): ``iters`` Ruiz passes on [[P,A'],[A,0]]
    with the cost normalisation inside each pass.  Returns scaled copies and
    (D, E, c)."""
    P = np.array(P, float)
    A = np.array(A, float)
    q = np.array(q, float)
    n, m = P.shape[0], A.shape[0]
    D = np.ones(n)
    E = np.ones(m)
    c = 1.0
    for _ in range(iters):
        d_tmp = np.maximum(np.abs(P).max(axis=0), np.abs(A).max(axis=0) if m else 0.0)
        e_tmp = np.abs(A).max(axis=1) if m else np.zeros(0)
        d_tmp = 1.0 / np.sqrt(_limit_scaling(d_tmp))
        e_tmp = 1.0 / np.sqrt(_limit_scaling(e_tmp))
        P = d_tmp[:, None] * P * d_tmp[None, :]
        A = e_tmp[:, None] * A * d_tmp[None, :]
        q = d_tmp * q
        D = D * d_tmp
        E = E * e_tmp
        c_tmp = np.abs(P).max(axis=0).mean()
        inf_norm_q = float(_limit_scaling(np.array([np.abs(q).max()]))[0])
        c_tmp = max(c_tmp, inf_norm_q)
        c_tmp = float(_limit_scaling(np.array([c_tmp]))[0])
        c_tmp = 1.0 / c_tmp
        P = P * c_tmp
        q = q * c_tmp
        c = c * c_tmp
    return P, q, A, D, E, c


class OSQPRef:
    """Mimics the ``osqp.OSQP`` object protocol used by the reference:
    ``setup(P,q,A,l,u,**settings)``, ``solve()``, ``update(l=,u=,A=)``."""

    def __init__(self):
        self.work = None

    # ------------------------------------------------------------------ setup
    def setup(self, P, q, A, l, u, **settings):
        s = dict(DEFAULT_SETTINGS)
        # the python wrapper's default for warm_start is True as well
        s.update(settings)
        self.s = SimpleNamespace(**s)
        self.n = int(np.asarray(q).size)
        self.m = int(np.asarray(l).size)
        self.P0 = _dense(P)
        self.A0 = _dense(A)
        self.q0 = np.array(q, float)
        # python wrapper clamps infinities (osqp/interface.py)
        self.l0 = np.maximum(np.array(l, float), -OSQP_INFTY)
        self.u0 = np.minimum(np.array(u, float), OSQP_INFTY)
        self._scale()
        self.s.rho = min(max(self.s.rho, RHO_MIN), RHO_MAX)
        self.constr_type = np.zeros(self.m, int)
        self.rho_vec = np.zeros(self.m)
        self._set_rho_vec(force=True)
        self._factor()
        # cold start
        self.x = np.zeros(self.n)
        self.z = np.zeros(self.m)
        self.y = np.zeros(self.m)
        self.rho_updates = 0
        self.n_factor = 1
        return self

    def _scale(self):
        if self.s.scaling:
            self.P, self.q, self.A, self.D, self.E, self.c = ruiz_scale(self.P0, self.q0, self.A0, self.s.scaling)
        else:
            self.P, self.q, self.A = self.P0.copy(), self.q0.copy(), self.A0.copy()
            self.D, self.E, self.c = np.ones(self.n), np.ones(self.m), 1.0
        self.Dinv, self.Einv, self.cinv = 1.0 / self.D, 1.0 / self.E, 1.0 / self.c
        self.l = self.E * self.l0
        self.u = self.E * self.u0

    def _classify(self):
        ct = np.zeros(self.m, int)
        free = (self.l < -OSQP_INFTY * MIN_SCALING) & (self.u > OSQP_INFTY * MIN_SCALING)
        eq = (~free) & (self.u - self.l < RHO_TOL)
        ct[free] = -1
        ct[eq] = 1
        return ct

    def _set_rho_vec(self, force=False):
        """``set_rho_vec`` / ``update_rho_vec`` (osqp/src/auxil.c). Returns True when
        a constraint type changed (=> refactor)."""
        ct = self._classify()
        changed = force or bool(np.any(ct != self.constr_type))
        self.constr_type = ct
        self._fill_rho_vec()
        return changed

    def _fill_rho_vec(self):
        rho = self.s.rho
        self.rho_vec = np.where(self.constr_type == -1, RHO_MIN,
                                np.where(self.constr_type == 1, RHO_EQ_OVER_RHO_INEQ * rho, rho))
        self.rho_inv_vec = 1.0 / self.rho_vec

    def _factor(self):
        n, m = self.n, self.m
        K = np.zeros((n + m, n + m))
        K[:n, :n] = self.P + self.s.sigma * np.eye(n)
        K[:n, n:] = self.A.T
        K[n:, :n] = self.A
        K[n:, n:] = -np.diag(self.rho_inv_vec)
        self._lu = sla.lu_factor(K)
        self.n_factor = getattr(self, "n_factor", 0) + 1

    # ----------------------------------------------------------------- update
    def update(self, q=None, l=None, u=None, A=None):
        """``prob.update(l=,u=)`` / ``prob.update(Ax=,l=,u=)``.  ``A`` is the full dense
        matrix (the reference passes the CSC value array; same information)."""
        if A is not None:
            # osqp_update_A: unscale, overwrite, re-run Ruiz from scratch, refactor
            self.A0 = _dense(A)
            self._scale()          # also rescales l,u with the (new) E
            self._factor()
            self.rho_updates = 0
        if q is not None:
            self.q0 = np.array(q, float)
            self.q = self.c * self.D * self.q0
        if l is not None or u is not None:
            if l is not None:
                self.l0 = np.maximum(np.array(l, float), -OSQP_INFTY)
            if u is not None:
                self.u0 = np.minimum(np.array(u, float), OSQP_INFTY)
            self.l = self.E * self.l0
            self.u = self.E * self.u0
            self.rho_updates = 0      # reset_info
            if self._set_rho_vec():
                self._factor()

    def warm_start(self, x=None, y=None):
        if x is not None:
            self.x = self.Dinv * np.asarray(x, float)
            self.z = self.A @ self.x
        if y is not None:
            self.y = self.Einv * np.asarray(y, float) * self.c

    # ------------------------------------------------------------------ solve
    def _update_info(self):
        """``update_info`` -> ``compute_pri_res`` / ``compute_dua_res`` (auxil.c).  Keeps
        the *scaled* residual vectors, which ``compute_rho_estimate`` reuses."""
        self.Ax = self.A @ self.x
        self._pri_vec = self.Ax - self.z
        self.Px = self.P @ self.x
        self.Aty = self.A.T @ self.y
        self._dua_vec = self.q + self.Px + self.Aty
        if self.s.scaling and not self.s.scaled_termination:
            self.pri_res = _ninf(self.Einv * self._pri_vec)
            self.dua_res = self.cinv * _ninf(self.Dinv * self._dua_vec)
        else:
            self.pri_res = _ninf(self._pri_vec)
            self.dua_res = _ninf(self._dua_vec)

    def _pri_tol(self, eps_abs, eps_rel):
        if self.s.scaling and not self.s.scaled_termination:
            mx = max(_ninf(self.Einv * self.z), _ninf(self.Einv * self.Ax))
        else:
            mx = max(_ninf(self.z), _ninf(self.Ax))
        return eps_abs + eps_rel * mx

    def _dua_tol(self, eps_abs, eps_rel):
        if self.s.scaling and not self.s.scaled_termination:
            mx = max(_ninf(self.Dinv * self.q), _ninf(self.Dinv * self.Aty), _ninf(self.Dinv * self.Px))
            mx *= self.cinv
        else:
            mx = max(_ninf(self.q), _ninf(self.Aty), _ninf(self.Px))
        return eps_abs + eps_rel * mx

    def _is_primal_infeasible(self, eps):
        inf_u = self.u > OSQP_INFTY * MIN_SCALING
        inf_l = self.l < -OSQP_INFTY * MIN_SCALING
        dy = self.delta_y
        dy = np.where(inf_u & inf_l, 0.0, np.where(inf_u, np.minimum(dy, 0.0), np.where(inf_l, np.maximum(dy, 0.0), dy)))
        self.delta_y = dy            # OSQP projects in place
        if self.s.scaling and not self.s.scaled_termination:
            norm_dy = _ninf(self.E * dy)
        else:
            norm_dy = _ninf(dy)
        if norm_dy > OSQP_DIVISION_TOL:
            lhs = float(np.sum(self.u * np.maximum(dy, 0.0) + self.l * np.minimum(dy, 0.0)))
            if lhs < -eps * norm_dy:
                Atdy = self.A.T @ dy
                if self.s.scaling and not self.s.scaled_termination:
                    Atdy = self.Dinv * Atdy
                return _ninf(Atdy) < eps * norm_dy
        return False

    def _is_dual_infeasible(self, eps):
        dx = self.delta_x
        if self.s.scaling and not self.s.scaled_termination:
            norm_dx = _ninf(self.D * dx)
            cs = self.c
        else:
            norm_dx = _ninf(dx)
            cs = 1.0
        if norm_dx > OSQP_DIVISION_TOL:
            if float(self.q @ dx) < -cs * eps * norm_dx:
                Pdx = self.P @ dx
                if self.s.scaling and not self.s.scaled_termination:
                    Pdx = self.Dinv * Pdx
                if _ninf(Pdx) < cs * eps * norm_dx:
                    Adx = self.A @ dx
                    if self.s.scaling and not self.s.scaled_termination:
                        Adx = self.Einv * Adx
                    bad = ((self.u < OSQP_INFTY * MIN_SCALING) & (Adx > eps * norm_dx)) | \
                          ((self.l > -OSQP_INFTY * MIN_SCALING) & (Adx < -eps * norm_dx))
                    return not bool(np.any(bad))
        return False

    def _check_termination(self, approximate):
        s = self.s
        k = 10.0 if approximate else 1.0
        eps_abs, eps_rel = s.eps_abs * k, s.eps_rel * k
        eps_pinf, eps_dinf = s.eps_prim_inf * k, s.eps_dual_inf * k
        if self.pri_res > OSQP_INFTY or self.dua_res > OSQP_INFTY:
            self.status_val = OSQP_NON_CVX
            return True
        prim_ok = dual_ok = prim_inf = dual_inf = False
        if self.m == 0:
            prim_ok = True
        elif self.pri_res < self._pri_tol(eps_abs, eps_rel):
            prim_ok = True
        else:
            prim_inf = self._is_primal_infeasible(eps_pinf)
        if self.dua_res < self._dua_tol(eps_abs, eps_rel):
            dual_ok = True
        else:
            dual_inf = self._is_dual_infeasible(eps_dinf)
        if prim_ok and dual_ok:
            self.status_val = OSQP_SOLVED_INACCURATE if approximate else OSQP_SOLVED
            return True
        if prim_inf:
            self.status_val = OSQP_PRIMAL_INFEASIBLE_INACCURATE if approximate else OSQP_PRIMAL_INFEASIBLE
            return True
        if dual_inf:
            self.status_val = OSQP_DUAL_INFEASIBLE_INACCURATE if approximate else OSQP_DUAL_INFEASIBLE
            return True
        return False

    def _rho_estimate(self):
        """``compute_rho_estimate`` (auxil.c): uses the SCALED residual vectors."""
        pri = _ninf(self._pri_vec)
        dua = _ninf(self._dua_vec)
        pri /= (max(_ninf(self.z), _ninf(self.Ax)) + 1e-10)
        dua /= (max(_ninf(self.q), _ninf(self.Aty), _ninf(self.Px)) + 1e-10)
        est = self.s.rho * np.sqrt(pri / (dua + 1e-10))
        return min(max(est, RHO_MIN), RHO_MAX)

    def _adapt_rho(self):
        s = self.s
        rho_new = self._rho_estimate()
        self.rho_estimate = rho_new
        if rho_new > s.rho * s.adaptive_rho_tolerance or rho_new < s.rho / s.adaptive_rho_tolerance:
            s.rho = min(max(rho_new, RHO_MIN), RHO_MAX)
            self._fill_rho_vec()
            self._factor()
            self.rho_updates += 1

    def solve(self):
        s = self.s
        n, m = self.n, self.m
        if not s.warm_start:
            self.x[:] = 0.0
            self.z[:] = 0.0
            self.y[:] = 0.0
        self.status_val = OSQP_UNSOLVED
        it = 0
        can_check = False
        rhs = np.empty(n + m)
        for it in range(1, s.max_iter + 1):
            x_prev, z_prev = self.x, self.z
            # update_xz_tilde
            rhs[:n] = s.sigma * x_prev - self.q
            rhs[n:] = z_prev - self.rho_inv_vec * self.y
            sol = sla.lu_solve(self._lu, rhs)
            xt = sol[:n]
            zt = z_prev + self.rho_inv_vec * (sol[n:] - self.y)
            # update_x / update_z / update_y
            self.x = s.alpha * xt + (1.0 - s.alpha) * x_prev
            self.delta_x = self.x - x_prev
            zr = s.alpha * zt + (1.0 - s.alpha) * z_prev
            self.z = np.minimum(np.maximum(zr + self.rho_inv_vec * self.y, self.l), self.u)
            self.delta_y = self.rho_vec * (zr - self.z)
            self.y = self.y + self.delta_y

            can_check = bool(s.check_termination) and (it % s.check_termination == 0)
            if can_check:
                self._update_info()
                if self._check_termination(False):
                    break
            if s.adaptive_rho and s.adaptive_rho_interval and (it % s.adaptive_rho_interval == 0):
                if not can_check:
                    self._update_info()
                self._adapt_rho()
        if not can_check:
            self._update_info()
            self._check_termination(False)
        if self.status_val == OSQP_UNSOLVED:
            if not self._check_termination(True):
                self.status_val = OSQP_MAX_ITER_REACHED
        self.iter = it
        infeas = self.status_val in (OSQP_PRIMAL_INFEASIBLE, OSQP_PRIMAL_INFEASIBLE_INACCURATE,
                                     OSQP_DUAL_INFEASIBLE, OSQP_DUAL_INFEASIBLE_INACCURATE, OSQP_NON_CVX)
        if infeas:
            xs = np.full(n, np.nan)
            ys = np.full(m, np.nan)
        else:
            xs = self.D * self.x
            ys = self.E * self.y * self.cinv
        info = SimpleNamespace(status=STATUS_STR[self.status_val], status_val=self.status_val, iter=it,
                               pri_res=self.pri_res, dua_res=self.dua_res, rho_updates=self.rho_updates,
                               rho=self.s.rho)
        return SimpleNamespace(x=xs, y=ys, info=info)


def _dense(M):
    return np.array(M.toarray() if hasattr(M, "toarray") else M, float)


def _ninf(v):
    return float(np.abs(v).max()) if v.size else 0.0
