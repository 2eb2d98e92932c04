"""Lane-batched numpy restatement of the discrete closed loop (ORACLE -- test infrastructure).

Same algorithm as ``oracle.sim_ref.trajectory_simulate`` (reference
``src/trajectorySimulate.py:285-356``) for the debris-free problem family, but vectorised
over B independent trajectories and written in the *algebraic form the CUDA engine uses*:

  * reduced KKT: ``(P+sI+A'RA) xt = s*x - q + A'(R z - y)``, ``zt = A xt`` (identical to
    OSQP's quasi-definite solve, ``oracle/osqp_ref.py`` ``solve``);
  * spectral operator ``M(rho)^-1 = V diag(1/(1+rho*lam)) V'`` per sign variant, so every
    lane may carry its own adaptive rho;
  * only the data that change per step are per-lane: ``x_hat`` (rows 0..3), the velocity
    1-norm bound (rows ``row3``), the disturbance pin (last two rows), the sign variant.

``tests/test_batched_ref.py`` checks it against the scalar oracle; the GPU parity tests
check the CUDA engine against both.  It is also the faster CPU baseline for ``bench.py``.
"""
import numpy as np
import scipy.linalg as sla

from .osqp_ref import (ruiz_scale, OSQP_INFTY, MIN_SCALING, RHO_MIN, RHO_MAX, RHO_EQ_OVER_RHO_INEQ, RHO_TOL,
                       OSQP_DIVISION_TOL, OSQP_SOLVED, OSQP_SOLVED_INACCURATE, OSQP_PRIMAL_INFEASIBLE,
                       OSQP_PRIMAL_INFEASIBLE_INACCURATE, OSQP_MAX_ITER_REACHED, OSQP_UNSOLVED, DEFAULT_SETTINGS)
from .sim_ref import build_setup
from .ukf_ref import cholesky_psd_clamp


def _ninf(a):
    return np.abs(a).max(axis=1)


class BatchedQP:
    """Shared tables + per-lane ADMM state for B lanes."""

    def __init__(self, setup, B, settings=None, spectral=None):
        """``spectral``: optional ``(V[4,n,n], lam[4,n])`` to use instead of this class's own generalised eigen-decomposition.
        Two ``eigh`` calls on inputs that differ in the last bit reconstruct ``M(rho)^-1`` to 1e-11..1e-9 relative only (the
        pencil has condition ~1e6), which is the largest difference between the engine and this oracle; passing the engine's
        tables (``Problem.V, Problem.lam``, themselves checked in tests/test_problem.py) isolates the device arithmetic."""
        s = setup
        st = dict(DEFAULT_SETTINGS)
        st.update(settings or {})
        self.st = st
        assert st['adaptive_rho_interval'] % st['check_termination'] == 0 and st['max_iter'] % st['check_termination'] == 0
        self.n, self.m, self.B = s.P.shape[0], s.A.shape[0], B
        A0 = s.A.copy()
        nX = (s.Nx + 1) * 4
        self.sgn_rows = nX + 5 * np.arange(s.Nx + 1) + 3
        self.c1c, self.c2c = 4 * np.arange(s.Nx + 1) + 2, 4 * np.arange(s.Nx + 1) + 3
        A0[self.sgn_rows, self.c1c] = 1.0
        A0[self.sgn_rows, self.c2c] = 1.0
        self.row3 = self.sgn_rows[:s.Nb + 1]
        self.P, self.q, A_s, self.D, self.E, self.c = ruiz_scale(s.P, s.q, A0, st['scaling'])
        self.Dinv, self.Einv, self.cinv = 1 / self.D, 1 / self.E, 1 / self.c
        lt = self.E * np.maximum(s.l, -OSQP_INFTY)
        ut = self.E * np.minimum(s.u, OSQP_INFTY)
        self.inf_l, self.inf_u = lt < -OSQP_INFTY * MIN_SCALING, ut > OSQP_INFTY * MIN_SCALING
        ct = np.where(self.inf_l & self.inf_u, -1, np.where(ut - lt < RHO_TOL, 1, 0))
        ct[self.row3] = 0
        self.ctype = ct
        self.lt, self.ut = lt, ut
        self.Av, self.V, self.lam = [], [], []
        w = np.where(ct == 1, RHO_EQ_OVER_RHO_INEQ, 1.0)
        for v in range(4):
            A = A_s.copy()
            if v & 1:
                A[self.sgn_rows, self.c1c] *= -1
            if v & 2:
                A[self.sgn_rows, self.c2c] *= -1
            Af, Ac = A[ct == -1], A[ct != -1]
            Bm = self.P + st['sigma'] * np.eye(self.n) + RHO_MIN * Af.T @ Af
            Gm = Ac.T @ (w[ct != -1][:, None] * Ac)
            if spectral is None:
                lam, V = sla.eigh(0.5 * (Gm + Gm.T), 0.5 * (Bm + Bm.T))
            else:
                V, lam = np.asarray(spectral[0][v], float), np.asarray(spectral[1][v], float)
            self.Av.append(A)
            self.V.append(V)
            self.lam.append(np.maximum(lam, 0))
        self.x = np.zeros((B, self.n))
        self.z = np.zeros((B, self.m))
        self.y = np.zeros((B, self.m))
        self.rho = np.full(B, min(max(st['rho'], RHO_MIN), RHO_MAX))
        self.l = np.tile(lt, (B, 1))
        self.u = np.tile(ut, (B, 1))
        self.variant = np.zeros(B, int)
        self.flip_flag = np.zeros(B, bool)
        self.eqm = np.zeros((B, self.row3.size), bool)       # per lane: which velocity-bound rows are equalities right now
        self.retype = True
        self.dy = np.zeros((B, self.m))
        self.is_reject = 1.0

    def set_params(self, idx, xhat4, val, dhat2, variant):
        """The two ``prob.update`` calls of trajectorySimulate.py:340-348 for lanes ``idx``."""
        E = self.E
        self.l[idx, :4] = self.u[idx, :4] = -xhat4 * E[:4]
        self.u[np.ix_(idx, self.row3)] = val[:, None] * E[self.row3][None, :]
        eq = val[:, None] * E[self.row3][None, :] - self.lt[self.row3][None, :] < RHO_TOL
        self.flip_flag[idx] |= eq.any(axis=1)
        if self.retype:
            # OSQP update_rho_vec (auxil.c): a row whose scaled bounds come within RHO_TOL of each other is re-classified as an
            # equality (rho_vec = 1e3 rho) and the KKT matrix refactored; it returns to an inequality when they part again
            self.eqm[idx] = eq
        self.l[idx, -2:] = self.u[idx, -2:] = self.is_reject * dhat2 * E[-2:]
        self.variant[idx] = variant

    def _rho_vecs(self, rho, g=None):
        rv = np.where(self.ctype[None, :] == -1, RHO_MIN,
                      np.where(self.ctype[None, :] == 1, RHO_EQ_OVER_RHO_INEQ * rho[:, None], rho[:, None]))
        if g is not None and self.eqm[g].any():
            sub = rv[:, self.row3]
            rv[:, self.row3] = np.where(self.eqm[g], RHO_EQ_OVER_RHO_INEQ * rho[:, None], sub)
        return rv, 1.0 / rv

    def solve(self, idx):
        """Run ``prob.solve()`` for lanes ``idx``; returns (status_val, iters, u0 unscaled)."""
        st = self.st
        idx = np.asarray(idx)
        status = np.full(idx.size, OSQP_UNSOLVED)
        iters = np.zeros(idx.size, int)
        act = np.arange(idx.size)
        it = 0
        while act.size and it < st['max_iter']:
            g = idx[act]
            for v in range(4):
                sel = np.nonzero(self.variant[g] == v)[0]
                if sel.size:
                    self._block(g[sel], v, st['check_termination'])
            it += st['check_termination']
            done = np.zeros(act.size, bool)
            for v in range(4):
                sel = np.nonzero(self.variant[g] == v)[0]
                if not sel.size:
                    continue
                gl = g[sel]
                info = self._info(gl, v)
                sv = self._check(gl, v, info, False)
                fin = sv != OSQP_UNSOLVED
                if st['adaptive_rho'] and it % st['adaptive_rho_interval'] == 0:
                    self._adapt(gl[~fin], {k: a[~fin] for k, a in info.items()})
                if it >= st['max_iter']:
                    sva = self._check(gl, v, info, True)
                    sva = np.where(sva == OSQP_UNSOLVED, OSQP_MAX_ITER_REACHED, sva)
                    sv = np.where(fin, sv, sva)
                    fin = np.ones_like(fin)
                status[act[sel]] = np.where(fin, sv, status[act[sel]])
                iters[act[sel]] = it
                done[sel] = fin
            act = act[~done]
        return status, iters

    def _block(self, g, v, niter):
        st = self.st
        A, V, lam = self.Av[v], self.V[v], self.lam[v]
        x, z, y = self.x[g], self.z[g], self.y[g]
        rv, rinv = self._rho_vecs(self.rho[g], g)
        dscale = 1.0 / (1.0 + self.rho[g][:, None] * lam[None, :])
        l, u = self.l[g], self.u[g]
        # re-typed rows F of a lane add (1e3 - 1) rho a_i a_i' to M(rho): Woodbury on the spectral inverse S,
        #   (M + A_F' D A_F)^-1 = S - S A_F' (D^-1 + A_F S A_F')^-1 A_F S,  D = 999 rho I
        wood = []
        for k in np.nonzero(self.eqm[g].any(axis=1))[0]:
            AF = A[self.row3[self.eqm[g[k]]]]                                   # [f, n]
            U = (V * dscale[k][None, :]) @ (V.T @ AF.T)                         # S A_F'  [n, f]
            C = np.linalg.inv(np.eye(AF.shape[0]) / ((RHO_EQ_OVER_RHO_INEQ - 1.0) * self.rho[g[k]]) + AF @ U)
            wood.append((k, AF, U @ C))
        for _ in range(niter):
            r = st['sigma'] * x - self.q[None, :] + (rv * z - y) @ A
            xt = ((r @ V) * dscale) @ V.T
            for k, AF, UC in wood:
                xt[k] = xt[k] - UC @ (AF @ xt[k])
            zt = xt @ A.T
            x = st['alpha'] * xt + (1 - st['alpha']) * x
            zr = st['alpha'] * zt + (1 - st['alpha']) * z
            z = np.minimum(np.maximum(zr + rinv * y, l), u)
            dy = rv * (zr - z)
            y = y + dy
        self.x[g], self.z[g], self.y[g] = x, z, y
        self.dy[g] = dy

    def _info(self, g, v):
        A = self.Av[v]
        x, z, y = self.x[g], self.z[g], self.y[g]
        Ax, Px, Aty = x @ A.T, x @ self.P, y @ A
        pv, dv = Ax - z, self.q[None, :] + Px + Aty
        return dict(Ax=Ax, Px=Px, Aty=Aty, pv=pv, dv=dv, dy=self.dy[g],
                    pri=_ninf(self.Einv * pv), dua=self.cinv * _ninf(self.Dinv * dv))

    def _check(self, g, v, info, approx):
        st = self.st
        k = 10.0 if approx else 1.0
        z = self.z[g]
        eps_p = k * st['eps_abs'] + k * st['eps_rel'] * np.maximum(_ninf(self.Einv * z), _ninf(self.Einv * info['Ax']))
        qn = np.abs(self.Dinv * self.q).max()
        eps_d = k * st['eps_abs'] + k * st['eps_rel'] * self.cinv * np.maximum(
            qn, np.maximum(_ninf(self.Dinv * info['Aty']), _ninf(self.Dinv * info['Px'])))
        prim_ok, dual_ok = info['pri'] < eps_p, info['dua'] < eps_d
        # primal infeasibility certificate (osqp auxil.c is_primal_infeasible)
        eps_i = k * st['eps_prim_inf']
        dy = info['dy']
        dy = np.where((self.inf_u & self.inf_l)[None, :], 0.0,
                      np.where(self.inf_u[None, :], np.minimum(dy, 0), np.where(self.inf_l[None, :], np.maximum(dy, 0), dy)))
        ndy = _ninf(self.E * dy)
        lhs = (self.u[g] * np.maximum(dy, 0) + self.l[g] * np.minimum(dy, 0)).sum(axis=1)
        atdy = _ninf(self.Dinv * (dy @ self.Av[v]))
        pinf = (~prim_ok) & (ndy > OSQP_DIVISION_TOL) & (lhs < -eps_i * ndy) & (atdy < eps_i * ndy)
        out = np.full(g.size, OSQP_UNSOLVED)
        out[pinf] = OSQP_PRIMAL_INFEASIBLE_INACCURATE if approx else OSQP_PRIMAL_INFEASIBLE
        out[prim_ok & dual_ok] = OSQP_SOLVED_INACCURATE if approx else OSQP_SOLVED
        return out

    def _adapt(self, g, info):
        if not g.size:
            return
        st = self.st
        z = self.z[g]
        pri = _ninf(info['pv']) / (np.maximum(_ninf(z), _ninf(info['Ax'])) + 1e-10)
        dua = _ninf(info['dv']) / (np.maximum(np.abs(self.q).max(), np.maximum(_ninf(info['Aty']), _ninf(info['Px']))) + 1e-10)
        est = np.clip(self.rho[g] * np.sqrt(pri / (dua + 1e-10)), RHO_MIN, RHO_MAX)
        tol = st['adaptive_rho_tolerance']
        ch = (est > self.rho[g] * tol) | (est < self.rho[g] / tol)
        self.rho[g[ch]] = est[ch]


def simulate_discrete_batch(sc, mp, fp, x0_batch, noise_batch=None, settings=None, regen_sigmas=True, nsteps=None,
                            chol_fail='raise', spectral=None, retype=True):
    """Batched ``trajectorySimulate`` (debris-free).  ``x0_batch[B,4]``; ``noise_batch[R,2,B]``
    holds sigma-scaled position disturbances, refreshed every ``noise_length`` steps
    (R >= nsim//noise_length + 1).  Returns a dict of SoA arrays."""
    import copy
    B = x0_batch.shape[0]
    sc0 = copy.copy(sc)
    s = build_setup(sc0, mp, fp, None)
    nsim = int(sc.T_final / sc.time_stp) if nsteps is None else nsteps
    qp = BatchedQP(s, B, settings, spectral=spectral)
    qp.retype = retype          # False: the pre-round-2 behaviour (re-typed rows only flagged), kept for the sensitivity check in the tests
    qp.is_reject = 1.0 if sc.isReject else 0.0
    has_noise = sc.noise is not None
    nrep = s.noiseRepeat
    Ad, Bd, Ao, Bou = s.Ad, s.Bd, s.Ao, s.Bou
    xtrue = np.full((nsim + 1, B, 4), np.nan)
    xest = np.full((nsim + 1, B, 6), np.nan)
    ctrls = np.full((nsim + 1, B, 2), np.nan)
    seq = np.zeros((nsim, B), np.uint8)
    status = np.zeros((nsim, B), int)
    iters = np.zeros((nsim, B), int)
    u_raw = np.full((nsim, B, 2), np.nan)
    rho_hist = np.full((nsim, B), np.nan)          # rho each lane ended solve i with (parity reports: rho at a divergence)
    xtrue[0] = x0_batch
    xest[0, :, :4] = x0_batch
    xest[0, :, 4:] = 0
    ctrls[0] = 0
    iterm = np.full(B, nsim)
    alive = np.ones(B, bool)
    xintf = np.zeros(B)
    noise = np.zeros((B, 4))
    if has_noise:
        noise[:, :2] = noise_batch[0].T
    # UKF state
    ux = xest[0].copy()
    uP = np.tile(np.diag([1e-20] * 4 + [1.0] * 2), (B, 1, 1))
    Qw = np.zeros((6, 6))
    Qw[:4, :4] = 0.001 * np.eye(4)
    Qw[4, 4], Qw[5, 5] = (s.T * s.sigMat[0, 0]) ** 2, (s.T * s.sigMat[1, 1]) ** 2
    lam_u = 0.1 ** 2 * (6 - 1) - 6
    Wm = np.full(13, 0.5 / (6 + lam_u))
    Wc = Wm.copy()
    Wm[0] = lam_u / (6 + lam_u)
    Wc[0] = Wm[0] + (1 - 0.01 + 2.0)

    clamped = np.zeros(B, bool)

    def sigmas(x, P, idx):
        try:
            U = np.linalg.cholesky((6 + lam_u) * P).transpose(0, 2, 1)     # upper, rows U[k]
        except np.linalg.LinAlgError:
            # filterpy/scipy raise here and the reference run dies; with chol_fail='clamp' the lane
            # continues on the positive semi-definite factor like the engine (oracle/ukf_ref.py)
            if chol_fail != 'clamp':
                raise
            U = np.empty_like(P)
            for k in range(P.shape[0]):
                try:
                    U[k] = np.linalg.cholesky((6 + lam_u) * P[k]).T
                except np.linalg.LinAlgError:
                    U[k], _ = cholesky_psd_clamp((6 + lam_u) * P[k])
                    clamped[idx[k]] = True
        return np.concatenate([x[:, None, :], x[:, None, :] + U, x[:, None, :] - U], axis=1)

    def set_qp(idx, xe):
        val = np.abs(xe[:, 0] - s.xr[0]) + np.abs(xe[:, 1] - s.xr[1])
        var = (xe[:, 2] < 0).astype(int) + 2 * (xe[:, 3] < 0).astype(int)
        qp.set_params(idx, xe[:, :4], val, xe[:, 4:6], var)

    set_qp(np.arange(B), xest[0])
    xe_store = xest[0].copy()
    for i in range(nsim):
        xt = xtrue[i]
        chk = xt[:, 1] if sc.inTrack else xt[:, 0]
        term = alive & ((np.hypot(xt[:, 0], xt[:, 1]) < s.rp) | (chk < s.rp - s.rtot))
        iterm[term] = i
        alive &= ~term
        idx = np.nonzero(alive)[0]
        if not idx.size:
            break
        sv, its = qp.solve(idx)
        status[i, idx], iters[i, idx] = sv, its
        rho_hist[i, idx] = qp.rho[idx]
        xs = xe_store[idx, :4]
        solved = sv == OSQP_SOLVED
        xi = np.where(solved, 0.0, xintf[idx] + xs[:, 0] - s.xr[0])
        xintf[idx] = xi
        u_fs = -(xs @ s.Kpf.T) - xi[:, None] * s.Kif[:, 0][None, :]
        u_mpc = qp.x[idx][:, s.Nx * 4 + 4:s.Nx * 4 + 6] * qp.D[s.Nx * 4 + 4:s.Nx * 4 + 6][None, :]
        ctrl = np.where(solved[:, None], u_mpc, u_fs)
        u_raw[i, idx] = ctrl
        seq[i, idx] = np.where(solved, 1, 2)
        nrm = np.hypot(ctrl[:, 0], ctrl[:, 1])
        c0 = np.where(nrm > s.umax[0], ctrl[:, 0] * (s.umax[0] / np.where(nrm > 0, nrm, 1)), ctrl[:, 0])
        nrm2 = np.hypot(c0, ctrl[:, 1])
        c1 = np.where(nrm > s.umax[0], ctrl[:, 1] * (s.umax[0] / np.where(nrm2 > 0, nrm2, 1)), ctrl[:, 1])
        ctrls[i + 1, idx] = np.stack([c0, c1], axis=1)
        up = ctrls[i, idx]
        xn = xt[idx] @ Ad.T + up @ Bd.T + noise[idx]
        xtrue[i + 1, idx] = xn
        if has_noise:
            sg = sigmas(ux[idx], uP[idx], idx)
            sf = sg @ Ao.T + (up @ Bou.T)[:, None, :]
            xm = np.einsum('k,bkj->bj', Wm, sf)
            dfx = sf - xm[:, None, :]
            Pm = np.einsum('k,bki,bkj->bij', Wc, dfx, dfx) + Qw[None]
            if regen_sigmas:
                sf = sigmas(xm, Pm, idx)
            zs = np.stack([np.hypot(sf[:, :, 0], sf[:, :, 1]), np.arctan2(sf[:, :, 1], sf[:, :, 0])], axis=2)
            zp = np.einsum('k,bkj->bj', Wm, zs)
            dz = zs - zp[:, None, :]
            S = np.einsum('k,bki,bkj->bij', Wc, dz, dz)
            Pxz = np.einsum('k,bki,bkj->bij', Wc, sf - xm[:, None, :], dz)
            K = Pxz @ np.linalg.inv(S)
            zmeas = np.stack([np.hypot(xn[:, 0], xn[:, 1]), np.arctan2(xn[:, 1], xn[:, 0])], axis=1)
            ux[idx] = xm + np.einsum('bij,bj->bi', K, zmeas - zp)
            uP[idx] = Pm - K @ S @ K.transpose(0, 2, 1)
            xe = ux[idx].copy()
        else:
            xe = np.concatenate([xn, np.zeros((idx.size, 2))], axis=1)
        set_qp(idx, xe)
        if sc.inTrack:                     # in-place x/y swap of the stored estimate, simhelpers.py:72
            xe = xe.copy()
            xe[:, [0, 1]] = xe[:, [1, 0]]
        xest[i + 1, idx] = xe
        xe_store[idx] = xe
        if has_noise and (i + 1) % nrep == 0:
            noise[:, :2] = noise_batch[(i + 1) // nrep].T
    return dict(i_term=iterm, x_true=xtrue, x_est=xest, ctrl_hist=ctrls, ctrlr_seq=seq, status=status, iters=iters,
                u_raw=u_raw, rho=qp.rho.copy(), rho_hist=rho_hist, flip_flag=qp.flip_flag.copy(), ukf_clamped=clamped)
