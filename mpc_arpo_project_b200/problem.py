"""Host-side, once-per-problem setup for the batched engine.

Everything here runs once on the CPU when a problem is created (the reference redoes it on
every ``trajectorySimulate`` call, ``src/trajectorySimulate.py:73-245``) and produces the
constant tables the CUDA kernels share across the whole batch:

* CW model + discretisation (ref ``:73-111``), observer model (``:114-118``), LOS-cone
  matrix (``:133-156``), virtual-LQR terminal cost / tail gain (``:175-177``), failsafe
  integral-LQR gains (``:180-187``; python-control ``dlqr(integral_action=)`` restated);
* the QP ``P, q, A, l, u`` of ``:216-236`` + ``src/simhelpers.py:109-113,137-138``,
  written down directly from the block structure instead of through
  ``scipy.sparse`` kron/stack calls;
* OSQP's Ruiz equilibration ``D, E, c`` (the reference gets it from ``osqp.setup`` and
  again from every ``prob.update(Ax=...)``; |A| does not depend on the per-step signs, so
  one equilibration serves the whole run and the whole batch);
* the *spectral KKT operator*: OSQP refactors ``[[P+sI, A'],[A, -1/rho]]`` whenever rho
  adapts, and rho is per-trajectory.  With ``M(rho) = B + rho*G`` (``B = P+sigma*I+rho_min*Af'Af``,
  ``G = Ac' W Ac``) the generalised eigen-decomposition ``G V = B V diag(lam)``, ``V'BV = I``
  gives ``M(rho)^-1 = V diag(1/(1+rho*lam)) V'`` for EVERY rho, so the batch shares one
  dense ``V`` per sign variant and each ADMM linear solve is two dense mat-vecs.

Four sign variants exist because row 3 of every LOS block carries ``sign(vx_hat), sign(vy_hat)``
(``simhelpers.py:66-67,106-107``): variant = (C1<0) + 2*(C2<0).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import scipy.linalg as sla

OSQP_INFTY = 1e30
MIN_SCALING, MAX_SCALING = 1e-4, 1e4
RHO_MIN, RHO_MAX, RHO_EQ_FACTOR, RHO_TOL = 1e-6, 1e6, 1e3, 1e-4


@dataclass
class SolverSettings:
    """OSQP settings as the reference leaves them (defaults of osqp 0.6.x, ref
    ``trajectorySimulate.py:245`` passes only ``warm_start=True, verbose=False``).
    ``adaptive_rho_interval`` is fixed (OSQP 0.6's default picks it from wall-clock time)."""
    rho: float = 0.1
    sigma: float = 1e-6
    alpha: float = 1.6
    max_iter: int = 4000
    eps_abs: float = 1e-3
    eps_rel: float = 1e-3
    eps_prim_inf: float = 1e-4
    eps_dual_inf: float = 1e-4
    scaling: int = 10
    adaptive_rho: bool = True
    adaptive_rho_interval: int = 50
    adaptive_rho_tolerance: float = 5.0
    check_termination: int = 25
    # State estimator (not an OSQP setting; carried here because every entry point already takes ``settings``):
    # "ukf" = the reference's range/bearing UKF (``trajectorySimulate.py:121-130, 329-337``), "kf" = the linear Kalman
    # filter on position measurements of the reference's prototype (``misc/MPCrendezKALMANdisturb.py:261-266``).
    estimator: str = "ukf"


def _dense(M):
    return np.array(M.toarray() if hasattr(M, "toarray") else M, dtype=float)


def cw_discretise(n: float, T: float, delta_v: bool):
    """Zero-order-hold discretisation of the planar CW equations (ref ``:73-111``).
    Acceleration inputs: Van-Loan block exponential (the reference integrates a sympy
    matrix exponential entry by entry; equal to 1e-12).  Impulsive delta-v: ``Bd = Ad[:, 2:]``."""
    Ap = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [3 * n * n, 0, 0, 2 * n], [0, 0, -2 * n, 0]], float)
    Bp = np.array([[0, 0], [0, 0], [1, 0], [0, 1]], float)
    blk = np.zeros((6, 6))
    blk[:4, :4], blk[:4, 4:] = Ap, Bp
    e = sla.expm(blk * T)
    Ad = sla.expm(Ap * T)
    Bd = Ad @ np.vstack([np.zeros((2, 2)), np.eye(2)]) if delta_v else e[:4, 4:]
    return Ad, Bd


def _lqr_gain(A, B, Q, R):
    X = sla.solve_discrete_are(A, B, Q, R)
    return np.linalg.solve(R + B.T @ X @ B, B.T @ X @ A), X


def _ackermann(A, B, poles):
    """Single-input pole placement by Ackermann's formula (``ct.acker`` in the reference,
    ``src/trajectorySimulate.py:198``): K = e_n' ctrb(A,B)^-1 p(A)."""
    n = A.shape[0]
    B = np.asarray(B, float).reshape(n, 1)
    ctrb = np.hstack([np.linalg.matrix_power(A, i) @ B for i in range(n)])
    coef = np.real(np.poly(poles))                   # p(s) = s^n + c1 s^(n-1) + ... + cn
    pA = sum(coef[i] * np.linalg.matrix_power(A, n - i) for i in range(n + 1))
    return np.linalg.solve(ctrb, pA)[-1:, :]


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


def ruiz_equilibrate(P, q, A, passes: int):
    """OSQP ``scale_data``: Ruiz passes on the KKT matrix with per-pass cost normalisation."""
    P, q, A = P.copy(), q.copy(), A.copy()
    D, E, c = np.ones(P.shape[0]), np.ones(A.shape[0]), 1.0
    for _ in range(passes):
        dn = _limit(np.maximum(np.abs(P).max(0), np.abs(A).max(0)))
        en = _limit(np.abs(A).max(1))
        d, e = 1 / np.sqrt(dn), 1 / np.sqrt(en)
        P = d[:, None] * P * d[None, :]
        A = e[:, None] * A * d[None, :]
        q = d * q
        D, E = D * d, E * e
        qn = np.abs(q).max()
        qn = 1.0 if qn < MIN_SCALING else min(qn, MAX_SCALING)
        ct = max(np.abs(P).max(0).mean(), qn)
        ct = 1.0 if ct < MIN_SCALING else min(ct, MAX_SCALING)
        inv = 1.0 / ct
        P, q, c = P * inv, q * inv, c * inv
    return P, q, A, D, E, c


def _ukf_process_noise(scale, sc):
    """UKF process noise (ref :272-275 / C-variant :310-313)."""
    sig = np.array(sc.noise.noise_std, float) if sc.noise is not None else np.zeros(2)
    Qw = np.zeros((6, 6))
    Qw[:4, :4] = 0.001 * np.eye(4)
    Qw[4, 4], Qw[5, 5] = (scale * sig[0]) ** 2, (scale * sig[1]) ** 2
    return Qw


@dataclass
class Problem:
    """All constant tables of one problem family (one (SimConditions, MPCParams,
    FailsafeParams) triple).  Scaled quantities carry a ``_s`` suffix."""
    Nx: int
    Nc: int
    Nb: int
    n: int
    m: int
    in_track: bool
    delta_v: bool
    is_reject: bool
    has_noise: bool
    T: float
    mean_mtn: float
    r_p: float
    r_tol: float
    xr: np.ndarray
    umax0: float
    suc_dist: float
    suc_ang: float
    sig: np.ndarray              # (sigma_x, sigma_y)
    noise_length: int
    Ad: np.ndarray
    Bd: np.ndarray
    Ao: np.ndarray
    Bou: np.ndarray
    Qw: np.ndarray
    Kpf: np.ndarray
    Kif: np.ndarray
    C: np.ndarray
    P: np.ndarray
    q: np.ndarray
    A: np.ndarray                # variant 0 (C1=C2=+1), unscaled
    l: np.ndarray
    u: np.ndarray
    D: np.ndarray
    E: np.ndarray
    c: float
    P_s: np.ndarray
    q_s: np.ndarray
    A_s: np.ndarray              # variant 0, scaled
    l_s: np.ndarray              # scaled template bounds (dynamic rows hold placeholders)
    u_s: np.ndarray
    ctype: np.ndarray            # -1 free, 0 inequality, 1 equality (OSQP constr_type)
    sgn_rows: np.ndarray         # rows carrying C1/C2
    sgn_c1_col: np.ndarray
    sgn_c2_col: np.ndarray
    row3: np.ndarray             # rows whose upper bound is E*val (k=0..Nb)
    V: np.ndarray                # [4, n, n] generalised eigenvectors per sign variant
    lam: np.ndarray              # [4, n]
    settings: SolverSettings = field(default_factory=SolverSettings)
    u_off: int = 0               # offset of u_0 in the decision vector
    # debris-avoidance lanes (src/mpcsim.py:99-123): per-lane path, no shared tables
    has_debris: bool = False
    estimator: int = 0           # 0 = UKF, 1 = linear KF (include/mpcb.h MPCB_EST_*)
    debris_center: np.ndarray = field(default_factory=lambda: np.zeros(2))
    debris_side: float = 0.0
    debris_detect: float = 0.0
    debris_verts: np.ndarray = field(default_factory=lambda: np.zeros((4, 2)))
    K_dead: np.ndarray = field(default_factory=lambda: np.zeros((2, 4)))     # K_total (:199-201)
    Ki_dead: np.ndarray = field(default_factory=lambda: np.zeros(2))         # K_i (:202)

    def A_variant(self, v: int, scaled: bool = True):
        A = (self.A_s if scaled else self.A).copy()
        if v & 1:
            A[self.sgn_rows, self.sgn_c1_col] *= -1
        if v & 2:
            A[self.sgn_rows, self.sgn_c2_col] *= -1
        return A


def build_problem(sim_conditions, mpc_params, fail_params, debris=None,
                  settings: Optional[SolverSettings] = None, ukf_interval_scale: Optional[float] = None) -> Problem:
    """Build the constant tables.  ``ukf_interval_scale``: the reference's continuous
    simulator scales the disturbance process noise by ``T*int(T/T_cont)``
    (``trajectorySimulateC.py:310``) instead of ``T`` (``trajectorySimulate.py:272``)."""
    st = settings or SolverSettings()
    if st.estimator not in ("ukf", "kf"):
        raise ValueError(f"unknown estimator {st.estimator!r} (\"ukf\" or \"kf\")")
    est = 1 if st.estimator == "kf" else 0
    sc, mp, fp = sim_conditions, mpc_params, fail_params
    Nx, Nc, Nb = int(mp.Nx), int(mp.Nc), int(mp.Nb)
    nx, nu, ny, nd = 4, 2, 5, 2
    T, nmm = float(sc.time_stp), float(sc.mean_mtn)
    x0 = np.asarray(sc.x0, float)
    xr = np.asarray(sc.xr, float)
    Ad, Bd = cw_discretise(nmm, T, bool(sc.isDeltaV))

    Ao = np.eye(6)
    Ao[:4, :4] = Ad
    Ao[0, 4] = Ao[1, 5] = 1.0
    Bou = np.vstack([Bd, np.zeros((2, 2))])

    gam, phi = float(sc.los_ang), float(sc.hatch_ofst)
    den = (sc.r_p - sc.r_tol) * np.sin(gam)
    C = np.array([[np.sin(phi + gam) / den, -np.cos(phi + gam) / den, 0, 0],
                  [-np.sin(phi - gam) / den, np.cos(phi - gam) / den, 0, 0],
                  [0, 1, 0, 0] if sc.inTrack else [1, 0, 0, 0],
                  [0, 0, 1, 1],
                  [0, 1, 0, 0]], float)

    Q, Ru, Rs = _dense(mp.Q_state), _dense(mp.R_input), _dense(mp.R_slack)
    Vecr = np.asarray(mp.V_ecr, float)
    K, S = _lqr_gain(Ad, Bd, Q, Ru)
    if not np.all(np.linalg.eigvals(S) > 0):
        raise Exception("Riccati solution not positive definite")   # ref :206-207
    Acl = Ad - Bd @ K

    Cint = np.atleast_2d(np.asarray(fp.C_int, float))
    Aaug = np.block([[Ad, np.zeros((4, Cint.shape[0]))], [Cint, np.eye(Cint.shape[0])]])
    Baug = np.vstack([Bd, np.zeros((Cint.shape[0], 2))])
    Kf, _ = _lqr_gain(Aaug, Baug, np.asarray(fp.Q_fail, float), np.asarray(fp.R_fail, float))
    if Cint.shape[0] != 1 or not np.array_equal(Cint, np.eye(1, 4)):
        raise NotImplementedError("only the reference's C_int = [1 0 0 0] integrator is supported")

    # ---- decision vector: [x_0..x_Nx | (u_k, s_k) k<Nc | d]
    nX, nU = nx * (Nx + 1), (nu + ny) * Nc
    n = nX + nU + nd
    m = nX + ny * (Nx + 1) + nU + nd
    P = np.zeros((n, n))
    q = np.zeros(n)
    for k in range(Nx + 1):
        Qk = Q if k < Nx else S
        P[4 * k:4 * k + 4, 4 * k:4 * k + 4] = Qk
        q[4 * k:4 * k + 4] = -Qk @ xr
    for k in range(Nc):
        o = nX + 7 * k
        P[o:o + 2, o:o + 2] = Ru
        P[o + 2:o + 7, o + 2:o + 7] = Rs
    P[n - 2:, n - 2:] = np.eye(2)
    # OSQP's check_termination also tests a DUAL-infeasibility certificate (osqp auxil.c is_dual_infeasible: q'dx < 0,
    # ||P dx||_inf <= eps_dual_inf ||dx||_inf, A dx in the recession cone of [l, u]).  The device kernels do not carry it:
    # P is positive definite for this problem family (Q, S, Ru, Rs, I_2 on the diagonal blocks), so
    # ||P dx||_inf >= lam_min(P) ||dx||_inf / sqrt(n) and the certificate cannot hold while lam_min / sqrt(n) > eps_dual_inf.
    # The premise is checked here rather than assumed (a semidefinite P must not silently run without the test).
    lam_min_P = float(np.linalg.eigvalsh(0.5 * (P + P.T)).min())
    if lam_min_P / np.sqrt(n) <= st.eps_dual_inf:
        raise NotImplementedError(f"cost Hessian is not safely positive definite (lam_min = {lam_min_P:.3g}): the engine omits "
                                  "OSQP's dual-infeasibility certificate, which this problem could trigger")

    A = np.zeros((m, n))
    l = np.full(m, -np.inf)
    u = np.full(m, np.inf)
    # (1) dynamics rows
    A[:nX, :nX] = -np.eye(nX)
    for k in range(1, Nx + 1):
        r = 4 * k
        A[r:r + 4, r - 4:r] += Ad if k <= Nc else Acl
        if k <= Nc:
            A[r:r + 4, nX + 7 * (k - 1):nX + 7 * (k - 1) + 2] = Bd
        A[r, n - 2] = 1.0
        A[r + 1, n - 1] = 1.0
    l[:nX] = u[:nX] = 0.0
    l[:4] = u[:4] = -x0
    # (2) LOS / velocity / debris rows
    sgn_rows, c1c, c2c, row3 = [], [], [], []
    val0 = abs(x0[0] - xr[0]) + abs(x0[1] - xr[1])
    for k in range(Nx + 1):
        r = nX + 5 * k
        A[r:r + 5, 4 * k:4 * k + 4] = C
        if k < Nc:
            for j in range(5):
                A[r + j, nX + 7 * k + 2 + j] = Vecr[j]
        sgn_rows.append(r + 3)
        c1c.append(4 * k + 2)
        c2c.append(4 * k + 3)
        if k <= Nb:
            l[r:r + 5] = [1.0, 1.0, sc.r_p, 0.0, -np.inf]
            u[r + 3] = val0
            row3.append(r + 3)
    # (3) input / slack box
    r0 = nX + 5 * (Nx + 1)
    A[r0:r0 + nU, nX:nX + nU] = np.eye(nU)
    for k in range(Nc):
        l[r0 + 7 * k:r0 + 7 * k + 7] = [-mp.u_lim[0], -mp.u_lim[1], 0, 0, 0, 0, 0]
        u[r0 + 7 * k:r0 + 7 * k + 2] = [mp.u_lim[0], mp.u_lim[1]]
    # (4) disturbance pin
    A[m - 2, n - 2] = A[m - 1, n - 1] = 1.0
    l[m - 2:] = u[m - 2:] = 0.0

    # deadbeat avoidance law on the (y, ydot, int y) subsystem, all poles at 0 (ref :190-203)
    Ap_ = Ad[[1, 3], :][:, [1, 3]]
    Bp_ = Bd[[1, 3], 1].reshape(2, 1)
    Aaug_d = np.block([[Ap_, np.zeros((2, 1))], [np.array([[1.0, 0.0]]), np.eye(1)]])
    Baug_d = np.vstack([Bp_, np.zeros((1, 1))])
    Kp_ = _ackermann(Aaug_d, Baug_d, np.zeros(3))
    K_dead = np.zeros((2, 4))
    K_dead[1, 1], K_dead[1, 3] = Kp_[0, 0], Kp_[0, 1]
    Ki_dead = np.array([0.0, Kp_[0, 2]])
    if debris is not None:
        # The half-plane row C[4,:] = [-slope, 1, 0, 0] is rebuilt from the estimate every step and OSQP then
        # re-scales and refactors (ref :345-348): no table can be shared, the device does the whole setup per lane.
        verts = np.asarray(debris.constructVertArr(), float)
        if sc.inTrack:
            verts = verts[[1, 2, 3, 0]]                    # simhelpers.py:52-53
        return Problem(Nx=Nx, Nc=Nc, Nb=Nb, n=n, m=m, in_track=bool(sc.inTrack), delta_v=bool(sc.isDeltaV),
                       is_reject=bool(sc.isReject), has_noise=sc.noise is not None, T=T, mean_mtn=nmm,
                       r_p=float(sc.r_p), r_tol=float(sc.r_tol), xr=xr, umax0=float(mp.u_lim[0]),
                       suc_dist=float(sc.suc_cond[0]), suc_ang=float(sc.suc_cond[1]),
                       sig=np.array(sc.noise.noise_std, float) if sc.noise is not None else np.zeros(2),
                       noise_length=int(sc.noise.noise_length) if sc.noise is not None else 1,
                       Ad=Ad, Bd=Bd, Ao=Ao, Bou=Bou, Qw=_ukf_process_noise(T if ukf_interval_scale is None else ukf_interval_scale, sc),
                       Kpf=Kf[:, :4], Kif=Kf[:, 4:5], C=C, P=P, q=q, A=A, l=l, u=u,
                       D=np.ones(n), E=np.ones(m), c=1.0, P_s=None, q_s=None, A_s=None, l_s=None, u_s=None, ctype=None,
                       sgn_rows=np.array(sgn_rows), sgn_c1_col=np.array(c1c), sgn_c2_col=np.array(c2c), row3=np.array(row3),
                       V=None, lam=None, settings=st, u_off=nX, has_debris=True,
                       debris_center=np.asarray(debris.center, float), debris_side=float(debris.side_length),
                       debris_detect=float(debris.detect_distance), debris_verts=verts, K_dead=K_dead, Ki_dead=Ki_dead, estimator=est)
    P_s, q_s, A_s, D, E, c = ruiz_equilibrate(P, q, A, st.scaling) if st.scaling else (P, q, A, np.ones(n), np.ones(m), 1.0)
    l_s = E * np.maximum(l, -OSQP_INFTY)
    u_s = E * np.minimum(u, OSQP_INFTY)
    free = (l_s < -OSQP_INFTY * MIN_SCALING) & (u_s > OSQP_INFTY * MIN_SCALING)
    ctype = np.where(free, -1, np.where(u_s - l_s < RHO_TOL, 1, 0)).astype(np.int32)
    ctype[np.array(row3)] = 0       # dynamic rows; the engine flags lanes where E*val < RHO_TOL

    prob = Problem(Nx=Nx, Nc=Nc, Nb=Nb, n=n, m=m, in_track=bool(sc.inTrack), delta_v=bool(sc.isDeltaV),
                   is_reject=bool(sc.isReject), has_noise=sc.noise is not None, T=T, mean_mtn=nmm,
                   r_p=float(sc.r_p), r_tol=float(sc.r_tol), xr=xr, umax0=float(mp.u_lim[0]),
                   suc_dist=float(sc.suc_cond[0]), suc_ang=float(sc.suc_cond[1]),
                   sig=np.array(sc.noise.noise_std, float) if sc.noise is not None else np.zeros(2),
                   noise_length=int(sc.noise.noise_length) if sc.noise is not None else 1,
                   Ad=Ad, Bd=Bd, Ao=Ao, Bou=Bou, Qw=np.zeros((6, 6)), Kpf=Kf[:, :4], Kif=Kf[:, 4:5], C=C,
                   P=P, q=q, A=A, l=l, u=u, D=D, E=E, c=c, P_s=P_s, q_s=q_s, A_s=A_s, l_s=l_s, u_s=u_s, ctype=ctype,
                   sgn_rows=np.array(sgn_rows), sgn_c1_col=np.array(c1c), sgn_c2_col=np.array(c2c),
                   row3=np.array(row3), V=np.zeros((4, n, n)), lam=np.zeros((4, n)), settings=st, u_off=nX,
                   K_dead=K_dead, Ki_dead=Ki_dead, estimator=est)
    prob.Qw = _ukf_process_noise(T if ukf_interval_scale is None else ukf_interval_scale, sc)

    # ---- spectral operator per sign variant
    w = np.where(ctype == 1, RHO_EQ_FACTOR, 1.0)
    for v in range(4):
        Av = prob.A_variant(v)
        Af, Ac = Av[ctype == -1], Av[ctype != -1]
        Bm = P_s + st.sigma * np.eye(n) + RHO_MIN * Af.T @ Af
        Gm = Ac.T @ (w[ctype != -1][:, None] * Ac)
        lam, Vv = sla.eigh(0.5 * (Gm + Gm.T), 0.5 * (Bm + Bm.T))
        prob.V[v], prob.lam[v] = Vv, np.maximum(lam, 0.0)
    return prob
