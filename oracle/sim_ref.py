"""Restatement of the reference's simulators and QP assembly (ORACLE -- test infrastructure).

Follows, block for block:
  ``/root/reference/src/trajectorySimulate.py:17-388``   -> :func:`build_setup`, :func:`trajectory_simulate`
  ``/root/reference/src/trajectorySimulateC.py:17-446``  -> :func:`trajectory_simulate_c`
  ``/root/reference/src/simhelpers.py:11-140``           -> :func:`configure_dynamic_constraints`
  ``/root/reference/src/simhelpers.py:142-172``          -> :func:`construct_osqp_aeq`
  ``/root/reference/src/simhelpers.py:174-189``          -> :func:`continuous_append_index`

Third-party numerics are the restatements in ``oracle.osqp_ref`` / ``oracle.ukf_ref`` /
``oracle.control_ref``.  Matrices are dense numpy (the reference uses scipy.sparse; same
values, see ``tests/golden/assembly_*.npz`` which were captured from the reference's own
code).  Parameter objects are duck-typed (``src.mpcsim`` attribute names).

Extra, oracle-only knobs (all default to the reference's behaviour):
  * ``draw``: callable returning 4 standard normals (reference: numpy legacy global RNG,
    re-seeded with 123 on every discrete call, ``trajectorySimulate.py:28,268,352``);
  * ``solver_settings``: dict merged into the OSQP settings;
  * ``integrator``: 'rk45' (= ``solve_ivp`` defaults, reference) or 'rk4' (fixed step h=T_cont);
  * ``chol_fail``: 'raise' (reference: scipy's Cholesky raises LinAlgError when the UKF covariance,
    singular because R = 0, rounds to a non-positive pivot -- about 2 % of randomly placed lanes) or
    'clamp' (the engine's continuation on the semi-definite factor, ``oracle/ukf_ref.py``).
"""
import math
from types import SimpleNamespace

import numpy as np
import scipy as sp
import scipy.linalg
import scipy.integrate

from .control_ref import dlqr_integral, acker, white_noise
from .osqp_ref import OSQPRef
from .ukf_ref import UKFRef, MerweScaledSigmaPointsRef
from .kf_ref import LinearKFRef


def _dense(M):
    return np.array(M.toarray() if hasattr(M, "toarray") else M, float)


# --------------------------------------------------------------------------- helpers
def construct_osqp_aeq(Nx, Nc, Ad, Bd, K, ny):
    """simhelpers.py:142-172 -- dynamics equality block [Ax | Bu]."""
    nx, nu = Bd.shape
    Ax1 = np.kron(np.eye(Nc + 1), -np.eye(nx)) + np.kron(np.eye(Nc + 1, k=-1), Ad)
    Acl = Ad - Bd @ K
    Ax2 = np.kron(np.eye(Nx - Nc), -np.eye(nx)) + np.kron(np.eye(Nx - Nc, k=-1), Acl)
    Ax = sp.linalg.block_diag(Ax1, Ax2)
    Ax4 = np.zeros((Nx + 1, Nx + 1))
    Ax4[Nc + 1, Nc] = 1
    Ax = Ax + np.kron(Ax4, Acl)
    BuI = np.vstack([np.zeros((1, Nc)), np.eye(Nc), np.zeros((Nx - Nc, Nc))])
    Bdaug = np.hstack([Bd, np.zeros((nx, ny))])
    return np.hstack([Ax, np.kron(BuI, Bdaug)])


def configure_dynamic_constraints(sc, mp, debris, xest, block_mats, u_lim):
    """simhelpers.py:11-140.  Mutates ``block_mats.C`` and (in-track) swaps
    ``xest[0], xest[1]`` IN PLACE on the caller's array, like the reference."""
    rp = sc.r_p
    xr = sc.xr
    isReject = sc.isReject
    rx, ry = xr[0], xr[1]
    Nx, Nc, Nb = mp.Nx, mp.Nc, mp.Nb
    Aeq, Aineq2, Block12, Block21, AextRow, AextCol, C = block_mats
    ny = C.shape[0]
    umin, umax = u_lim

    if debris is not None:
        sqVerts = debris.constructVertArr()
        if sc.inTrack:
            sqVerts = np.vstack([sqVerts[1], sqVerts[2], sqVerts[3], sqVerts[0]])
        center = debris.center
        sideLength = debris.side_length
        hasDebris = True
        detect_dist = debris.detect_distance
    else:
        center = (-np.inf, -np.inf)
        sideLength = 0
        hasDebris = False
        detect_dist = np.inf

    C1 = 1 if xest[2] >= 0 else -1      # (-1, 1)[xest[2] >= 0], simhelpers.py:66
    C2 = 1 if xest[3] >= 0 else -1

    if sc.inTrack:
        xestCalc = np.copy(xest)
        xest[0], xest[1] = xest[1], xest[0]
        center = [center[1], center[0]]
    else:
        xestCalc = xest

    inside = (xest[0] - (center[0] + sideLength / 2) < 0 and xest[0] - (center[0] - sideLength / 2) > 0)
    if xest[1] >= 0:
        v = 1 if inside else 0
    else:
        v = 2 if inside else 3
    inter = None
    if inside or hasDebris:
        slope = (xestCalc[1] - sqVerts[v, 1]) / (xestCalc[0] - sqVerts[v, 0])
        inter = -slope * xestCalc[0] + xestCalc[1]
    else:
        slope = 0

    C[3, 2] = C1
    C[3, 3] = C2
    C[4, 0] = -slope
    Aineq1 = np.kron(np.eye(Nx + 1), C)
    Aineq = np.block([[Aineq1, Block12], [Block21, Aineq2]])
    A = np.vstack([Aeq, Aineq])
    A = np.hstack([A, AextCol])
    A = np.vstack([A, AextRow])

    near = (xest[0] - (center[0] + sideLength / 2) < detect_dist and xest[0] - (center[0] + sideLength / 2) > 0)
    val = np.absolute(xestCalc[0] - rx) + np.absolute(xestCalc[1] - ry)
    if xest[1] >= 0:
        xmin = np.array([1., 1., rp, 0., inter if (inside or near) else -np.inf])
        xmax = np.array([np.inf, np.inf, np.inf, val, np.inf])
    else:
        xmax = np.array([np.inf, np.inf, np.inf, val, inter if (inside or near) else np.inf])
        xmin = np.array([1., 1., rp, 0., -np.inf])

    dpin = (1.0 if isReject else 0.0) * np.asarray(xest[4:6], float)
    lineq = np.hstack([np.kron(np.ones(Nb + 1), xmin), np.kron(np.ones(Nx - Nb), -np.inf * np.ones(ny)),
                       np.kron(np.ones(Nc), umin), dpin])
    uineq = np.hstack([np.kron(np.ones(Nb + 1), xmax), np.kron(np.ones(Nx - Nb), np.inf * np.ones(ny)),
                       np.kron(np.ones(Nc), umax), dpin])
    return A, lineq, uineq


def continuous_append_index(impc, ifailsf, ifailsd, i):
    """simhelpers.py:174-189."""
    if impc and impc[-1] == i - 1:
        impc.append(i)
    elif ifailsf and ifailsf[-1] == i - 1:
        ifailsf.append(i)
    elif ifailsd and ifailsd[-1] == i - 1:
        ifailsd.append(i)


def discretise(n, T, isDeltaV, use_sympy=False):
    """trajectorySimulate.py:73-111.  ``use_sympy`` follows the reference literally
    (sympy matrix exponential + 16 scalar ``quad`` calls); the default uses the Van-Loan
    block exponential, equal to it to ~1e-12."""
    Ap = np.array([[0., 0., 1., 0.], [0., 0., 0., 1.], [3 * n ** 2, 0., 0., 2 * n], [0., 0., -2 * n, 0.]])
    Bp = np.array([[0., 0.], [0., 0.], [1., 0.], [0., 1.]])
    Ad = sp.linalg.expm(Ap * T)
    if isDeltaV:
        Bd = Ad @ np.vstack([np.zeros([2, 2]), np.eye(2)])
    elif use_sympy:
        import sympy as sy
        x = sy.symbols('x')
        eAx = (sy.Matrix(Ap) * x).exp()
        eAxInt = np.empty([4, 4])
        for (i, j), f in np.ndenumerate(eAx):
            eAxInt[i, j] = sp.integrate.quad(sy.lambdify((x), f), 0., T)[0]
        Bd = eAxInt @ Bp
    else:
        M = np.zeros((6, 6))
        M[:4, :4] = Ap
        M[:4, 4:] = Bp
        Bd = sp.linalg.expm(M * T)[:4, 4:]
    return Ap, Bp, Ad, Bd


def build_setup(sc, mp, fp, debris, use_sympy=False):
    """Everything ``trajectorySimulate.py:30-245`` computes before the closed loop
    (identical in ``trajectorySimulateC.py:30-272``), returned as one namespace."""
    s = SimpleNamespace()
    noise = sc.noise
    if noise is not None:
        s.sigMat = noise.constructSigMat()
        s.noiseRepeat = noise.noise_length
    else:
        s.sigMat = np.diag([0., 0., 0., 0.])
        s.noiseRepeat = 1
    gam, rp, rtot, phi, n, T = sc.los_ang, sc.r_p, sc.r_tol, sc.hatch_ofst, sc.mean_mtn, sc.time_stp
    x0 = np.asarray(sc.x0, float)
    xr = np.asarray(sc.xr, float)
    if debris is not None:
        sqVerts = debris.constructVertArr()
        center, sideLength, hasDebris = debris.center, debris.side_length, True
    else:
        center, sideLength, hasDebris = (-np.inf, -np.inf), 0, False
    s.center, s.sideLength, s.hasDebris = center, sideLength, hasDebris

    Ap, Bp, Ad, Bd = discretise(n, T, sc.isDeltaV, use_sympy)
    nx, nu = Bp.shape
    ndi = 2
    Cm = np.array([[1., 0., 0., 0.], [0., 1., 0., 0.]])
    nym = 2
    Ao = sp.linalg.block_diag(Ad, np.eye(ndi))
    Ao[0, 4] = 1.
    Ao[1, 5] = 1.
    Bou = np.vstack([Bd, np.zeros([2, 2])])

    den = (rp - rtot) * math.sin(gam)
    C_11 = math.sin(phi + gam) / den
    C_12 = -math.cos(phi + gam) / den
    C_21 = -math.sin(phi - gam) / den
    C_22 = math.cos(phi - gam) / den
    if (x0[0] - (center[0] + sideLength / 2) < 0 and x0[0] - (center[0] - sideLength / 2) > 0):
        slope = (x0[1] - sqVerts[1, 1]) / (x0[0] - sqVerts[1, 0])
    elif hasDebris:
        slope = (x0[1] - sqVerts[0, 1]) / (x0[0] - sqVerts[0, 0])
    else:
        slope = 0
    C = np.array([[C_11, C_12, 0., 0.], [C_21, C_22, 0., 0.], [1., 0., 0., 0.], [0., 0., 1., 1.], [-slope, 1., 0., 0.]])
    if sc.inTrack:
        C[2, :] = np.array([0., 1., 0., 0.])
    ny = C.shape[0]

    ulim = mp.u_lim
    umin = np.hstack([-ulim[0], -ulim[1], np.zeros(ny)])
    umax = np.hstack([ulim[0], ulim[1], np.inf * np.ones(ny)])
    Vecr = np.asarray(mp.V_ecr, float)
    D = np.hstack([np.zeros([ny, nu]), np.diag(Vecr)])

    Q = _dense(mp.Q_state)
    Ru = _dense(mp.R_input)
    Rs = _dense(mp.R_slack)
    R = sp.linalg.block_diag(Ru, Rs)
    S = sp.linalg.solve_discrete_are(Ad, Bd, Q, Ru)
    K = np.linalg.inv(Ru + Bd.T @ S @ Bd) @ (Bd.T @ S @ Ad)
    QN = S

    Crefx = np.atleast_2d(np.asarray(fp.C_int, float))
    nr = Crefx.shape[0]
    Kf = dlqr_integral(Ad, Bd, fp.Q_fail, fp.R_fail, Crefx)
    Kpf = Kf[:, :nx]
    Kif = Kf[:, nx:nx + nr]

    Crefy = np.array([[0., 1., 0., 0.]])
    Bd_prune = Bd[:, 1].reshape(nx, 1)[[1, 3], ]
    Ad_prune = Ad[[1, 3], :][:, [1, 3]]
    C_prune = np.array([1, 0])
    A_aug = np.block([[Ad_prune, np.zeros([2, 1])], [C_prune, np.eye(1)]])
    B_aug = np.block([[Bd_prune], [np.zeros([1, 1])]])
    K_prune = acker(A_aug, B_aug, np.array([0, 0, 0]))
    K_total = np.zeros([nu, nx])
    K_total[1, 1] = K_prune[0, 0]
    K_total[1, 3] = K_prune[0, 1]
    K_i = np.vstack([0, K_prune[0, 2]])

    if not np.all(np.linalg.eigvals(S) > 0):
        raise Exception("Riccati solution not positive definite")

    Nx, Nc, Nb = mp.Nx, mp.Nc, mp.Nb
    P = sp.linalg.block_diag(np.kron(np.eye(Nx), Q), QN, np.kron(np.eye(Nc), R), np.eye(ndi))
    q = np.hstack([np.kron(np.ones(Nx), -Q @ xr), -QN @ xr, np.zeros(Nc * (nu + ny)), np.zeros(ndi)])
    Aeq = construct_osqp_aeq(Nx, Nc, Ad, Bd, K, ny)
    leq = np.hstack([-x0, np.zeros(Nx * nx)])
    ueq = leq

    Aineq2 = np.eye(Nc * (nu + ny))
    Block12 = np.vstack([np.kron(np.eye(Nc), D), np.zeros(((Nx + 1 - Nc) * ny, Nc * (nu + ny)))])
    Block21 = np.zeros((Nc * (nu + ny), (Nx + 1) * nx))
    AextCol = np.vstack([np.zeros([nx, ndi]),
                         np.kron(np.ones([Nx, 1]), np.vstack([np.eye(ndi), np.zeros([nx - ndi, ndi])])),
                         np.zeros([(Nx + 1) * ny, ndi]), np.zeros([Nc * (nu + ny), ndi])])
    AextRow = np.hstack([np.zeros([ndi, (Nx + 1) * nx]), np.zeros([ndi, Nc * (nu + ny)]), np.eye(ndi)])
    block_mats = (Aeq, Aineq2, Block12, Block21, AextRow, AextCol, C)
    u_lim = (umin, umax)

    A, lineq, uineq = configure_dynamic_constraints(sc, mp, debris, np.hstack([np.copy(x0), 0, 0]), block_mats, u_lim)
    l = np.hstack([leq, lineq])
    u = np.hstack([ueq, uineq])

    s.__dict__.update(dict(Ap=Ap, Bp=Bp, Ad=Ad, Bd=Bd, Ao=Ao, Bou=Bou, Cm=Cm, C=C, umin=umin, umax=umax, D=D,
                           Q=Q, Ru=Ru, Rs=Rs, S=S, K=K, QN=QN, Kpf=Kpf, Kif=Kif, Crefx=Crefx, Crefy=Crefy,
                           K_total=K_total, K_i=K_i, P=P, q=q, Aeq=Aeq, A=A, l=l, u=u, block_mats=block_mats,
                           u_lim=u_lim, nx=nx, nu=nu, ny=ny, ndi=ndi, nym=nym, Nx=Nx, Nc=Nc, Nb=Nb,
                           x0=x0, xr=xr, T=T, n=n, rp=rp, rtot=rtot))
    return s


def _measure(estimator, x):
    """UKF: range / bearing of the true state (trajectorySimulate.py:329-332); linear KF: its position (Cm x)."""
    if estimator == 'kf':
        return np.array([x[0], x[1]])
    return np.array([np.linalg.norm(x[:2]), math.atan2(x[1], x[0])])


def _make_ukf(s, xest0, Bnoise_scale, regen_sigmas, chol_fail='raise', estimator='ukf'):
    """trajectorySimulate.py:121-130, 250, 272-282 (UKF model, P0, Q, R)."""
    Ao, Bou = s.Ao, s.Bou

    def fx(x, u):
        return Ao @ x + Bou @ u

    def hx(x):
        return np.array([np.linalg.norm(x[:2]), math.atan2(x[1], x[0])])

    pts = MerweScaledSigmaPointsRef(6, alpha=0.1, beta=2., kappa=-1, chol_fail=chol_fail)
    Bnoise = np.vstack([np.zeros([s.nx, s.ndi]), Bnoise_scale * np.eye(s.ndi)])
    Qw = np.diag([s.sigMat[0, 0] ** 2, s.sigMat[1, 1] ** 2])
    Qw = Bnoise @ Qw @ Bnoise.T
    Qw[:4, :][:, :4] = 0.001 * np.eye(s.nx)
    kf = UKFRef(6, 2, fx, hx, pts, regen_sigmas=regen_sigmas) if estimator == 'ukf' else LinearKFRef(Ao, Bou)
    kf.x = np.array(xest0, float)
    kf.P = sp.linalg.block_diag(1e-20 * np.eye(s.nx), np.eye(s.ndi))
    kf.R = np.zeros([s.nym, s.nym])
    kf.Q = Qw
    return kf, Qw


def _select_control(s, status, xest4, xintf, res_x):
    """trajectorySimulate.py:299-319: status branch, failsafe / deadbeat laws, norm clip."""
    center, side = s.center, s.sideLength
    if status != 'solved':
        if (xest4[0] - (center[0] + side / 2) < 0 and xest4[0] - (center[0] - side / 2) > 0
                and xest4[1] < (center[1] + side / 2) and xest4[1] > (center[1] - side / 2)):
            which = 3
            xintf = xintf + s.Crefy @ xest4 - (center[1] + side / 2)
            ctrl = -s.K_total @ xest4 - s.K_i @ xintf
        else:
            which = 2
            xintf = xintf + s.Crefx @ xest4 - s.xr[0]
            ctrl = -s.Kpf @ xest4 - s.Kif @ xintf
    else:
        which = 1
        xintf = 0
        ctrl = np.array(res_x[(s.Nx + 1) * s.nx:(s.Nx + 1) * s.nx + s.nu], float)
    ctrl = np.array(ctrl, float).reshape(-1)
    raw = ctrl.copy()
    if np.linalg.norm(ctrl) > s.umax[0]:
        ctrl[0] = ctrl[0] * (s.umax[0] / np.linalg.norm(ctrl))
        ctrl[1] = ctrl[1] * (s.umax[0] / np.linalg.norm(ctrl))
    return ctrl, raw, xintf, which


def _success_scan(xtruePiece, iterm, xr, distTol, angTol):
    """trajectorySimulate.py:369-376."""
    with np.errstate(all='ignore'):
        for i in range(iterm - 1, 0, -1):
            dist = np.linalg.norm(xtruePiece[0:2, i] - xr[0:2])
            ang = abs(math.atan(np.float64(xtruePiece[3, i]) / np.float64(xtruePiece[2, i]))) * (180 / np.pi)
            if dist <= distTol and ang <= angTol:
                return True
    return False


def _terminated(inTrack, x, rp, rtot):
    """trajectorySimulate.py:288-293."""
    if not inTrack:
        return np.linalg.norm(x[0:2]) < rp or x[0] < rp - rtot
    return np.linalg.norm(x[0:2]) < rp or x[1] < rp - rtot


def _legacy_draw():
    return np.random.normal(0, 1, 4)


# --------------------------------------------------------------------------- the real package, if it is ever importable
def real_osqp_available():
    """SURVEY 8(c) / BASELINE.md section 3: osqp is not installable offline, so the oracle restates it.  If ``import osqp``
    succeeds at run time the closed loop below can run on the REAL solver (``solver='osqp'``), which is the only way the
    third-party boundary ever gets pinned; bench.py's reference arm switches to it and says so."""
    try:
        import osqp  # noqa: F401
        return True
    except Exception:
        return False


class _RealOSQP:
    """``osqp.OSQP`` behind the small protocol ``OSQPRef`` exposes (dense ``A`` on update -> the CSC value array the reference
    passes as ``Ax``, trajectorySimulate.py:348)."""

    def setup(self, P, q, A, l, u, **kw):
        import osqp
        from scipy import sparse
        self._A = sparse.csc_matrix(A)
        self._A.sort_indices()
        self._rows, self._cols = self._A.nonzero()
        order = np.lexsort((self._rows, self._cols))
        self._rows, self._cols = self._rows[order], self._cols[order]
        self._prob = osqp.OSQP()
        self._prob.setup(sparse.triu(sparse.csc_matrix(P)).tocsc(), np.asarray(q, float), self._A, np.asarray(l, float),
                         np.asarray(u, float), **kw)

    def solve(self):
        res = self._prob.solve()
        res.info.rho = getattr(res.info, 'rho_estimate', float('nan'))
        return res

    def update(self, l=None, u=None, A=None):
        kw = {}
        if l is not None:
            kw['l'] = np.asarray(l, float)
        if u is not None:
            kw['u'] = np.asarray(u, float)
        if A is not None:
            kw['Ax'] = np.asarray(A)[self._rows, self._cols].astype(float)
        self._prob.update(**kw)


# --------------------------------------------------------------------------- discrete
def trajectory_simulate(sc, mp, fp, debris, draw=None, solver_settings=None, regen_sigmas=True,
                        use_sympy=False, max_steps=None, chol_fail='raise', estimator='ukf', solver='restated'):
    """trajectorySimulate.py:17-388.  Returns a SimRun-like namespace plus per-step solver
    telemetry (``status_val``, ``iters``, ``rho``, ``u_raw``) used by the parity tests."""
    if draw is None:
        np.random.seed(123)                       # :28
        draw = _legacy_draw
    s = build_setup(sc, mp, fp, debris, use_sympy)
    nx, nu, ndi, Nx = s.nx, s.nu, s.ndi, s.Nx
    Ad, Bd = s.Ad, s.Bd
    noise = sc.noise
    nsim = int(sc.T_final / sc.time_stp)
    if max_steps is not None:
        nsim = min(nsim, max_steps)
    l, u = s.l.copy(), s.u.copy()

    prob = _RealOSQP() if solver == 'osqp' else OSQPRef()
    prob.setup(s.P, s.q, s.A, l, u, **dict(dict(warm_start=True, verbose=False), **(solver_settings or {})))

    xest0 = np.hstack([s.x0, 0., 0.])
    iterm = nsim
    ifailsd, ifailsf, impc = [], [], []
    xtrueP = np.full([nx, nsim + 1], np.nan)
    xestO = np.full([nx + ndi, nsim + 1], np.nan)
    xintf = 0
    noiseStored = np.full([nx, nsim + 1], np.nan)
    ctrls = np.full([nu, nsim + 1], np.nan)
    ctrls[:, 0] = 0.
    xtrueP[:, 0] = s.x0
    xestO[:, 0] = xest0
    noiseVec = s.sigMat @ draw()
    noiseStored[:, 0] = noiseVec
    kf, _ = _make_ukf(s, xest0, s.T, regen_sigmas, chol_fail, estimator)

    status_val = np.zeros(nsim, int)
    iters = np.zeros(nsim, int)
    rhos = np.zeros(nsim)
    u_raw = np.full([nu, nsim], np.nan)

    for i in range(nsim):
        if _terminated(sc.inTrack, xtrueP[:, i], s.rp, s.rtot):
            iterm = i
            break
        res = prob.solve()
        status_val[i], iters[i], rhos[i] = res.info.status_val, res.info.iter, res.info.rho
        ctrl, raw, xintf, which = _select_control(s, res.info.status, xestO[:4, i], xintf, res.x)
        u_raw[:, i] = raw
        (impc if which == 1 else ifailsf if which == 2 else ifailsd).append(i)

        ctrls[:, i + 1] = ctrl
        xtrueP[:, i + 1] = Ad @ xtrueP[:, i] + Bd @ ctrls[:, i] + noiseVec

        if noise is not None:
            ymeas = _measure(estimator, xtrueP[:, i + 1])
            kf.predict(ctrls[:, i])
            kf.update(ymeas)
            xestO[:, i + 1] = kf.x
        else:
            xestO[:, i + 1] = np.hstack([xtrueP[:, i + 1], [0., 0.]])

        l[:nx] = -xestO[:4, i + 1]
        u[:nx] = -xestO[:4, i + 1]
        prob.update(l=l, u=u)
        A, lineq, uineq = configure_dynamic_constraints(sc, mp, debris, xestO[:, i + 1], s.block_mats, s.u_lim)
        l[(Nx + 1) * nx:] = lineq
        u[(Nx + 1) * nx:] = uineq
        prob.update(A=A, l=l, u=u)

        if (i + 1) % s.noiseRepeat == 0:
            noiseVec = s.sigMat @ draw()
        noiseStored[:, i + 1] = noiseVec

    xtruePiece = np.full([nx, iterm], np.nan)
    for idx in (impc, ifailsf, ifailsd):
        xtruePiece[:, idx] = xtrueP[:, idx]
    succ = _success_scan(xtruePiece, iterm, s.xr, sc.suc_cond[0], sc.suc_cond[1])
    seq = np.full(iterm, np.nan)
    seq[impc] = 1
    seq[ifailsf] = 2
    seq[ifailsd] = 3
    return SimpleNamespace(i_term=iterm, isSuccess=succ, x_true_pcw=xtruePiece, x_est=xestO, ctrl_hist=ctrls,
                           ctrlr_seq=seq, noise_hist=noiseStored, x_true=xtrueP, status_val=status_val[:iterm],
                           iters=iters[:iterm], rho=rhos[:iterm], u_raw=u_raw[:, :iterm], setup=s,
                           ukf_clamped=kf.points.clamped)


# --------------------------------------------------------------------------- continuous
def state_eqn_n(x, u, n):
    """trajectorySimulateC.py:64-79 (nonlinear planar relative motion, 500 km orbit)."""
    R_T = 500e+03 + 6378.1e+03
    mu = (n ** 2) * (R_T ** 3)
    r3 = ((R_T + x[0]) ** 2 + x[1] ** 2) ** (3 / 2)
    return np.array([x[2], x[3],
                     2 * n * x[3] + (n ** 2) * x[0] - (mu * (R_T + x[0])) / r3 + mu / (R_T ** 2) + u[0],
                     -2 * n * x[2] + (n ** 2) * x[1] - (mu * x[1]) / r3 + u[1]])


def _substep(x, u, n, t, h, integrator):
    if integrator == 'rk45':
        sol = sp.integrate.solve_ivp(lambda tt, xx: state_eqn_n(xx, u, n), (t, t + h), x)
        return sol.y[:, -1]
    k1 = state_eqn_n(x, u, n)
    k2 = state_eqn_n(x + 0.5 * h * k1, u, n)
    k3 = state_eqn_n(x + 0.5 * h * k2, u, n)
    k4 = state_eqn_n(x + h * k3, u, n)
    return x + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)


def trajectory_simulate_c(sc, mp, fp, debris, V=None, solver_settings=None, regen_sigmas=True,
                          integrator='rk45', use_sympy=False, chol_fail='raise', estimator='ukf'):
    """trajectorySimulateC.py:17-446.  ``V`` (2 x n_refresh) replaces the ``ct.white_noise``
    draw (:301) when given.  The loop's literal start index 500 (:325) is restated as
    ``int(T/T_cont)``, which is what it equals for every shipped parameter set."""
    s = build_setup(sc, mp, fp, debris, use_sympy)
    nx, nu, ndi, Nx = s.nx, s.nu, s.ndi, s.Nx
    T, T_cont, time_final = sc.time_stp, sc.T_cont, sc.T_final
    noise = sc.noise
    nsimD = int(time_final / T)
    nsimC = int(time_final / T_cont)
    ratio = int(T / T_cont)
    xTimeD = np.arange(0, time_final, T)
    xTimeC = np.arange(0, time_final, T_cont)
    l, u = s.l.copy(), s.u.copy()

    prob = OSQPRef()
    prob.setup(s.P, s.q, s.A, l, u, **dict(dict(warm_start=True, verbose=False), **(solver_settings or {})))

    xest0 = np.hstack([s.x0, 0., 0.])
    iterm = nsimC
    ifailsd, ifailsf, impc = [], [], []
    xtrueP = np.full([nx, nsimC], np.nan)
    xestO = np.full([nx + ndi, nsimD + 1], np.nan)
    xintf = 0
    noiseStored = np.full([nx, nsimC], np.nan)
    ctrls = np.full([nu, nsimC], np.nan)
    ctrls[:, :ratio + 1] = 0.
    xtrueP[:, :ratio + 1] = s.x0.reshape(-1, 1)
    xestO[:, 0] = xest0

    Qcont = np.diag([s.sigMat[0, 0] ** 2, s.sigMat[0, 0] ** 2])          # sigma_x twice, :296
    noiseRepeat = s.noiseRepeat
    noiseTimes = np.arange(0, time_final, T * noiseRepeat)
    noiseIntC = int((noiseRepeat * T) / T_cont)
    if V is None:
        V = white_noise(noiseTimes, Qcont, lambda size: np.random.normal(0, 1, size))
    V = np.asarray(V, float)
    sum_vec = np.full([nx, nsimD], np.nan)
    for j, col in enumerate(V.T):
        noiseStored[:, j * noiseIntC:noiseIntC * (1 + j)] = np.vstack([col.reshape(ndi, 1), np.zeros([2, 1])])
        sum_vec[:, j * noiseRepeat:noiseRepeat * (1 + j)] = ratio * np.concatenate([col, np.zeros(2)]).reshape(-1, 1)
    kf, _ = _make_ukf(s, xest0, T * ratio, regen_sigmas, chol_fail, estimator)

    status_val, iters, u_raw, solve_at = [], [], [], []
    disc_j = 1
    time = T
    i = ratio - 1
    for i in range(ratio, nsimC - 1):
        if _terminated(sc.inTrack, xtrueP[:, i], s.rp, s.rtot):
            iterm = i
            break
        sample = (disc_j < nsimD) and (xTimeC[i] == xTimeD[disc_j])
        if sample:
            res = prob.solve()
            ctrl, raw, xintf, which = _select_control(s, res.info.status, xestO[:4, disc_j - 1], xintf, res.x)
            (impc if which == 1 else ifailsf if which == 2 else ifailsd).append(i)
            status_val.append(res.info.status_val)
            iters.append(res.info.iter)
            u_raw.append(raw)
            solve_at.append(i)
            ctrls[:, i + 1] = ctrl
        else:
            continuous_append_index(impc, ifailsf, ifailsd, i)
            ctrls[:, i + 1] = ctrls[:, i]

        if not sc.isDeltaV:
            xtrueP[:, i + 1] = _substep(xtrueP[:, i], ctrls[:, i], s.n, time, T_cont, integrator) + noiseStored[:, i]
        else:
            xn = _substep(xtrueP[:, i], np.zeros(nu), s.n, time, T_cont, integrator)
            if sample:
                xtrueP[:, i + 1] = xn + np.hstack([np.zeros(2), ctrls[:, i]]) + noiseStored[:, i]
            else:
                xtrueP[:, i + 1] = xn + noiseStored[:, i]

        if sample:
            if noise is not None:
                ymeas = _measure(estimator, xtrueP[:, i + 1])
                kf.predict(ctrls[:, i])
                kf.update(ymeas)
                xestO[:, disc_j] = kf.x
            else:
                xestO[:, disc_j] = np.hstack([xtrueP[:, i + 1], [0., 0.]])
            l[:nx] = -xestO[:4, disc_j]
            u[:nx] = -xestO[:4, disc_j]
            prob.update(l=l, u=u)
            A, lineq, uineq = configure_dynamic_constraints(sc, mp, debris, xestO[:, disc_j], s.block_mats, s.u_lim)
            l[(Nx + 1) * nx:] = lineq
            u[(Nx + 1) * nx:] = uineq
            prob.update(A=A, l=l, u=u)
            disc_j += 1
        time = time + T_cont
    continuous_append_index(impc, ifailsf, ifailsd, i + 1)

    xtruePiece = np.full([nx, iterm], np.nan)
    xtruePiece[:, :ratio + 1] = xtrueP[:, :ratio + 1]
    for idx in (impc, ifailsf, ifailsd):
        idx = [k for k in idx if k < iterm]
        xtruePiece[:, idx] = xtrueP[:, idx]
    succ = _success_scan(xtruePiece, iterm, s.xr, sc.suc_cond[0], sc.suc_cond[1])
    seq = np.full(iterm, np.nan)
    seq[:ratio] = 0
    for code, idx in ((1, impc), (2, ifailsf), (3, ifailsd)):
        seq[[k for k in idx if k < iterm]] = code
    seq[-1] = seq[-2]
    return SimpleNamespace(i_term=iterm, isSuccess=succ, x_true_pcw=xtruePiece, x_est=xestO, ctrl_hist=ctrls,
                           ctrlr_seq=seq, noise_hist=sum_vec, x_true=xtrueP, status_val=np.array(status_val),
                           iters=np.array(iters), u_raw=np.array(u_raw).T if u_raw else np.zeros((2, 0)),
                           solve_at=np.array(solve_at), n_est=disc_j, setup=s, ukf_clamped=kf.points.clamped)
