"""Restatement of filterpy's UKF *as the reference uses it* (ORACLE -- test infrastructure).

[3P: ``filterpy.kalman.UnscentedKalmanFilter`` + ``MerweScaledSigmaPoints``, filterpy
1.4.5 (final release), unpinned; source not in /root/reference.  PARITY UNPINNED.]

Reference call sites: ``src/trajectorySimulate.py:121-130`` (fx/hx/points),
``:278-282`` (construction, x/P/R/Q), ``:333-335`` (predict(ctrl), update(z));
``src/trajectorySimulateC.py:148-157,316-320,390-392``.

Known ambiguity: filterpy 1.4.5 regenerates the sigma points from
the predicted (x-,P-) at the end of ``predict``; older releases reused the
propagated points.  ``regen_sigmas`` selects (default True = 1.4.5).
"""
import numpy as np
import scipy.linalg as sla


def cholesky_psd_clamp(M):
    """Upper factor ``U'U = M`` that never raises: a pivot <= 0 leaves that row of U zero.
    NOT filterpy behaviour (scipy.linalg.cholesky raises LinAlgError there and the reference run
    dies); it is the engine's documented continuation for such lanes, restated here so that the
    parity tests can cover them.  Returns (U, clamped)."""
    n = M.shape[0]
    U = np.zeros((n, n))
    clamped = False
    for i in range(n):
        d = M[i, i] - U[:i, i] @ U[:i, i]
        if not d > 0.0:
            clamped = True
            continue
        r = np.sqrt(d)
        U[i, i] = r
        for j in range(i + 1, n):
            U[i, j] = (M[i, j] - U[:i, i] @ U[:i, j]) / r
    return U, clamped


class MerweScaledSigmaPointsRef:
    def __init__(self, n, alpha, beta, kappa, chol_fail='raise'):
        self.n, self.alpha, self.beta, self.kappa = n, alpha, beta, kappa
        self.chol_fail = chol_fail          # 'raise' = filterpy/scipy; 'clamp' = engine continuation
        self.clamped = False
        lam = alpha ** 2 * (n + kappa) - n
        c = 0.5 / (n + lam)
        self.Wc = np.full(2 * n + 1, c)
        self.Wm = np.full(2 * n + 1, c)
        self.Wc[0] = lam / (n + lam) + (1.0 - alpha ** 2 + beta)
        self.Wm[0] = lam / (n + lam)
        self.lam = lam

    def sigma_points(self, x, P):
        n = self.n
        try:
            U = sla.cholesky((self.lam + n) * P)   # upper, U'U = (n+lam) P
        except np.linalg.LinAlgError:
            if self.chol_fail != 'clamp':
                raise
            U, _ = cholesky_psd_clamp((self.lam + n) * P)
            self.clamped = True
        sig = np.zeros((2 * n + 1, n))
        sig[0] = x
        for k in range(n):
            sig[k + 1] = x + U[k]
            sig[n + k + 1] = x - U[k]
        return sig


class UKFRef:
    """``UnscentedKalmanFilter(dim_x, dim_z, dt, fx, hx, points)`` subset: predict/update
    with additive Q/R, plain-subtraction residuals, ``np.dot`` means (no angle wrap)."""

    def __init__(self, dim_x, dim_z, fx, hx, points, regen_sigmas=True):
        self.x = np.zeros(dim_x)
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.R = np.eye(dim_z)
        self.fx, self.hx, self.points = fx, hx, points
        self.Wm, self.Wc = points.Wm, points.Wc
        self.regen = regen_sigmas
        self.sigmas_f = np.zeros((2 * dim_x + 1, dim_x))

    @staticmethod
    def _ut(sigmas, Wm, Wc, noise_cov):
        x = Wm @ sigmas
        y = sigmas - x[None, :]
        P = y.T @ (np.diag(Wc) @ y)
        return x, P + noise_cov

    def predict(self, u):
        sig = self.points.sigma_points(self.x, self.P)
        for i, s in enumerate(sig):
            self.sigmas_f[i] = self.fx(s, u)
        self.x, self.P = self._ut(self.sigmas_f, self.Wm, self.Wc, self.Q)
        if self.regen:
            self.sigmas_f = self.points.sigma_points(self.x, self.P)

    def update(self, z):
        sig_h = np.atleast_2d([self.hx(s) for s in self.sigmas_f])
        zp, S = self._ut(sig_h, self.Wm, self.Wc, self.R)
        SI = np.linalg.inv(S)
        Pxz = np.zeros((self.sigmas_f.shape[1], sig_h.shape[1]))
        for i in range(self.sigmas_f.shape[0]):
            Pxz += self.Wc[i] * np.outer(self.sigmas_f[i] - self.x, sig_h[i] - zp)
        K = Pxz @ SI
        self.x = self.x + K @ (np.asarray(z, float) - zp)
        self.P = self.P - K @ (S @ K.T)
        self.S, self.K = S, K
