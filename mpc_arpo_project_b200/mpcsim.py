"""Parameter and telemetry objects of the drop-in boundary.

Same class names, constructor signatures and attribute names as the reference's
``src/mpcsim.py:13-176`` (``Noise``, ``SimConditions``, ``SimRun``, ``Debris``,
``MPCParams``, ``FailsafeParams``) so scripts written against the reference keep
working; the plotting helper ``figurePlotSave`` (``src/mpcsim.py:179-416``) is out of
scope.  ``BatchSimRun`` is new: the SoA result of a batched run.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional, Sequence, Tuple

import numpy as np
from scipy import sparse


class Noise:
    """Additive plant disturbance statistics (ref ``src/mpcsim.py:13-32``)."""

    def __init__(self, noise_std: Tuple[float, float], noise_length: float):
        self.noise_std = noise_std
        self.noise_length = noise_length

    def constructSigMat(self):
        sx, sy = self.noise_std
        out = np.zeros((4, 4))
        out[0, 0], out[1, 1] = sx, sy
        return out


class SimConditions:
    """Controller-independent conditions (ref ``src/mpcsim.py:35-73``).  ``hatch_ofst``
    is derived: 90 degrees for in-track approaches, else 0 (``:64``)."""

    def __init__(self, x0, xr, r_p: float, los_ang: float, r_tol: float, mean_mtn: float, time_stp: float,
                 isReject: bool, suc_cond: Tuple[float, float], noise: Optional[Noise] = None,
                 inTrack: bool = False, T_cont: float = float('nan'), T_final: int = 100, isDeltaV: bool = False):
        self.x0 = x0
        self.xr = xr
        self.r_p = r_p
        self.los_ang = los_ang
        self.r_tol = r_tol
        self.hatch_ofst = (90.0 if inTrack else 0.0) * (np.pi / 180)
        self.mean_mtn = mean_mtn
        self.time_stp = time_stp
        self.isReject = isReject
        self.suc_cond = suc_cond
        self.noise = noise
        self.inTrack = inTrack
        self.T_cont = T_cont
        self.T_final = T_final
        self.isDeltaV = isDeltaV


class SimRun:
    """Per-trajectory telemetry (ref ``src/mpcsim.py:75-97``)."""

    def __init__(self, i_term: int, isSuccess: bool, x_true_pcw, x_est, ctrl_hist, ctrlr_seq, noise_hist):
        self.i_term = i_term
        self.isSuccess = isSuccess
        self.x_true_pcw = x_true_pcw
        self.x_est = x_est
        self.ctrl_hist = ctrl_hist
        self.ctrlr_seq = ctrlr_seq
        self.noise_hist = noise_hist


class Debris:
    """Square debris bounding box (ref ``src/mpcsim.py:99-123``)."""

    def __init__(self, center: Tuple[float, float], side_length: float, detect_distance: float):
        self.center = center
        self.side_length = side_length
        self.detect_distance = detect_distance

    def constructVertArr(self):
        h = self.side_length / 2
        cx, cy = self.center
        return np.array([[cx + h, cy + h], [cx - h, cy + h], [cx - h, cy - h], [cx + h, cy - h]])


class MPCParams:
    """MPC tuning (ref ``src/mpcsim.py:127-157``).  ``swap_xy`` exchanges the x/y entries
    of the diagonal ``Q_state`` / ``R_input`` for in-track approaches (``:145-151``)."""

    def __init__(self, Q_state, R_input, R_slack, V_ecr, horizons, u_lim: Tuple[float, float], swap_xy: bool = False):
        if swap_xy:
            Qd = np.array(Q_state.toarray() if hasattr(Q_state, "toarray") else Q_state, float)
            Rd = np.array(R_input.toarray() if hasattr(R_input, "toarray") else R_input, float)
            Qd[[0, 1, 2, 3], [0, 1, 2, 3]] = Qd[[1, 0, 3, 2], [1, 0, 3, 2]]
            Rd[[0, 1], [0, 1]] = Rd[[1, 0], [1, 0]]
            Q_state, R_input = sparse.dia_array(Qd), sparse.dia_array(Rd)
        self.Q_state = Q_state
        self.R_input = R_input
        self.R_slack = R_slack
        self.V_ecr = V_ecr
        self.Nx = horizons["Nx"]
        self.Nc = horizons["Nc"]
        self.Nb = horizons["Nb"]
        self.u_lim = u_lim


class FailsafeParams:
    """Failsafe LQR / deadbeat tuning (ref ``src/mpcsim.py:160-176``)."""

    def __init__(self, Q_fail, R_fail, C_int, K_dead):
        self.Q_fail = Q_fail
        self.R_fail = R_fail
        self.C_int = C_int
        self.K_dead = K_dead


@dataclass
class BatchSimRun:
    """Result of a batched run: one lane per trajectory, SoA host arrays.

    ``x_true[4, T, B]``, ``x_est[6, T, B]``, ``ctrl_hist[2, T, B]`` (T = nsim+1 for the
    discrete simulator; decimated sample instants for the continuous one),
    ``ctrlr_seq[T, B]`` uint8 codes 1/2/3 as in ``trajectorySimulate.py:378-385``,
    ``i_term[B]``, ``isSuccess[B]``, ``final_dist[B]`` (= ``||x_true_pcw[:, i_term-1]-xr||``,
    the statistic of ``test/disturbRejComp.py:87-88``) plus solver telemetry.
    """
    i_term: np.ndarray
    isSuccess: np.ndarray
    final_dist: np.ndarray
    x_true: Optional[np.ndarray] = None
    x_est: Optional[np.ndarray] = None
    ctrl_hist: Optional[np.ndarray] = None
    ctrlr_seq: Optional[np.ndarray] = None
    status: Optional[np.ndarray] = None
    iters: Optional[np.ndarray] = None
    u_raw: Optional[np.ndarray] = None
    rho: Optional[np.ndarray] = None             # [T-1, B] ADMM step size after each solve (record 'rho')
    ukf_clamped: Optional[np.ndarray] = None     # [B] 1 where the reference's UKF would have raised LinAlgError
    x_true_sub: Optional[np.ndarray] = None      # continuous simulator, every substep: [4, NS, B] (record 'x_true_sub')
    ctrl_sub: Optional[np.ndarray] = None        # [2, NS, B]
    ctrlr_sub: Optional[np.ndarray] = None       # [NS, B]
    stats: dict = field(default_factory=dict)
    stats_vec: Optional[np.ndarray] = None       # the MPCB_NSTATS doubles ranks all-reduce (sum)

    @property
    def batch(self) -> int:
        return int(self.i_term.shape[0])
