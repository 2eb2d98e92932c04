"""Parameter sets of the reference's scripts and the synthetic workloads of SURVEY.md section 8(d).

``make_params(case)`` restates the literals of ``test/traj_eval_radial.py:17-72``, ``test/traj_eval_radialC.py:17-75``,
``test/traj_eval_in_track.py:14-66`` and ``test/disturbRejComp.py:17-72`` (weights, horizons, cone, thrust limit,
failsafe weights) as ``mpcsim`` objects; ``case`` overrides what the scripts vary (horizon, noise, plant, debris).
``WORKLOADS`` / ``make_inputs`` are the benchmark batches (BASELINE.json configs 2-5); ``bench.py``, the Monte-Carlo
helpers and the tests all draw their lanes from here so that every arm sees the same distribution.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse

from . import mpcsim as _mpcsim


def make_params(case: dict, M=None):
    """-> (SimConditions, MPCParams, FailsafeParams, Debris | None).  ``M``: a module exposing the mpcsim classes
    (default: this package's; the golden-fixture generator passes the reference's own ``src.mpcsim``)."""
    M = M or _mpcsim
    Q = 8e+02 * sparse.diags([0.2 ** 2., 10 ** 2., 3.8 ** 2, 900.])
    R = 1000 ** 2 * sparse.diags([1., 1.])
    Rs = 5 ** 2 * sparse.eye(5)
    v = 50000 * np.ones(5)
    v[-2] = -v[-2]
    v[-1] = 0
    fp = M.FailsafeParams(0.005 * np.diag([0.0001, 1, 100000., 1., 0.01]), 100 * np.diag([1, 1]), np.eye(1, 4), np.zeros([2, 2]))
    c = dict(case)
    Nx = c.get('Nx', 10)
    in_track = c.get('inTrack', False)
    if in_track:
        x0 = np.array(c.get('x0', [-10., 100., 0., 0.]))
        xr = np.array([0., 2.5, 0., 0.])
        Rs = 5 ** 2 * sparse.diags([1.5, 1.5, 1, 1, 1e5])
        v[-1] = 1e-09
    else:
        x0 = np.array(c.get('x0', [100., 10., 0., 0.]))
        xr = np.array([2.5, 0., 0., 0.])
    noise = M.Noise((c['sigma'], c['sigma']), c.get('noise_length', 50)) if c.get('sigma') else None
    sc = M.SimConditions(x0, xr, 2.5, 10 * (np.pi / 180), 1.5, 1.107e-3, 0.5, c.get('isReject', True), (0.2, 45), noise,
                         in_track, T_cont=c.get('T_cont', float('nan')), T_final=c.get('T_final', 150),
                         isDeltaV=c.get('isDeltaV', False))
    mp = M.MPCParams(Q, R, Rs, v, {"Nx": Nx, "Nc": 5, "Nb": 5}, (0.2, 0.2), swap_xy=in_track)
    debris = M.Debris(*c['debris']) if c.get('debris') else None
    return sc, mp, fp, debris


# disturbRejComp.py:75-82: the ten disturbance hold lengths of the sweep
DISTURB_REJ_LENGTHS = (1, 10, 20, 30, 50, 70, 100, 150, 200, 250)

# SURVEY.md section 8(d) synthetic workloads.  lanes = per-GPU shard.
WORKLOADS = {
    "config1": dict(kind="D", lanes=8192, case=dict(Nx=40, sigma=0.75, noise_length=50, T_final=150, debris=((40., 0.), 5., 20)),
                    desc="test/traj_eval_radial.py as shipped (Nx=40, debris (40,0,5,20), sigma=0.75 held 50 steps, 300 steps), batched"),
    "config2": dict(kind="D", lanes=4096, case=dict(Nx=10, sigma=0.75, noise_length=50, T_final=150),
                    desc="trajectorySimulate batched: 4096 linear-CW radial lanes, Nx=10, sigma=0.75 held 50 steps, 300 steps"),
    "config2_quiet": dict(kind="D", lanes=4096, case=dict(Nx=10, sigma=0.1, noise_length=50, T_final=150),
                          desc="config2 with sigma=0.1 (MPC stays feasible: solver-throughput variant)"),
    "config3": dict(kind="C", lanes=65536, case=dict(Nx=10, sigma=0.0012, noise_length=50, T_cont=0.001, T_final=150),
                    desc="trajectorySimulateC batched: 65536 nonlinear-plant lanes, accel inputs, RK4 h=1ms"),
    "config4": dict(kind="D", lanes=32768, case=dict(Nx=20, inTrack=True, isDeltaV=True, isReject=False, sigma=None, T_final=150),
                    desc="in-track delta-v sweep, Nx=20, 262144 lanes over 8 GPUs (32768 per GPU)"),
    "config5": dict(kind="S", lanes=131072, case=dict(Nx=30, sigma=0.7, noise_length=50, T_final=150),
                    desc="disturbRejComp Monte Carlo: 10 hold lengths x {reject, no reject}, Nx=30, 1M lanes over 8 GPUs "
                         "(131072 per GPU = 20 cells x 6553 realisations)"),
    "config5_cell": dict(kind="D", lanes=131072, case=dict(Nx=30, sigma=0.7, noise_length=50, T_final=150),
                         desc="one disturbRejComp cell (hold 50, reject), Nx=30, 131072 lanes per GPU"),
}


def make_inputs(wl: dict, B: int, seed: int):
    """Synthetic lanes of SURVEY.md 8(d): x0 = nominal + U(-10,10) x U(-5,5) (in-track: U(-15,15) x 100+U(-10,10));
    N(0,1)*sigma disturbances, one row per hold interval.  -> (x0[4, B], noise[R, 2, B] | None)."""
    case = wl["case"]
    rng = np.random.default_rng(seed)
    if case.get("inTrack"):
        x0 = np.stack([rng.uniform(-15, 15, B), 100 + rng.uniform(-10, 10, B), np.zeros(B), np.zeros(B)])
    else:
        x0 = np.stack([100 + rng.uniform(-10, 10, B), 10 + rng.uniform(-5, 5, B), np.zeros(B), np.zeros(B)])
    sig = case.get("sigma")
    noise = None
    if sig:
        T, Tf, nl = 0.5, case["T_final"], case["noise_length"]
        R = (int(Tf / T) // nl + 1) if wl["kind"] != "C" else np.arange(0, Tf, T * nl).size
        noise = sig * rng.standard_normal((R, 2, B))
    return np.ascontiguousarray(x0), noise
