#!/usr/bin/env python
"""Tiny runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

No torch import (numpy host arrays through the C ABI), a few lanes and control steps each, so that the 10-100x
slow-down of the tools stays within a minute.  Families: team kernel n = 81 / 121 / 161 (whole-loop mode), team kernel in
list mode + post kernel (continuous simulator), the block and tile kernels under the round loop, the per-lane debris
kernel, the QP seam.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpc_arpo_project_b200 as M                                   # noqa: E402
from mpc_arpo_project_b200.presets import make_params               # noqa: E402

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["team10", "team20", "team30", "list", "block", "tile", "wave", "debris", "qp"]
rng = np.random.default_rng(0)


def lanes(B, in_track=False):
    base = np.array([-10., 100., 0, 0]) if in_track else np.array([100., 10., 0, 0])
    return base[None, :] + np.concatenate([rng.uniform(-5, 5, (B, 2)), np.zeros((B, 2))], axis=1)


def discrete(case, B, nsteps, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    sc, mp, fp, debris = make_params(case)
    sig = case.get("sigma") or 0.0
    noise = sig * rng.standard_normal((nsteps // case.get("noise_length", 50) + 1, 2, B)) if sig else None
    r = M.trajectorySimulateBatch(sc, mp, fp, debris, lanes(B, case.get("inTrack", False)), noise, nsteps=nsteps)
    for k in (env or {}):
        os.environ.pop(k)
    return int(r.stats["qp_solves"]), int(r.stats["admm_iterations"])


if "team10" in which:
    print("team n=81 :", discrete(dict(Nx=10, sigma=0.75, noise_length=3), 12, 8), flush=True)
if "team20" in which:
    print("team n=121:", discrete(dict(Nx=20, inTrack=True, isDeltaV=True, isReject=False, sigma=None), 6, 6), flush=True)
if "team30" in which:
    print("team n=161:", discrete(dict(Nx=30, sigma=0.7, noise_length=3, isReject=False), 4, 5), flush=True)
if "block" in which:
    print("block     :", discrete(dict(Nx=10, sigma=0.3, noise_length=3), 12, 4, {"MPCB_SOLVER": "block"}), flush=True)
if "tile" in which:
    print("tile      :", discrete(dict(Nx=10, sigma=0.3, noise_length=3), 19, 4, {"MPCB_SOLVER": "tile"}), flush=True)
if "wave" in which and os.environ.get("MPCB_HAVE_WAVE"):
    print("wave      :", discrete(dict(Nx=10, sigma=0.3, noise_length=3), 19, 4, {"MPCB_SOLVER": "wave"}), flush=True)
if "debris" in which:
    print("debris    :", discrete(dict(Nx=10, sigma=0.3, noise_length=3, debris=((60., 0.), 5., 20)), 2, 3), flush=True)
if "list" in which:
    case = dict(Nx=10, sigma=0.0012, noise_length=2, T_cont=0.001, T_final=1.5)
    sc, mp, fp, _ = make_params(case)
    B = 5
    noise = 0.0012 * rng.standard_normal((np.arange(0, 1.5, 1.0).size, 2, B))
    r = M.trajectorySimulateCBatch(sc, mp, fp, None, lanes(B), noise)
    print("list mode :", int(r.stats["qp_solves"]), int(r.stats["admm_iterations"]), flush=True)
if "qp" in which:
    sc, mp, fp, _ = make_params(dict(Nx=10, sigma=0.1))
    eng = M.Engine(M.build_problem(sc, mp, fp, None))
    xh = np.zeros((6, 9))
    xh[0], xh[1] = 100 + rng.uniform(-5, 5, 9), 10 + rng.uniform(-5, 5, 9)
    u0, st, it = eng.qp_solve(xh)
    u0, st, it = eng.qp_solve(xh)
    print("qp seam   :", st.tolist(), it.tolist(), flush=True)
    eng.close()
print("done")
