#!/usr/bin/env python
"""Dump the engine's full telemetry of a workload slice to an .npz (for offline comparison with the oracle)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpc_arpo_project_b200 as M
from mpc_arpo_project_b200.presets import WORKLOADS, make_inputs, make_params
name, B, seed, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
wl = WORKLOADS[name]
sc, mp, fp, _ = make_params(wl["case"])
x0, noise = make_inputs(wl, B, seed)
got = M.trajectorySimulateBatch(sc, mp, fp, None, np.ascontiguousarray(x0.T), noise)
np.savez_compressed(out, x0=x0, noise=noise if noise is not None else np.zeros(0), i_term=got.i_term, iters=got.iters, status=got.status,
                    ctrlr_seq=got.ctrlr_seq, ctrl_hist=got.ctrl_hist, u_raw=got.u_raw, x_true=got.x_true, x_est=got.x_est,
                    ukf_clamped=got.ukf_clamped, final_dist=got.final_dist, isSuccess=got.isSuccess)
print(name, B, seed, "->", out, int(got.stats["qp_solves"]), int(got.stats["admm_iterations"]))
