"""Nx = 10 in-track delta-v batch (config 4's scenario at the short horizon): how much do sign-flip rebuilds cost there?"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import mpc_arpo_project_b200 as M
from oracle.gen_golden import make_params
case = dict(Nx=10, inTrack=True, isDeltaV=True, isReject=False, sigma=None, T_final=150)
sc, mp, fp, _ = make_params(M, case)
B = 8192
rng = np.random.default_rng(0)
x0 = np.stack([rng.uniform(-15, 15, B), 100 + rng.uniform(-10, 10, B), np.zeros(B), np.zeros(B)], axis=1)
eng = M.Engine(M.build_problem(sc, mp, fp, None))
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    r = M.trajectorySimulateBatch(sc, mp, fp, None, x0, None, engine=eng, record=())
    torch.cuda.synchronize()
    dt = time.time() - t0
c = eng.counters()
print(f"Nx=10 in-track dv, {B} lanes: {r.stats['qp_solves']/dt/1e6:.2f} M solves/s, {dt*1e3:.1f} ms, rebuilds/run {c['operator_rebuilds']/3:.0f}, solves {r.stats['qp_solves']:.0f}")
