import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import mpc_arpo_project_b200 as M
from oracle.gen_golden import make_params
for Nx, B in ((10, 1184), (30, 592), (40, 296)):
    case = dict(Nx=Nx, sigma=0.7, noise_length=50, T_final=150, debris=((40., 0.), 5., 20))
    sc, mp, fp, debris = make_params(M, case)
    rng = np.random.default_rng(0)
    x0 = np.array([100., 10., 0, 0])[None, :] + np.concatenate([rng.uniform(-5, 5, (B, 2)), np.zeros((B, 2))], axis=1)
    noise = 0.7 * rng.standard_normal((7, 2, B))
    eng = M.Engine(M.build_problem(sc, mp, fp, debris))
    r = M.trajectorySimulateBatch(sc, mp, fp, debris, x0, noise, engine=eng, record=())
    torch.cuda.synchronize()
    t0 = time.time()
    r = M.trajectorySimulateBatch(sc, mp, fp, debris, x0, noise, engine=eng, record=())
    torch.cuda.synchronize()
    dt = time.time() - t0
    s = r.stats
    print(f"debris lanes Nx={Nx} B={B}: {s['qp_solves']/dt:9.0f} solves/s, {dt*1e3:7.1f} ms, {s['admm_iterations']/max(1,s['qp_solves']):.1f} its/solve, {s['qp_solves']/B:.0f} steps/lane")
    eng.close()
