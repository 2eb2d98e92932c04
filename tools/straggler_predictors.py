import sys, json, numpy as np
sys.path.insert(0, '/root/repo')
import mpc_arpo_project_b200 as M
import bench
wl = bench.WORKLOADS["config2"]
from oracle.gen_golden import make_params
sc, mp, fp, _ = make_params(M, wl["case"])
eng = M.Engine(M.build_problem(sc, mp, fp, None))
out = {}
for seed in (1234, 1235, 1236, 1237):
    x0, noise = bench.make_inputs(wl, 4096, seed)
    r = eng.simulate_discrete(x0, noise, 300, ("iters", "status"))
    it = r.iters.astype(np.int64).sum(0)
    nmax = (r.status == -2).sum(0)
    feats = {
        "absnoise_max": np.abs(noise).max(axis=(0, 1)),
        "absnoise_first": np.abs(noise[0]).max(0),
        "noise_norm_sum": np.sqrt((noise ** 2).sum(1)).sum(0),
        "noise_x_min": noise[:, 0, :].min(0), "noise_x_max": noise[:, 0, :].max(0),
        "noise_y_min": noise[:, 1, :].min(0), "noise_y_max": noise[:, 1, :].max(0),
        "x0_y": x0[1], "x0_x": x0[0],
    }
    top = np.argsort(-it)[:20]
    print("seed", seed, "total its", it.sum(), "max lane", it.max(), "top20 share", it[top].sum() / it.sum())
    for k, v in feats.items():
        c = np.corrcoef(v, it)[0, 1]
        # rank of the top-5 heaviest lanes when ordering by this feature (descending and ascending)
        rk_d = np.argsort(np.argsort(-v))[top[:5]]
        rk_a = np.argsort(np.argsort(v))[top[:5]]
        print(f"   {k:16s} corr {c:+.3f}  ranks desc {rk_d.tolist()}  asc {rk_a.tolist()}")
    print("   top5 lanes", top[:5].tolist(), "its", it[top[:5]].tolist(), "maxiter solves", nmax[top[:5]].tolist(), "noise of top1", noise[:, :, top[0]].round(2).tolist())
