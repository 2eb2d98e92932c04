"""How much of the full-horizon 'diverged' count is the rho criterion of oracle/parity.py?  (GPU; test tooling.)"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_arpo_project_b200 as M
from mpc_arpo_project_b200.presets import WORKLOADS, make_inputs, make_params
from oracle.batched_ref import simulate_discrete_batch
from oracle import c_ref
import oracle.parity as P

out = {}
for name, B in (("config2", 256), ("config2_quiet", 256), ("config4", 192), ("config5_cell", 128)):
    wl = WORKLOADS[name]
    sc, mp, fp, _ = make_params(wl["case"])
    x0, noise = make_inputs(wl, B, 4321)
    x0T = np.ascontiguousarray(x0.T)
    prob = M.build_problem(sc, mp, fp, None)
    with M.Engine(prob) as eng:
        got = M.trajectorySimulateBatch(sc, mp, fp, None, x0T, noise, engine=eng)
    shared = simulate_discrete_batch(sc, mp, fp, x0T, noise, chol_fail='clamp', spectral=(prob.V, prob.lam))
    own = simulate_discrete_batch(sc, mp, fp, x0T, noise, chol_fail='clamp')
    twin = c_ref.simulate_discrete(prob, x0, noise, 300, nthreads=0)
    it_g = np.asarray(got.iters).astype(int); st_g = np.asarray(got.status).astype(int); rho_g = np.asarray(got.rho)
    res = {}
    for nm, ref in (("shared", shared), ("own", own), ("twin", twin)):
        live = np.arange(300)[:, None] < np.minimum(np.asarray(got.i_term), ref["i_term"])[None, :]
        dis = ((it_g != ref["iters"]) | (st_g != ref["status"])) & live
        first_dis = np.where(dis.any(0), dis.argmax(0), 300)
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.abs(rho_g - ref["rho_hist"]) / np.abs(ref["rho_hist"])
        rel = np.where(live, np.nan_to_num(rel, nan=0.0), 0.0)
        r = {"discrete_exact": float((first_dis >= 300).mean())}
        for tol in (1e-6, 1e-4, 1e-2):
            bad = (rel > tol)
            first_rho = np.where(bad.any(0), bad.argmax(0), 300)
            r[f"exact_with_rho_tol_{tol:g}"] = float(((first_dis >= 300) & (first_rho >= 300)).mean())
            r[f"rho_first_before_discrete_{tol:g}"] = int((first_rho < first_dis).sum())
        # max relative rho gap on the prefix where the discrete record still matches
        pre = np.arange(300)[:, None] < first_dis[None, :]
        r["max_rel_rho_gap_on_discrete_prefix"] = float((rel * pre).max())
        res[nm] = r
    # oracle (own tables) against the C twin: two CPU implementations
    live = np.arange(300)[:, None] < np.minimum(own["i_term"], twin["i_term"])[None, :]
    dis = ((own["iters"] != twin["iters"]) | (own["status"] != twin["status"])) & live
    res["own_vs_twin_discrete_exact"] = float((~dis.any(0)).mean())
    dis = ((shared["iters"] != own["iters"]) | (shared["status"] != own["status"])) & (np.arange(300)[:, None] < np.minimum(own["i_term"], shared["i_term"])[None, :])
    res["shared_vs_own_discrete_exact"] = float((~dis.any(0)).mean())
    out[name] = res
    print(name, json.dumps(res), flush=True)
json.dump(out, open("gpurun_out/r2b_parity_rho_probe.json", "w"), indent=1)
