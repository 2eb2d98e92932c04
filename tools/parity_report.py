#!/usr/bin/env python
"""Full-horizon parity table: every benchmarked workload, all 300 control steps, the CUDA engine
through the C ABI against the oracle on the same seeded lanes.  Runs on a GPU box:

    python tools/parity_report.py [--lanes 512] [--out gpurun_out/parity.json]

Discrete workloads (config 2, its sigma = 0.1 variant, config 4, config 5's cell) are compared with the lane-batched
oracle (``oracle/batched_ref.py``) and, on a few lanes, with the scalar OSQP-shaped oracle (``oracle/sim_ref.py``); the
continuous simulator (config 3) with the scalar oracle (RK4) on ``--lanes-c`` lanes.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mpc_arpo_project_b200 as M                                  # noqa: E402
from mpc_arpo_project_b200.presets import WORKLOADS, make_inputs, make_params     # noqa: E402
from oracle.batched_ref import simulate_discrete_batch             # noqa: E402
from oracle.parity import full_horizon_report, markdown_row, MARKDOWN_HEADER      # noqa: E402
from oracle.sim_ref import trajectory_simulate, trajectory_simulate_c              # noqa: E402


def discrete(name, B, seed, nsteps=None, floor_control=True):
    """-> report with three comparisons: (a) oracle with ITS OWN spectral tables (independent eigen-decomposition: the two
    reconstruct M(rho)^-1 to 1e-11..1e-9 only), (b) oracle fed the engine's tables (isolates the device arithmetic),
    (c) the control: the oracle against itself with x0 perturbed by 1e-13 relative -- the fraction of lanes ANY two float64
    implementations can be expected to keep bit-for-bit in their discrete decisions over the horizon."""
    wl = WORKLOADS[name]
    sc, mp, fp, _ = make_params(wl["case"])
    x0, noise = make_inputs(wl, B, seed)
    nsim = int(sc.T_final / sc.time_stp) if nsteps is None else nsteps
    x0T = np.ascontiguousarray(x0.T)
    t0 = time.time()
    prob = M.build_problem(sc, mp, fp, None)
    eng = M.Engine(prob)
    got = M.trajectorySimulateBatch(sc, mp, fp, None, x0T, noise, nsteps=nsim, engine=eng)
    eng.close()
    t1 = time.time()
    ref = simulate_discrete_batch(sc, mp, fp, x0T, noise, nsteps=nsim, chol_fail='clamp')
    t2 = time.time()
    rep = full_horizon_report(got, ref)
    refs = simulate_discrete_batch(sc, mp, fp, x0T, noise, nsteps=nsim, chol_fail='clamp', spectral=(prob.V, prob.lam))
    rep["shared_tables"] = full_horizon_report(got, refs)
    if floor_control:
        rng = np.random.default_rng(99)
        x0p = np.ascontiguousarray((x0 * (1 + 1e-13 * rng.standard_normal(x0.shape))).T)
        refp = simulate_discrete_batch(sc, mp, fp, x0p, noise, nsteps=nsim, chol_fail='clamp')
        ctl = _as_run(refp)
        rep["control_oracle_vs_perturbed_oracle"] = {k: v for k, v in full_horizon_report(ctl, ref).items()
                                                     if k in ("lanes", "exact_lanes", "exact_frac", "first_divergence_step")}
    rep["engine_s"], rep["oracle_s"] = t1 - t0, t2 - t1
    rep["seed"] = seed
    return rep, got, (sc, mp, fp, x0, noise)


def _as_run(ref):
    """An oracle result dressed as an engine result (engine layout [field, T, B])."""
    import types
    fin = np.array([ref["x_true"][max(int(t) - 1, 0), b] for b, t in enumerate(ref["i_term"])])
    return types.SimpleNamespace(iters=ref["iters"], status=ref["status"], ctrlr_seq=ref["ctrlr_seq"], i_term=ref["i_term"],
                                 ctrl_hist=ref["ctrl_hist"].transpose(2, 0, 1), x_true=ref["x_true"].transpose(2, 0, 1),
                                 rho=ref["rho_hist"], final_dist=np.linalg.norm(fin[:, :2], axis=1),
                                 isSuccess=np.zeros(len(fin), int))


def scalar_check(got, ctx, lanes, wl):
    """A few lanes of the same run against the scalar oracle (per-step OSQP update + KKT LU)."""
    sc, mp, fp, x0, noise = ctx
    case = wl["case"]
    exact, worst_u = 0, 0.0
    first = []
    for b in lanes:
        sc.x0 = x0[:, b].copy()
        if noise is not None:
            d = np.concatenate([noise[:, :, b] / case["sigma"], np.zeros((noise.shape[0], 2))], axis=1)
        else:
            d = np.zeros((2, 4))
        it = iter(d)
        r = trajectory_simulate(sc, mp, fp, None, draw=lambda: next(it, np.zeros(4)), chol_fail='clamp')
        T = min(int(r.i_term), int(got.i_term[b]))
        gi, gs = np.asarray(got.iters[:T, b], int), np.asarray(got.status[:T, b], int)
        ri, rs = np.asarray(r.iters[:T], int), np.asarray(r.status_val[:T], int)
        bad = np.nonzero((gi != ri) | (gs != rs))[0]
        f = int(bad[0]) if bad.size else T
        if not bad.size and int(r.i_term) == int(got.i_term[b]):
            exact += 1
        else:
            first.append(f)
        worst_u = max(worst_u, float(np.nanmax(np.abs(got.ctrl_hist[:, :f + 1, b] - r.ctrl_hist[:, :f + 1]))))
    return {"lanes": len(lanes), "exact_lanes": exact, "max_du_prefix": worst_u, "first_divergence_steps": first}


def continuous(B, seed):
    wl = WORKLOADS["config3"]
    case = wl["case"]
    sc, mp, fp, _ = make_params(case)
    x0, noise = make_inputs(wl, B, seed)
    got = M.trajectorySimulateCBatch(sc, mp, fp, None, np.ascontiguousarray(x0.T), noise)
    exact, worst_u, worst_x, solves, first = 0, 0.0, 0.0, 0, []
    for b in range(B):
        sc.x0 = x0[:, b].copy()
        r = trajectory_simulate_c(sc, mp, fp, None, V=noise[:, :, b].T, integrator='rk4', chol_fail='clamp')
        ns = len(r.iters)
        solves += ns
        gi, gs = np.asarray(got.iters[:ns, b], int), np.asarray(got.status[:ns, b], int)
        bad = np.nonzero((gi != np.asarray(r.iters, int)) | (gs != np.asarray(r.status_val, int)))[0]
        f = int(bad[0]) if bad.size else ns
        if not bad.size and int(got.i_term[b]) == int(r.i_term):
            exact += 1
        else:
            first.append(f)
        if f:
            worst_u = max(worst_u, float(np.nanmax(np.abs(got.u_raw[:, :f, b] - r.u_raw[:, :f]))))
            for j, i_sub in enumerate(r.solve_at[:f]):
                worst_x = max(worst_x, float(np.max(np.abs(got.x_true[:, j + 1, b] - r.x_true[:, i_sub + 1]))))
    return {"lanes": B, "solves": solves, "exact_lanes": exact, "exact_frac": exact / B, "max_du_prefix": worst_u,
            "max_dx_prefix": worst_x, "first_divergence_steps": first, "seed": seed}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=512)
    ap.add_argument("--lanes-c", type=int, default=8)
    ap.add_argument("--scalar-lanes", type=int, default=16)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--only", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.json"))
    args = ap.parse_args()
    res = {}
    rows = [MARKDOWN_HEADER]
    for name in ("config2", "config2_quiet", "config4", "config5_cell"):
        if args.only and name not in args.only.split(","):
            continue
        B = args.lanes if name != "config5_cell" else max(64, args.lanes // 2)
        rep, got, ctx = discrete(name, B, args.seed)
        if args.scalar_lanes:
            rep["scalar_oracle"] = scalar_check(got, ctx, list(range(0, B, max(1, B // args.scalar_lanes)))[:args.scalar_lanes],
                                                WORKLOADS[name])
        res[name] = rep
        rows.append(markdown_row(name + " (oracle's own tables)", rep))
        rows.append(markdown_row(name + " (engine's spectral tables)", rep["shared_tables"]))
        if "control_oracle_vs_perturbed_oracle" in rep:
            c = rep["control_oracle_vs_perturbed_oracle"]
            rows.append(f"| {name}: control, oracle vs oracle with x0*(1+1e-13) | {c['lanes']} x {rep['steps']} | {c['exact_lanes']} "
                        f"({100 * c['exact_frac']:.1f} %) | | | | | | |")
        print(json.dumps({name: rep}), flush=True)
    if not args.only or "config3" in args.only.split(","):
        rep = continuous(args.lanes_c, args.seed)
        res["config3"] = rep
        print(json.dumps({"config3": rep}), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    open(os.path.splitext(args.out)[0] + ".md", "w").write("\n".join(rows) + "\n")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
