#!/usr/bin/env python
"""Benchmark of the batched closed-loop MPC engine (BASELINE.json metric: closed-loop MPC QP
solves per second = live trajectory control steps per second).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2] [--impl b200|reference]

One "step" = one full pass of the hot path over one batch: a whole closed-loop simulation of the
workload (config2 = 4096 linear-CW radial trajectories, 10-step horizon, 300 control steps, one QP
per live trajectory per control step; config5 = the whole disturbRejComp sweep, 10 hold lengths x
{reject, no reject}).  Every step draws NEW lanes: seed = 1234 + step * world + rank, so N = 1 and
N = 8 see the same lane distribution and no figure rests on one lucky seed.  For N > 1 the driver
launches this file under torchrun; every rank runs its own shard of lanes (weak scaling), steps do
not synchronise the ranks, and the only collective is ONE NCCL all-reduce of the accumulated
MPCB_NSTATS statistics after the last step, inside the timed region.

Prints ONE JSON line on rank 0 (see README / DESIGN.md section "Measurement").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mpc_arpo_project_b200.presets import DISTURB_REJ_LENGTHS, WORKLOADS, make_inputs, make_params    # noqa: E402

METRIC = "closed-loop MPC QP solves/sec (traj-steps/s)"
UNIT = "solves/s"
SEED0 = 1234


def f_it(n, m, nnzA):
    """Algorithmic flops per ADMM iteration per trajectory, SURVEY.md section 8(d): 2n^2 + 4 nnz(A) + 12 m + 8 n."""
    return 2 * n * n + 4 * nnzA + 12 * m + 8 * n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid, wait_first=8.0):
        self.p, self.lines = None, []
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={uuid}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
            return
        import threading
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()
        # nvidia-smi's start-up (NVML init, a second or more) stalls the GPU it attaches to: wait for its first sample so
        # that none of it lands in the timed region; the steady polling that follows does not show in the timings
        t0 = time.time()
        while not self.lines and time.time() - t0 < wait_first and self.p.poll() is None:
            time.sleep(0.05)

    def _pump(self):
        for line in self.p.stdout:
            self.lines.append(line)

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.t.join(timeout=2)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in "".join(self.lines).strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s in sm if s > 0.5 * max(sm)] or sm          # "under load" = samples above the idle clock
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU arms (oracle = checker / baseline)
def _cell_problems(wl):
    """The problem families a workload runs: one for kinds D / C, 20 (hold length x mode) for the sweep."""
    import copy
    import mpc_arpo_project_b200 as M
    from mpc_arpo_project_b200.mpcsim import Noise
    sc, mp, fp, _ = make_params(wl["case"])
    if wl["kind"] != "S":
        return [(sc, mp, fp)]
    out = []
    for nl in DISTURB_REJ_LENGTHS:
        for rej in (False, True):
            s2 = copy.copy(sc)
            s2.isReject = rej
            s2.noise = Noise(tuple(sc.noise.noise_std), int(nl))
            out.append((s2, mp, fp))
    return out


def cpu_twin_run(wl, lanes, threads, seed):
    """The compiled oracle twin (oracle/c/mpc_ref.c: per-trajectory loop, dense Cholesky of the reduced KKT system, one
    trajectory per thread) on `lanes` lanes of the workload's distribution.  Returns (solves, wall seconds)."""
    import mpc_arpo_project_b200 as M
    from oracle import c_ref
    cells = _cell_problems(wl)
    per = max(1, lanes // len(cells))
    solves, wall = 0, 0.0
    for ci, (sc, mp, fp) in enumerate(cells):
        prob = M.build_problem(sc, mp, fp, None)
        nsteps = int(sc.T_final / sc.time_stp)
        w2 = dict(wl, kind="D", case=dict(wl["case"], noise_length=int(sc.noise.noise_length) if sc.noise is not None else 50))
        x0, noise = make_inputs(w2, per, seed + 101 * ci)
        t0 = time.perf_counter()
        r = c_ref.simulate_discrete(prob, x0, noise, nsteps, nthreads=threads, record=False)
        wall += time.perf_counter() - t0
        solves += r["qp_solves"]
    return solves, wall


def _py_worker(args):
    """One trajectory of the restated Python reference path (oracle/sim_ref.py) -- the continuous simulator's CPU arm."""
    case, x0, draws, solver = args
    from oracle.sim_ref import trajectory_simulate, trajectory_simulate_c
    sc, mp, fp, _ = make_params(case)
    sc.x0 = np.array(x0, float)
    t0 = time.perf_counter()
    if "T_cont" in case:
        r = trajectory_simulate_c(sc, mp, fp, None, V=draws, integrator="rk4", chol_fail="clamp")
        solves = len(r.iters)
    else:
        it = iter(draws)
        r = trajectory_simulate(sc, mp, fp, None, draw=lambda: next(it, np.zeros(4)), chol_fail="clamp", solver=solver)
        solves = int(r.i_term)
    return solves, time.perf_counter() - t0


def cpu_python_run(wl, lanes, cores, seed, solver="restated"):
    """The per-trajectory Python loop of the oracle (numpy + restated OSQP / UKF, or the REAL osqp if importable), one
    process per core."""
    import multiprocessing as mpx
    case = dict(wl["case"])
    x0, noise = make_inputs(wl, lanes, seed)
    jobs = []
    for b in range(lanes):
        if noise is None:
            d = np.zeros((2, 4)) if wl["kind"] != "C" else None
        elif wl["kind"] != "C":
            d = np.concatenate([noise[:, :, b] / case["sigma"], np.zeros((noise.shape[0], 2))], axis=1)
        else:
            d = noise[:, :, b].T
        jobs.append((case, x0[:, b], d, solver))
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"
    try:
        t0 = time.perf_counter()
        if cores > 1:
            with mpx.get_context("spawn").Pool(cores) as pool:
                pool.map(_py_worker, jobs[:cores], chunksize=1)          # interpreter / scipy import cost outside the timing
                t0 = time.perf_counter()
                res = pool.map(_py_worker, jobs, chunksize=1)
        else:
            res = [_py_worker(j) for j in jobs]
        wall = time.perf_counter() - t0
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return sum(r[0] for r in res), wall


def cpu_arm(wl, lanes, cores, seed):
    """-> (solves, wall, kind, description).  Discrete workloads: the compiled twin; the continuous simulator: the Python
    restatement (the twin has no RK4 plant); the real osqp package if it can be imported."""
    from oracle.sim_ref import real_osqp_available
    if real_osqp_available() and wl["kind"] == "D":
        s, w = cpu_python_run(wl, lanes, cores, seed, solver="osqp")
        return s, w, "reference", "the reference's per-trajectory Python loop on the REAL osqp package (oracle/sim_ref.py, solver='osqp')"
    if wl["kind"] == "C":
        s, w = cpu_python_run(wl, lanes, cores, seed)
        return s, w, "port", "oracle/sim_ref.py (restated per-trajectory Python + OSQP loop, RK4 plant)"
    s, w = cpu_twin_run(wl, lanes, cores, seed)
    note = ""
    if wl["case"].get("debris"):
        note = ("; the twin runs the DEBRIS-FREE problem of the same horizon (it has no obstacle geometry): a lower bound on the "
                "reference's cost per solve, which re-equilibrates and refactors at every step once a Debris object exists")
    return s, w, "port", ("oracle/c/mpc_ref.c (compiled per-trajectory loop of the reference: OSQP-equivalent ADMM with a dense "
                          "Cholesky of the reduced KKT system, UKF, one trajectory per thread; osqp / filterpy / control are not "
                          "installable offline)" + note)


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    lanes = args.ref_lanes or (cores if wl["kind"] == "C" else (10 * cores if wl["kind"] == "S" else 128 * cores))
    lanes = max(cores, lanes)
    times, solves, kind, desc = [], 0, "port", ""
    for i in range(args.warmup + args.steps):
        s, w, kind, desc = cpu_arm(wl, lanes, cores, SEED0 + i)
        if i >= args.warmup:
            times.append(w)
            solves += s
    value = solves / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "lanes_per_step": lanes, "note": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{lanes} lanes of {args.workload} per step, full horizon"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--lanes", type=int, default=None, help="lanes per GPU (default: the workload's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-lanes", type=int, default=0, help="reference arm: lanes per step (0: 128 per core, 1 per core for the continuous simulator)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-lanes", type=int, default=64, help="lanes of the full-horizon parity slice (0 = skip)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, wl)
    if args.warmup < 3:
        print("[bench] warning: timing rules ask for >= 3 warm-up steps", file=sys.stderr)

    import torch
    import torch.distributed as dist
    import mpc_arpo_project_b200 as M
    from mpc_arpo_project_b200 import _lib
    from mpc_arpo_project_b200.trajectorySimulateC import continuous_grid, noise_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL announces its version on STDOUT when the first communicator comes up; the contract is ONE JSON line there
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device(f"cuda:{local}")

    B = args.lanes or wl["lanes"]
    kind = wl["kind"]
    case = wl["case"]
    sc, mp, fp, debris = make_params(case)
    record = ("x_true", "x_est", "ctrl", "ctrlr_seq")          # the SimRun fields of the reference
    plan = None
    if kind == "D":
        prob = M.build_problem(sc, mp, fp, debris)
        nsteps = int(sc.T_final / sc.time_stp)
        engines = [M.Engine(prob, device=local, pin_outputs=True)]
    elif kind == "C":
        prob = M.build_problem_c(sc, mp, fp, None)
        nsimD, nsimC, ratio = continuous_grid(sc)
        _, hold = noise_plan(sc)
        engines = [M.Engine(prob, device=local, pin_outputs=True)]
    else:                                                      # the disturbRejComp sweep: 10 hold lengths x 2 modes
        plan = M.RatioSweep(sc, mp, fp, None, sc.noise.noise_std, DISTURB_REJ_LENGTHS, device=local, pin_outputs=True)
        engines = [e for row in plan.engines for e in row]
        prob = engines[0].problem
        Bc = B // len(engines)                                 # realisations per cell
        B = Bc * len(engines)
        record = ()                                            # the sweep keeps statistics only (disturbRejComp.py:87-98)
    eng = engines[0]
    for e in engines:
        e.batch_alloc(B if plan is None else Bc)

    # ---- inputs of every step, generated before any timing: pinned host copies (e2e leg) and device copies (value leg)
    n_dev, n_e2e = args.warmup + args.steps, 1 + args.steps
    step_inputs = {}

    def inputs(step):
        if step not in step_inputs:
            seed = SEED0 + step * world + rank
            if plan is None:
                x0_h, noise_h = make_inputs(wl, B, seed)
                hp = (torch.from_numpy(x0_h).pin_memory(), torch.from_numpy(noise_h).pin_memory() if noise_h is not None else None)
            else:
                rng = np.random.default_rng(seed)
                x0_h = np.ascontiguousarray(np.tile(np.asarray(sc.x0, float)[:, None], (1, Bc)))      # the script's fixed x0
                sig = np.asarray(sc.noise.noise_std, float)
                hp = (torch.from_numpy(x0_h).pin_memory(),
                      [torch.from_numpy(rng.standard_normal((plan.refreshes(i), 2, Bc)) * sig[None, :, None]).pin_memory()
                       for i in range(len(plan.noise_lengths))])
            if plan is None:
                dv = (hp[0].to(dev), hp[1].to(dev) if hp[1] is not None else None)
            else:
                dv = (hp[0].to(dev), [t.to(dev) for t in hp[1]])
            step_inputs[step] = (hp, dv)
        return step_inputs[step]

    for s in range(n_dev + n_e2e + 2):
        inputs(s)
    h2d = inputs(0)[0][0].numel() * 8 + (0 if inputs(0)[0][1] is None else
                                         (inputs(0)[0][1].numel() * 8 if plan is None else sum(t.numel() * 8 for t in inputs(0)[0][1])))
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stats_t = torch.zeros((len(engines), _lib.NSTATS), dtype=torch.float64, device=dev)

    def step(s, on_device):
        """-> (qp solves of this rank, statistics [cells, NSTATS], last result)."""
        hp, dv = inputs(s)
        if plan is not None:
            x0, nz = (dv[0], dv[1]) if on_device else (hp[0].numpy(), [t.numpy() for t in hp[1]])
            out = plan.run(x0, nz)
            return int(out["qp_solves"]), out["stats"].reshape(len(engines), -1), out
        x0, nz = (dv if on_device else (hp[0].numpy(), hp[1].numpy() if hp[1] is not None else None))
        if kind == "D":
            res = eng.simulate_discrete(x0, nz, nsteps, record)
        else:
            res = eng.simulate_continuous(x0, nz, nsimC, ratio, float(sc.T_cont), hold, record)
        return int(res.stats["qp_solves"]), res.stats_vec[None, :], res

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(first, nrep, on_device):
        """nrep steps starting at input index `first`; returns (device ms summed over steps, solves, last result).  L2 is
        flushed between steps.  Ranks run their steps independently (lanes are independent; a Monte-Carlo job needs its
        statistics once, at the end): the all-reduce of the accumulated MPCB_NSTATS statistics -- the path's only collective --
        follows the last step INSIDE the timed region: its event pair is recorded on torch's stream, which NCCL runs on."""
        tot_ms, solves, res = 0.0, 0, None
        cur = torch.cuda.current_stream(dev)
        acc = np.zeros((len(engines), _lib.NSTATS))
        for k in range(nrep):
            flush.fill_(1)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sv, st, res = step(first + k, on_device)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            tot_ms += e0.elapsed_time(e1)
            acc += st
            if os.environ.get("BENCH_VERBOSE"):
                print(f"[bench] step {first + k} (seed {SEED0 + (first + k) * world + rank}): {e0.elapsed_time(e1):.2f} ms, {sv} solves", file=sys.stderr)
            solves += sv
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            stats_t.copy_(torch.from_numpy(np.ascontiguousarray(acc)), non_blocking=True)
            dist.all_reduce(stats_t)
            e1.record(cur)
            torch.cuda.synchronize(dev)
            tot_ms += e0.elapsed_time(e1)       # includes the wait for the slowest rank: max over ranks is taken below anyway
        return tot_ms, solves, res

    # the sampler starts BEFORE the warm-up: nvidia-smi's own start-up stalls the GPU it queries
    sampler = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid)) if rank == 0 else None
    keep = None
    for s in range(args.warmup):
        keep = step(s, True)    # the previous result stays alive while the next step allocates, exactly as in the timed loop
    del keep
    c0 = [e.counters() for e in engines]
    sync_all()
    ms, solves, res = timed(args.warmup, args.steps, True)
    sync_all()
    clocks = sampler.stop() if sampler else None
    c1 = [e.counters() for e in engines]

    # end-to-end leg: same steps through the public API with pinned HOST buffers (H2D + D2H inside)
    # (the SAME lanes as the device-resident leg, step by step: per-seed step times spread 104-194 ms on config 2, so two legs on
    # different seeds would not be comparable)
    step(n_dev, False)
    sync_all()
    ms_e2e, solves_e2e, res_h = timed(args.warmup, args.steps, False)
    sync_all()
    if plan is None:
        d2h = sum(getattr(res_h, k).nbytes for k in ("x_true", "x_est", "ctrl_hist", "ctrlr_seq", "i_term", "isSuccess",
                                                     "final_dist", "ukf_clamped")) + 8 * _lib.NSTATS
    else:
        d2h = len(engines) * (Bc * (4 + 4 + 8 + 4) + 8 * _lib.NSTATS)       # i_term, isSuccess, final_dist, ukf_clamped + statistics

    # roofline leg: one more device-resident step with per-launch CUDA events around the ADMM kernels
    for e in engines:
        e.set_timing(True)
    t0 = [e.counters() for e in engines]
    step(n_dev + n_e2e, True)
    t1 = [e.counters() for e in engines]
    for e in engines:
        e.set_timing(False)

    def delta(a, b, key):
        return sum(y[key] - x[key] for x, y in zip(a, b))

    # solver status mix and iterations-per-solve histogram (SURVEY 8(d) caveat) + full-horizon parity slice, untimed
    mix, parity = None, None
    if rank == 0 and kind == "D":
        hp, dv = inputs(n_dev + n_e2e + 1)
        r2 = eng.simulate_discrete(dv[0], dv[1], nsteps, tuple(record) + ("status", "iters"))
        st = np.asarray(r2.status.cpu() if hasattr(r2.status, "cpu") else r2.status).astype(np.int64)
        it = np.asarray(r2.iters.cpu() if hasattr(r2.iters, "cpu") else r2.iters).astype(np.int64)
        iterm = np.asarray(r2.i_term.cpu() if hasattr(r2.i_term, "cpu") else r2.i_term).astype(np.int64)
        live = np.arange(st.shape[0])[:, None] < iterm[None, :]
        names = {1: "solved", 2: "solved_inaccurate", -3: "primal_infeasible", 3: "primal_infeasible_inaccurate",
                 -2: "max_iter_reached"}
        tot = max(1, int(live.sum()))
        edges = [25, 50, 75, 100, 200, 500, 1000, 3999, 4000]
        hist, lo = {}, 0
        for e in edges:
            hist[f"<={e}"] = float(((it > lo) & (it <= e) & live).sum() / tot)
            lo = e
        mix = {"status_fraction": {nm: float(((st == k) & live).sum() / tot) for k, nm in names.items()},
               "iters_per_solve_hist": hist, "max_iterations_one_lane": int((it * live).sum(axis=0).max())}
    if rank == 0 and args.parity_lanes > 0 and kind in ("D", "S") and not prob.has_debris:
        # the oracle as the CHECKER (never timed, never on the product path): a slice of the workload, all 300 steps, lane by
        # lane (oracle/parity.py; DESIGN.md section 4 explains what "exact" can and cannot mean here)
        from oracle.batched_ref import simulate_discrete_batch
        from oracle.parity import full_horizon_report
        pl = args.parity_lanes if prob.Nx <= 10 else max(16, args.parity_lanes // 2)
        wlp = wl if kind == "D" else dict(wl, kind="D")
        x0p, nzp = make_inputs(wlp, pl, SEED0 - 1)
        x0T = np.ascontiguousarray(x0p.T)
        psc = sc if kind == "D" else plan_sc(sc)
        pprob = prob if kind == "D" else M.build_problem(psc, mp, fp, None)
        with M.Engine(pprob, device=local) as pe:
            got = M.trajectorySimulateBatch(psc, mp, fp, None, x0T, nzp, engine=pe)
        ref = simulate_discrete_batch(psc, mp, fp, x0T, nzp, chol_fail="clamp", spectral=(pprob.V, pprob.lam))
        rep = full_horizon_report(got, ref)
        parity = {"lanes": rep["lanes"], "steps": rep["steps"], "exact_frac": rep["exact_frac"],
                  "exact_strict_frac": rep["exact_strict_frac"], "max_du": rep["max_du_prefix"],
                  "max_du_decision_prefix": rep["max_du_decision_prefix"],
                  "solves_on_exact_prefix": rep["solves_exact_prefix"], "solves_on_strict_prefix": rep["solves_strict_prefix"],
                  "solves_compared": rep["solves_compared"],
                  "checker": "oracle/batched_ref.py fed the engine's spectral tables, all 300 steps; exact = iterations, status, "
                             "controller of every solve and rho to 1 % (adaptation decisions) up to i_term; strict = rho to 1e-6 too; "
                             "max_du over the strict prefixes (task bar 1e-4), max_du_decision_prefix over the exact ones"}

    red = torch.tensor([ms, float(solves), ms_e2e, float(solves_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = red.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, ms_e2e = float(mx[0]), float(mx[2])
        solves, solves_e2e = int(sm[1]), int(sm[3])

    if rank == 0:
        n, m = prob.n, prob.m
        nnzA = int(np.count_nonzero(prob.A))
        fit = f_it(n, m, nnzA)
        admm_ms = delta(t0, t1, "admm_ms")
        admm_launches = delta(t0, t1, "admm_launches")
        admm_iters = delta(t0, t1, "admm_iterations")
        step_ms_timed = delta(t0, t1, "total_ms")
        step_solves = delta(t0, t1, "qp_solves")
        peak_dfma, peak_dmma = np.zeros(1), np.zeros(1)
        lib = _lib.load()
        _lib.check(lib.mpcb_measure_fp64_peak(local, 0, peak_dfma.ctypes.data_as(_lib.c_double_p)))
        _lib.check(lib.mpcb_measure_fp64_peak(local, 1, peak_dmma.ctypes.data_as(_lib.c_double_p)))
        peak = float(max(peak_dfma[0], peak_dmma[0]))
        achieved = fit * admm_iters / (admm_ms * 1e-3) / 1e12 if admm_ms > 0 else None
        forced = os.environ.get("MPCB_SOLVER", "")
        blocks = eng.solver_blocks()
        Bcell = B if plan is None else Bc
        wave_min = int(os.environ.get("MPCB_WAVE_MIN_LANES", "16384"))
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        ctas4 = (Bcell + 31) // 32                 # mpcb.cu wave_small_window: every SM (or at least 3 of 4) gets one 4-warp CTA
        wave4 = (not os.environ.get("MPCB_NO_WAVE4")) and 4 * ctas4 >= 3 * sms and ctas4 <= sms and prob.n == 81
        if forced in ("block", "tile", "wave"):
            kern = {"block": "admm_block_kernel", "tile": "admm_tile_kernel", "wave": "admm_wave_kernel + team_kernel (resume)"}[forced]
        elif prob.has_debris:
            kern = "generic_lane_kernel"
        elif not blocks["team"]:
            kern = "admm_block_kernel"
        elif forced != "team" and blocks["wave"] and kind != "C" and (Bcell >= wave_min or wave4):
            kern = ("admm_wave_kernel<%d warps> (multi-RHS DMMA rounds) + team_kernel (takes the last lanes over mid-flight)"
                    % (4 if (Bcell + 31) // 32 <= sms else 8))
        else:
            kern = "team_kernel" + (" (whole closed loop, one launch per step)" if kind != "C" else " (list mode: one launch per round of solves)")
        traffic = None
        try:        # DRAM bytes per launch of the dominant kernel from the committed ncu capture (null if never captured)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(f"{kern.split(' ')[0]}:{args.workload}")
            w4, ho = tj.get(f"admm_wave_kernel<4>:{args.workload}"), tj.get(f"team_kernel_handover:{args.workload}")
            if kern.startswith("admm_wave_kernel<4") and w4 and ho and admm_launches > 1:
                # rounds + one hand-over launch: launch-weighted mean of the two kernels' captured bytes
                traffic = int((w4 * (admm_launches - 1) + ho) / admm_launches)
        except Exception:
            pass
        tot_stats = np.asarray(res["stats"]).reshape(-1, _lib.NSTATS).sum(axis=0) if plan is not None else res.stats_vec
        line = {
            "metric": METRIC, "value": solves / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "lanes_per_gpu": B, "Nx": prob.Nx, "n": n, "m": m,
                       "seeds": f"{SEED0} + step * {world} + rank: every step (warm-up, timed, e2e) simulates newly drawn lanes",
                       "l2": "flushed between steps (256 MiB write)", "telemetry": list(record) or ["statistics only"],
                       "solver": "OSQP-equivalent ADMM, reference defaults (eps 1e-3, adaptive rho, check every 25)",
                       "mean_admm_iters_per_solve": float(tot_stats[6] / max(1.0, tot_stats[5])),
                       "live_steps_per_lane": float(tot_stats[5] / B),
                       "flip_lanes": float(tot_stats[7]), "ukf_clamped_lanes": float(tot_stats[8]),
                       "operator_rebuilds_per_step": delta(c0, c1, "operator_rebuilds") / max(1, args.steps),
                       "solver_mix": mix},
            "e2e": {"value": solves_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(delta(c0, c1, "kernel_launches")),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "kernel": kern,
                         "algorithmic_flops_per_iteration": fit, "iterations_in_step": int(admm_iters),
                         "launches": int(admm_launches), "avg_launch_ms": admm_ms / max(1, admm_launches),
                         "share_of_step": admm_ms / step_ms_timed if step_ms_timed else None,
                         "step_solves_per_s": step_solves / (step_ms_timed * 1e-3) if step_ms_timed else None,
                         "peak_source": "measured live on this GPU: float64 DFMA %.1f / DMMA m8n8k4 %.1f TFLOP/s "
                                        "(MEASURED_PEAKS.json has no float64 entry)" % (peak_dfma[0], peak_dmma[0])},
        }
        if plan is not None:
            line["config"]["dist_ratio"] = [float(v) for v in res["dist_ratios"]]
            line["config"]["noise_lengths"] = list(DISTURB_REJ_LENGTHS)
            line["config"]["realisations_per_cell"] = Bc
        if parity is not None:
            line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            if kind == "C":
                lanes = 2 * cores
            else:       # a pilot sizes the sample to 10-20 s of wall time on all cores (the families differ 40x in cost per solve)
                pilot = (2 if kind == "D" else 1) * cores * (20 if kind == "S" else 1)
                s0, w0 = cpu_twin_run(wl, pilot, cores, SEED0 - 7)
                lanes = int(min(65536, max(pilot, pilot * 24.0 / max(w0, 1e-3))))
            s, w, ckind, desc = cpu_arm(wl, lanes, cores, SEED0)
            cb = {"value": s / w, "unit": UNIT, "cores": cores, "kind": ckind,
                  "sample": f"{lanes} lanes of {args.workload}, full horizon, {w:.1f} s wall: {desc}"}
            if kind != "C":
                s1, w1 = cpu_twin_run(wl, max(2 if kind == "D" else 20, lanes // (3 * cores)), 1, SEED0)
                cb["one_core"] = s1 / w1
            if kind == "D" and not prob.has_debris and prob.Nx <= 20:
                # the restated PYTHON path (numpy + restated OSQP / UKF, structured like the reference's loop),
                # one trajectory per core -- what the reference's interpreter-bound loop costs, beside the compiled twin
                s2, w2 = cpu_python_run(wl, cores, cores, SEED0)
                cb["python_restatement"] = {"value": s2 / w2, "cores": cores, "sample": f"{cores} lanes, full horizon, {w2:.1f} s wall"}
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    for e in engines:
        e.close()
    if world > 1:
        dist.destroy_process_group()


def plan_sc(sc):
    """The sweep's parity slice runs one cell: hold 50, with rejection."""
    import copy
    from mpc_arpo_project_b200.mpcsim import Noise
    s2 = copy.copy(sc)
    s2.isReject = True
    s2.noise = Noise(tuple(sc.noise.noise_std), 50)
    return s2


if __name__ == "__main__":
    main()
