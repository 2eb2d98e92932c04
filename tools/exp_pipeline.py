"""Does a second batch fill the tail of the first?  depth 1 / 2 / 3 on BASELINE config 2 (and a large batch)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_arpo_project_b200 as M
from mpc_arpo_project_b200.pipeline import BatchPipeline
from mpc_arpo_project_b200.presets import WORKLOADS, make_inputs, make_params

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else WORKLOADS[name]["lanes"]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
wl = WORKLOADS[name]
sc, mp, fp, _ = make_params(wl["case"])
prob = M.build_problem(sc, mp, fp, None)
dev = torch.device("cuda:0")
ins = []
for s in range(K + 3):
    x0, nz = make_inputs(wl, B, 1234 + s)
    ins.append((torch.from_numpy(x0).to(dev), torch.from_numpy(nz).to(dev) if nz is not None else None))
rec = ("x_true", "x_est", "ctrl", "ctrlr_seq")
for depth in (1, 2, 3, 2, 1):
    with BatchPipeline(prob, 0, depth, lanes=B) as pipe:
        def one(eng, x0, nz):
            t0 = time.perf_counter()
            r = eng.simulate_discrete(x0, nz, 300, rec)
            return int(r.stats["qp_solves"]), t0, time.perf_counter()      # telemetry tensors die here: the allocator reuses them
        [f.result() for f in [pipe.submit(one, *i) for i in ins[:3]]]
        [f.result() for f in [pipe.submit(one, *i) for i in ins[:3]]]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = [f.result() for f in [pipe.submit(one, *i) for i in ins[3:]]]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        solves = sum(r[0] for r in res)
        if os.environ.get("PIPE_TIMELINE"):
            print("  calls (start, end) ms:", " ".join(f"({1e3 * (a - t0):.0f},{1e3 * (b - t0):.0f})" for _, a, b in res))
        print(f"{name} B={B} depth={depth}: {1e3 * dt / K:.2f} ms per step, {solves / dt / 1e6:.3f} M solves/s", flush=True)
        del res
