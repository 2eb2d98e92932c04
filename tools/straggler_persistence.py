"""Is a lane's solver cost persistent over its trajectory?  (Would a probe of the first K control steps find the
stragglers of config 2?)  Prints, per seed, where the heaviest lanes rank when lanes are ordered by the iterations they
used in their first K steps, and the makespan of list scheduling on 296 teams for index order vs probe order."""
import sys, heapq, numpy as np
sys.path.insert(0, '/root/repo')
import mpc_arpo_project_b200 as M
import bench
from oracle.gen_golden import make_params
wl = bench.WORKLOADS["config2"]
sc, mp, fp, _ = make_params(M, wl["case"])
eng = M.Engine(M.build_problem(sc, mp, fp, None))


def makespan(work, order, P=296):
    h = [0.0] * P
    heapq.heapify(h)
    for ln in order:
        t = heapq.heappop(h)
        heapq.heappush(h, t + work[ln])
    return max(h)


for seed in (1234, 1235, 1236, 1237, 1238, 1239, 1240, 1241):
    x0, noise = bench.make_inputs(wl, 4096, seed)
    r = eng.simulate_discrete(x0, noise, 300, ("iters",))
    its = r.iters.astype(np.int64)
    tot = its.sum(0)
    top = np.argsort(-tot)[:5]
    line = f"seed {seed}: max lane {tot.max():6d} mean {tot.mean():.0f} | index-order makespan {makespan(tot, range(4096)) / 1e3:6.1f}k ideal {max(tot.max(), tot.sum() / 296) / 1e3:6.1f}k"
    for K in (10, 25, 50):
        head = its[:K].sum(0)
        rest = tot - head
        order = np.argsort(-head)
        rk = np.argsort(np.argsort(-head))[top]
        line += f" | K={K}: top5 ranks {rk.tolist()} makespan {(head.sum() / 296 + makespan(rest, order)) / 1e3:6.1f}k"
    print(line)
