import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mpc_arpo_project_b200 as M
from oracle.gen_golden import make_params
from oracle.batched_ref import simulate_discrete_batch
from test_batched_ref import retype_lanes
Nx = int(sys.argv[1]) if len(sys.argv) > 1 else 10
case, x0, noise = retype_lanes(B=24, seed=11, Nx=Nx)
sc, mp, fp, _ = make_params(M, case)
got = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise)
ref = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp')
off = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp', retype=False)
gi = np.asarray(got.iters).astype(int)
for b in range(x0.shape[0]):
    T = min(int(got.i_term[b]), int(ref["i_term"][b]))
    d = np.nonzero(gi[:T, b] != ref["iters"][:T, b])[0]
    d0 = np.nonzero(gi[:T, b] != off["iters"][:T, b])[0]
    print(b, "flip" if ref["flip_flag"][b] else "    ", "i_term eng/ref/off", int(got.i_term[b]), int(ref["i_term"][b]), int(off["i_term"][b]),
          "first iters diff vs ref", (int(d[0]), gi[d[0], b], ref["iters"][d[0], b]) if d.size else None,
          "| vs off", (int(d0[0]), gi[d0[0], b], off["iters"][d0[0], b]) if d0.size else None)
