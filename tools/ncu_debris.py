"""One small batch of config 1's lanes (Nx = 40 + debris) for an ncu capture of generic_lane_kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_arpo_project_b200 as M
from mpc_arpo_project_b200.presets import WORKLOADS, make_inputs, make_params
wl = WORKLOADS["config1"]
sc, mp, fp, debris = make_params(dict(wl["case"], T_final=20))
B = 148
x0, noise = make_inputs(wl, B, 1234)
eng = M.Engine(M.build_problem(sc, mp, fp, debris))
for _ in range(2):
    r = M.trajectorySimulateBatch(sc, mp, fp, debris, np.ascontiguousarray(x0.T), noise, engine=eng, record=(), nsteps=40)
print(r.stats)
