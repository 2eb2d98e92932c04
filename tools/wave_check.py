"""wave kernel vs oracle: QP seam (cold + 2 warm), then a short closed loop."""
import os, sys
os.environ["MPCB_SOLVER"] = "wave"
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/tests") else ".")
import numpy as np
import mpc_arpo_project_b200 as M
from mpc_arpo_project_b200.presets import make_params
from oracle.batched_ref import BatchedQP, simulate_discrete_batch
from oracle.sim_ref import build_setup
B = int(sys.argv[1]) if len(sys.argv) > 1 else 77
sc, mp, fp, _ = make_params(dict(Nx=10, sigma=0.1))
rng = np.random.default_rng(10)
s = build_setup(sc, mp, fp, None)
prob = M.build_problem(sc, mp, fp, None)
qp = BatchedQP(s, B, spectral=(prob.V, prob.lam))
eng = M.Engine(prob)
print("solver blocks:", eng.solver_blocks())
eng.batch_alloc(B)
xh = np.zeros((B, 6)); xh[:, 0] = 100 + rng.uniform(-10, 10, B); xh[:, 1] = 10 + rng.uniform(-5, 5, B)
for rnd in range(3):
    if rnd:
        xh[:, :2] += rng.normal(0, 0.05, (B, 2)); xh[:, 2:4] = rng.normal(0, 0.1, (B, 2)); xh[:, 4:6] = rng.normal(0, 0.01, (B, 2))
    val = np.abs(xh[:, 0] - s.xr[0]) + np.abs(xh[:, 1] - s.xr[1]); var = (xh[:, 2] < 0).astype(int) + 2 * (xh[:, 3] < 0).astype(int)
    idx = np.arange(B)
    qp.set_params(idx, xh[:, :4], val, xh[:, 4:6], var)
    st_ref, it_ref = qp.solve(idx)
    u_ref = qp.x[:, 44:46] * qp.D[44:46][None, :]
    u0, st, it = eng.qp_solve(np.ascontiguousarray(xh.T))
    print(f"round {rnd}: iters equal {np.array_equal(it, it_ref)} ({(it != it_ref).sum()} differ) status equal {np.array_equal(st, st_ref)} max|du| {np.abs(u0.T - u_ref).max():.3e}")
    if not np.array_equal(it, it_ref):
        bad = np.nonzero(it != it_ref)[0][:8]
        print("   lanes", bad, "got", it[bad], st[bad], "ref", it_ref[bad], st_ref[bad], "variants", var[bad])
    x, z, y, rho = eng.qp_state(B // 2)
    print(f"   lane {B//2}: |dx| {np.abs(x - qp.x[B//2]).max():.2e} |dz| {np.abs(z - qp.z[B//2]).max():.2e} |dy| {np.abs(y - qp.y[B//2]).max():.2e} rho {rho} {qp.rho[B//2]}")
eng.close()
# closed loop
case = dict(Nx=10, sigma=0.75, noise_length=50, T_final=25)
sc, mp, fp, _ = make_params(case)
B2 = 96
x0 = np.array([100., 10., 0., 0.])[None, :] + np.concatenate([rng.uniform(-5, 5, (B2, 2)), np.zeros((B2, 2))], axis=1)
noise = 0.75 * rng.standard_normal((2, 2, B2))
got = M.trajectorySimulateBatch(sc, mp, fp, None, x0, noise)
prob = M.build_problem(sc, mp, fp, None)
ref = simulate_discrete_batch(sc, mp, fp, x0, noise, chol_fail='clamp', spectral=(prob.V, prob.lam))
print("closed loop: iters equal", np.array_equal(got.iters.astype(int), ref['iters']), "status", np.array_equal(got.status.astype(int), ref['status']),
      "max|du|", np.nanmax(np.abs(got.ctrl_hist.transpose(1, 2, 0) - ref['ctrl_hist'])), "i_term", np.array_equal(got.i_term, ref['i_term']))
