"""Full-horizon GPU parity: every benchmarked workload for ALL 300 control steps against the oracle, lane by lane.

What can and cannot be asserted (measured with tools/parity_report.py):

* The closed loop of the reference is chaotic at the level of OSQP's discrete decisions (terminate at this 25-iteration
  check or the next, adapt rho or not): the R = 0 UKF and the 1e-3 termination tolerances amplify a 1e-13 perturbation of
  x0 into a different decision on ~13 % of config 2's lanes within 300 steps -- measured on the ORACLE AGAINST ITSELF (the
  control below).  No two float64 implementations can do better than that floor, and the two CPU oracles (numpy
  eigen-decomposition vs the C twin's Cholesky) agree with each other on 71 % of config 2's lanes
  (tests/test_oracle_c.py::test_c_twin_full_horizon_against_the_batched_oracle, gpurun_out r2b_parity_rho_probe.json).
* "Exact" = iteration count, OSQP status, controller choice and rho (to 1 %: did both sides adapt, oracle/parity.py) of
  EVERY solve up to i_term, and i_term itself.  Measured (seed 4321; engine's tables / oracle's own tables):
  config 2 82 % / 63 %, config 2 quiet 99 % / 96 %, config 4 99 % / 97 %, config 5 cell 83 % / 73 %.  Engine vs oracle with
  the SAME spectral tables isolates the device arithmetic (it stays exact at the control's rate); the oracle's own tables
  add the 1e-11..1e-9 by which two `eigh` calls reconstruct M(rho)^-1.
* Controls are compared on the STRICT prefix of a lane: every decision equal and rho equal to 1e-6.  There they agree to
  1e-6 (measured 3e-8), far inside the task's 1e-4 bar.  Past the first solve whose adapted rho differs in the 5th digit
  the two sides run different OSQP solves -- same decisions, iterates apart by a fraction of OSQP's own 1e-3 termination
  tolerance (measured up to 6e-4 in the controls) -- which no implementation other than OSQP's own arithmetic can avoid.
"""
import numpy as np
import pytest

import mpc_arpo_project_b200 as M
from mpc_arpo_project_b200.presets import WORKLOADS, make_inputs, make_params
from oracle.batched_ref import simulate_discrete_batch
from oracle.parity import as_engine_layout, full_horizon_report
from oracle.sim_ref import trajectory_simulate_c

pytestmark = pytest.mark.gpu

U_BAR = 1e-4          # BASELINE.json north_star: per-step controls within 1e-4 abs
# name: (lanes, floor on exact lanes with the engine's tables, floor with the oracle's own tables, bound on |du| over the exact prefix)
CASES = {
    "config2": (256, 0.72, 0.53, 1e-6),
    "config2_quiet": (256, 0.95, 0.90, 1e-6),
    "config4": (192, 0.95, 0.90, 1e-6),
    "config5_cell": (128, 0.72, 0.62, 1e-6),
    # the large-batch path: rounds of the multi-RHS wave kernel (FP64 tensor cores, unscaled variables), forced at 256 lanes;
    # re-typing lanes and -- in the second case -- the last lanes go to the team kernel as they do in a 65 536-lane run
    "config2@wave_rounds_only": (256, 0.72, 0.53, 1e-6),          # measured 81 % / 63 %: the team path's figures
    "config2_quiet@wave_then_team": (256, 0.95, 0.90, 1e-6),      # measured 99 % / 96 %
}
ENV = {"config2@wave_rounds_only": {"MPCB_SOLVER": "wave", "MPCB_RESUME_BELOW": "0"},
       "config2_quiet@wave_then_team": {"MPCB_SOLVER": "wave", "MPCB_RESUME_BELOW": "100"}}


def _run(name, B, seed=4321):
    name = name.split("@")[0]
    wl = WORKLOADS[name]
    sc, mp, fp, _ = make_params(wl["case"])
    x0, noise = make_inputs(wl, B, seed)
    x0T = np.ascontiguousarray(x0.T)
    prob = M.build_problem(sc, mp, fp, None)
    with M.Engine(prob) as eng:
        got = M.trajectorySimulateBatch(sc, mp, fp, None, x0T, noise, engine=eng)
    return (sc, mp, fp), prob, x0, x0T, noise, got


@pytest.mark.parametrize("name", list(CASES))
def test_full_horizon_against_the_batched_oracle(name, monkeypatch):
    B, floor_shared, floor_own, du_bound = CASES[name]
    for k, v in ENV.get(name, {}).items():
        monkeypatch.setenv(k, v)
    (sc, mp, fp), prob, x0, x0T, noise, got = _run(name, B)
    nsim = int(sc.T_final / sc.time_stp)
    assert nsim == 300 and got.iters.shape == (nsim, B)
    shared = simulate_discrete_batch(sc, mp, fp, x0T, noise, chol_fail='clamp', spectral=(prob.V, prob.lam))
    rep_s = full_horizon_report(got, shared, U_BAR)
    own = simulate_discrete_batch(sc, mp, fp, x0T, noise, chol_fail='clamp')
    rep_o = full_horizon_report(got, own, U_BAR)
    print(name, "engine tables:", rep_s["exact_lanes"], "/", B, "max du", rep_s["max_du_prefix"],
          "| oracle tables:", rep_o["exact_lanes"], "/", B, "max du", rep_o["max_du_prefix"])
    # (1) discrete record identical to i_term on at least the measured floor of lanes
    assert rep_s["exact_frac"] >= floor_shared, rep_s
    assert rep_o["exact_frac"] >= floor_own, rep_o
    # (2) on the prefix where every decision (iterations, status, controller, rho) still coincides the controls agree
    assert rep_s["max_du_prefix"] <= du_bound <= U_BAR, rep_s
    assert rep_o["max_du_prefix"] <= U_BAR, rep_o
    # (3) most solves are on an exact prefix, and the batch statistic the Monte-Carlo drivers reduce (mean final distance,
    #     disturbRejComp.py:87-100) agrees: diverged lanes are the same controller on a perturbed path
    assert rep_s["solves_exact_prefix"] >= 0.80 * rep_s["solves_compared"], rep_s
    fin_o = np.array([own["x_true"][max(int(t) - 1, 0), b, :2] for b, t in enumerate(own["i_term"])])
    fin_g = np.array([got.x_true[:2, max(int(t) - 1, 0), b] for b, t in enumerate(got.i_term)])
    mo, mg = np.linalg.norm(fin_o, axis=1).mean(), np.linalg.norm(fin_g, axis=1).mean()
    assert abs(mo - mg) <= 0.05 * max(mo, 1.0), (mo, mg)
    assert (np.asarray(got.i_term) == own["i_term"]).mean() >= 0.9


def test_engine_is_as_faithful_as_the_oracle_is_to_itself():
    """The control: perturb x0 by 1e-13 (relative) and run the ORACLE twice.  The engine, fed the oracle's tables, keeps at
    least as many lanes exact as the oracle keeps against its own perturbed copy (less a sampling margin)."""
    name, B = "config2", 256
    (sc, mp, fp), prob, x0, x0T, noise, got = _run(name, B)
    ref = simulate_discrete_batch(sc, mp, fp, x0T, noise, chol_fail='clamp', spectral=(prob.V, prob.lam))
    rng = np.random.default_rng(99)
    x0p = np.ascontiguousarray((x0 * (1 + 1e-13 * rng.standard_normal(x0.shape))).T)
    refp = simulate_discrete_batch(sc, mp, fp, x0p, noise, chol_fail='clamp', spectral=(prob.V, prob.lam))
    control = full_horizon_report(as_engine_layout(refp), ref, U_BAR)["exact_frac"]
    rep = full_horizon_report(got, ref, U_BAR)
    print("control (oracle vs oracle, x0 * (1 + 1e-13)):", control, "engine vs oracle:", rep["exact_frac"])
    assert control < 0.999, "the control found no sensitivity: the floor argument of DESIGN.md section 4 would not hold"
    assert rep["exact_frac"] >= control - 0.10, (rep["exact_frac"], control)


def test_full_horizon_continuous_against_the_scalar_oracle():
    """BASELINE config 3's lanes (nonlinear plant, RK4 at 1 ms, T_final = 150 s) against oracle/sim_ref.trajectory_simulate_c:
    every solve of every lane until the lane terminates."""
    wl = WORKLOADS["config3"]
    sc, mp, fp, _ = make_params(wl["case"])
    B = 8
    x0, noise = make_inputs(wl, B, 4321)
    got = M.trajectorySimulateCBatch(sc, mp, fp, None, np.ascontiguousarray(x0.T), noise)
    exact, solves = 0, 0
    for b in range(B):
        sc.x0 = x0[:, b].copy()
        r = trajectory_simulate_c(sc, mp, fp, None, V=noise[:, :, b].T, integrator='rk4', chol_fail='clamp')
        ns = len(r.iters)
        solves += ns
        gi, gs = np.asarray(got.iters[:ns, b], int), np.asarray(got.status[:ns, b], int)
        bad = np.nonzero((gi != np.asarray(r.iters, int)) | (gs != np.asarray(r.status_val, int)))[0]
        f = int(bad[0]) if bad.size else ns
        exact += int(not bad.size and int(got.i_term[b]) == int(r.i_term))
        assert f >= 1
        np.testing.assert_allclose(got.u_raw[:, :f, b], r.u_raw[:, :f], rtol=0, atol=1e-6)
        for j, i_sub in enumerate(r.solve_at[:f]):
            np.testing.assert_allclose(got.x_true[:, j + 1, b], r.x_true[:, i_sub + 1], rtol=1e-7, atol=1e-6)
    assert solves >= 8 * 60
    assert exact >= B - 3, exact
