"""Host-side logic that needs no GPU: the Monte-Carlo plan's reduction, the noise-refresh bookkeeping, the workload
generator of bench.py / the tests, and the engine's refusal to run without a device."""
import numpy as np
import pytest

from mpc_arpo_project_b200 import _lib
from mpc_arpo_project_b200.montecarlo import RatioSweep, noise_refreshes
from mpc_arpo_project_b200.presets import DISTURB_REJ_LENGTHS, WORKLOADS, make_inputs, make_params


def test_noise_refreshes_cover_the_run():
    """trajectorySimulate.py:351-356 redraws the disturbance every noise_length control steps, starting with a draw before
    step 0 (:268): a run of nsteps steps consumes nsteps // noise_length + 1 rows."""
    for nsteps, nl in ((300, 50), (300, 1), (300, 250), (299, 50), (12, 5)):
        R = noise_refreshes(nsteps, nl)
        assert R == nsteps // nl + 1
        assert (nsteps - 1) // nl < R           # the last step's row exists


def test_ratio_sweep_reduce_is_the_scripts_ratio():
    """disturbRejComp.py:87-100: dist_ratio[len] = mean final distance with rejection / without, per hold length; the plan
    reduces summed statistics (what ranks all-reduce), so the ratio must not depend on how lanes were sharded."""
    plan = RatioSweep.__new__(RatioSweep)        # reduce() needs only the hold lengths
    plan.noise_lengths = [1, 50, 250]
    rng = np.random.default_rng(0)
    d = rng.uniform(1, 20, (3, 2, 64))           # per-lane final distances [length][mode][lane]
    solves = rng.integers(50, 300, (3, 2, 64))

    def stats_of(sl):
        st = np.zeros((3, 2, _lib.NSTATS))
        st[:, :, 0] = d[:, :, sl].sum(-1)
        st[:, :, 3] = d[:, :, sl].shape[-1]
        st[:, :, 5] = solves[:, :, sl].sum(-1)
        return st

    whole = plan.reduce(stats_of(slice(None)))
    sharded = plan.reduce(stats_of(slice(0, 20)) + stats_of(slice(20, 64)))      # two ranks, summed
    want = d[:, 1].mean(-1) / d[:, 0].mean(-1)
    np.testing.assert_allclose(whole["dist_ratios"], want, rtol=1e-13)
    np.testing.assert_allclose(sharded["dist_ratios"], want, rtol=1e-13)
    assert whole["qp_solves"] == float(solves.sum())


def test_workloads_are_the_baseline_configs():
    assert WORKLOADS["config2"]["lanes"] == 4096 and WORKLOADS["config2"]["case"]["Nx"] == 10
    assert WORKLOADS["config3"]["lanes"] == 65536 and WORKLOADS["config3"]["kind"] == "C"
    assert WORKLOADS["config4"]["lanes"] * 8 == 262144 and WORKLOADS["config4"]["case"]["Nx"] == 20
    assert WORKLOADS["config5"]["lanes"] * 8 == 1048576 and WORKLOADS["config5"]["case"]["Nx"] == 30
    assert WORKLOADS["config1"]["case"]["debris"] == ((40., 0.), 5., 20) and WORKLOADS["config1"]["case"]["Nx"] == 40
    assert DISTURB_REJ_LENGTHS == (1, 10, 20, 30, 50, 70, 100, 150, 200, 250)      # disturbRejComp.py:75-82


def test_make_inputs_is_seeded_and_shaped():
    wl = WORKLOADS["config2"]
    x0, nz = make_inputs(wl, 33, 7)
    x0b, nzb = make_inputs(wl, 33, 7)
    assert x0.shape == (4, 33) and nz.shape == (7, 2, 33)          # 300 steps held 50: 6 refreshes + the initial draw
    assert np.array_equal(x0, x0b) and np.array_equal(nz, nzb)
    assert not np.array_equal(x0, make_inputs(wl, 33, 8)[0])
    assert np.all(np.abs(x0[0] - 100) <= 10) and np.all(np.abs(x0[1] - 10) <= 5) and np.all(x0[2:] == 0)
    sc, mp, fp, debris = make_params(WORKLOADS["config1"]["case"])
    assert debris is not None and sc.noise.noise_length == 50
