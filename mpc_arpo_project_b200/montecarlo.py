"""Monte-Carlo studies of the reference as single batched calls.

* :func:`final_distance_ratio_sweep` is ``test/disturbRejComp.py:74-100``: for every noise hold length, the
  mean final distance to the target with disturbance rejection divided by the mean without it.
* :func:`success_rate` is ``test/saved_runs/success_rates_test.py:64-75``.

The reference re-seeds the global RNG with 123 inside every ``trajectorySimulate`` call
(``src/trajectorySimulate.py:28``), so its "100 Monte-Carlo runs" per setting are 100 copies of one
realisation and its rejection / no-rejection runs share that realisation.  Here every lane gets its own
draw (``numpy.random.default_rng(seed)``), and the rejection / no-rejection batches share the draws lane by
lane, which keeps the pairing and makes the average a real one.  Under ``torch.distributed`` every rank
simulates its own ``mc_num`` lanes and the statistics are all-reduced (sum), the path's only collective.
"""
from __future__ import annotations

import copy
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .engine import Engine
from .mpcsim import Noise
from .problem import build_problem
from .trajectorySimulate import n_control_steps, noise_refreshes


def _allreduce_sum(vec: np.ndarray, device: int) -> np.ndarray:
    try:
        import torch
        import torch.distributed as dist
    except Exception:
        return vec
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vec
    dev = f"cuda:{device}" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).to(dev)
    dist.all_reduce(t)
    return t.cpu().numpy()


def _rank() -> int:
    try:
        import torch.distributed as dist
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    except Exception:
        return 0


def _run(sc, mp, fp, debris, x0, noise, device, philox=None):
    """``philox = (n_refresh, seed, lane_offset)``: draw the disturbances on the GPU (``Engine.noise_fill``) instead of
    taking ``noise`` from the host."""
    eng = Engine(build_problem(sc, mp, fp, debris), device)
    try:
        if philox is not None and sc.noise is not None:
            noise = eng.noise_fill(x0.shape[1], philox[0], philox[1], philox[2])
        return eng.simulate_discrete(x0, noise, n_control_steps(sc), record=())
    finally:
        eng.close()


class RatioSweep:
    """``test/disturbRejComp.py:74-100`` as a prepared plan: one engine per (noise hold length, reject / no reject) cell, built
    once, so that repeated sweeps (Monte-Carlo batches, benchmark steps) pay no set-up.  ``run`` takes the lanes' initial
    states and one disturbance tensor per hold length (shared by the cell's two modes, lane by lane, which keeps the
    reference's pairing of the rejection / no-rejection runs) and returns the ratio curve."""

    def __init__(self, sim_conditions, mpc_params, fail_params, debris, noise_std: Sequence[float], noise_lengths: Sequence[float],
                 device: int = 0, pin_outputs: bool = False, settings=None):
        self.noise_lengths = [int(nl) for nl in noise_lengths]
        self.noise_std = tuple(float(v) for v in noise_std)
        self.nsteps = n_control_steps(sim_conditions)
        self.device = device
        self.engines = []                      # [length][mode]: mode 0 = no rejection, 1 = rejection
        for nl in self.noise_lengths:
            row = []
            for rej in (False, True):
                sc = copy.copy(sim_conditions)
                sc.isReject = rej
                sc.noise = Noise(self.noise_std, nl)
                row.append(Engine(build_problem(sc, mpc_params, fail_params, debris, settings), device, pin_outputs=pin_outputs))
            self.engines.append(row)

    def refreshes(self, i: int) -> int:
        return noise_refreshes(self.nsteps, self.noise_lengths[i])

    def close(self):
        for row in self.engines:
            for e in row:
                e.close()
        self.engines = []

    def run(self, x0, noises, record=()):
        """``x0[4, B]`` and ``noises[i][R_i, 2, B]`` (numpy or torch CUDA, all of one kind) -> dict with ``dist_ratios[len]``,
        the two mean curves, the summed statistics vectors ``stats[len][mode]`` (what ranks all-reduce) and ``qp_solves``."""
        L = len(self.noise_lengths)
        stats = np.zeros((L, 2, _lib.NSTATS))
        for i in range(L):
            for k in range(2):
                stats[i, k] = self.engines[i][k].simulate_discrete(x0, noises[i], self.nsteps, record=record).stats_vec
        return self.reduce(stats)

    def reduce(self, stats: np.ndarray) -> dict:
        """Ratio curve from (all-reduced) statistics: stats[..., 0] = sum of final distances, [..., 3] = lanes."""
        mean = stats[:, :, 0] / np.maximum(stats[:, :, 3], 1.0)
        return {"noise_lengths": np.asarray(self.noise_lengths, float), "mean_dist_norej": mean[:, 0], "mean_dist_rej": mean[:, 1],
                "dist_ratios": mean[:, 1] / mean[:, 0], "stats": stats, "qp_solves": float(stats[:, :, 5].sum())}


def final_distance_ratio_sweep(sim_conditions, mpc_params, fail_params, debris, noise_std: Sequence[float],
                               noise_lengths: Sequence[float], mc_num: int = 100, seed: int = 0, device: int = 0,
                               noise_rng: str = "numpy") -> dict:
    """``dist_ratios[i] = mean_j ||x_final(rej) - xr|| / mean_j ||x_final(no rej) - xr||`` for hold length i
    (``disturbRejComp.py:85-98``).  Returns the ratios and both means; ``mc_num`` lanes per rank and setting.
    ``noise_rng="philox"``: draws come from the device generator (stream = seed + hold-length index, lanes offset by rank)."""
    x0 = np.tile(np.asarray(sim_conditions.x0, float)[:, None], (1, mc_num))
    rng = np.random.default_rng(seed + 7919 * _rank())
    sig = np.asarray(noise_std, float)
    plan = RatioSweep(sim_conditions, mpc_params, fail_params, debris, noise_std, noise_lengths, device)
    try:
        noises = []
        for i in range(len(plan.noise_lengths)):
            R = plan.refreshes(i)
            if noise_rng == "numpy":
                noises.append(rng.standard_normal((R, 2, mc_num)) * sig[None, :, None])
            else:
                noises.append(plan.engines[i][0].noise_fill(mc_num, R, seed + i, _rank() * mc_num))
        out = plan.run(x0, noises)
        stats = _allreduce_sum(out["stats"].reshape(-1), device).reshape(out["stats"].shape)
        return plan.reduce(stats)
    finally:
        plan.close()


def success_rate(sim_conditions, mpc_params, fail_params, debris, mc_num: int = 300, seed: int = 0, device: int = 0,
                 x0_batch: Optional[np.ndarray] = None, noise_rng: str = "numpy") -> dict:
    """Fraction of successful approaches over ``mc_num`` noise realisations (``success_rates_test.py:66-75``)."""
    x0 = np.tile(np.asarray(sim_conditions.x0, float)[:, None], (1, mc_num)) if x0_batch is None else np.ascontiguousarray(x0_batch.T)
    nsteps = n_control_steps(sim_conditions)
    noise, philox = None, None
    if sim_conditions.noise is not None:
        R = noise_refreshes(nsteps, int(sim_conditions.noise.noise_length))
        if noise_rng == "numpy":
            rng = np.random.default_rng(seed + 7919 * _rank())
            noise = rng.standard_normal((R, 2, x0.shape[1])) * np.asarray(sim_conditions.noise.noise_std, float)[None, :, None]
        else:
            philox = (R, seed, _rank() * x0.shape[1])
    run = _run(sim_conditions, mpc_params, fail_params, debris, x0, noise, device, philox)
    s = _allreduce_sum(np.array([run.stats["n_success"], run.stats["n_lanes"]]), device)
    return {"success_count": int(s[0]), "runs": int(s[1]), "success_rate": s[0] / s[1]}
