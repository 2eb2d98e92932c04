"""Drop-in for ``src/trajectorySimulate.py:17-388`` (discrete-time linear CW plant) plus the
batched entry point the B200 engine exists for.

``trajectorySimulate(sim_conditions, mpc_params, fail_params, debris) -> SimRun`` keeps the
reference signature and result shapes (one trajectory; the legacy global numpy RNG is re-seeded
with 123 exactly like ``:28``).  ``trajectorySimulateBatch`` runs B trajectories in lockstep on
the GPU and returns a :class:`BatchSimRun`.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .engine import Engine, _RECORD_ALL
from .mpcsim import BatchSimRun, SimRun
from .problem import Problem, SolverSettings, build_problem


def n_control_steps(sim_conditions) -> int:
    """``nsimD = int(T_final / T)`` (ref ``:54``)."""
    return int(sim_conditions.T_final / sim_conditions.time_stp)


def noise_refreshes(nsteps: int, noise_length: int) -> int:
    """Disturbance draws a run of ``nsteps`` consumes: one before the loop (``:268``) and one after
    every ``noise_length``-th step (``:351-353``)."""
    return nsteps // max(1, int(noise_length)) + 1


def trajectorySimulateBatch(sim_conditions, mpc_params, fail_params, debris, x0_batch, noise_batch=None,
                            seed: Optional[int] = None, nsteps: Optional[int] = None,
                            record: Sequence[str] = _RECORD_ALL, settings: Optional[SolverSettings] = None,
                            device: int = 0, engine: Optional[Engine] = None, noise_rng: str = "numpy",
                            lane_offset: int = 0) -> BatchSimRun:
    """B trajectories of ``trajectorySimulate`` on one GPU.

    ``x0_batch[B, 4]`` initial states (``sim_conditions.x0`` is ignored); ``noise_batch[R, 2, B]``
    sigma-scaled position disturbances, row ``r`` held from control step ``r*noise_length`` (the
    reference draws ``sigMat @ N(0,1)^4`` there, ``:268,352``); if omitted and the conditions carry a
    ``Noise``, they are drawn from ``numpy.random.default_rng(seed)`` or, with ``noise_rng="philox"``, on the GPU
    (``Engine.noise_fill``: Philox4x32-10 keyed by ``seed``, lane ``lane_offset + b``; nothing crosses PCIe when ``x0_batch``
    is a device tensor).  Arrays may be numpy or torch
    CUDA tensors in the engine's SoA layout: pass ``x0_batch`` as ``[4, B]`` torch tensors to skip
    every host copy.
    """
    eng = engine or Engine(build_problem(sim_conditions, mpc_params, fail_params, debris, settings), device)
    try:
        p: Problem = eng.problem
        nsteps = n_control_steps(sim_conditions) if nsteps is None else int(nsteps)
        on_dev = type(x0_batch).__module__.startswith("torch")
        if on_dev:
            x0 = x0_batch                      # already [4, B] on the device
            B = x0.shape[1]
        else:
            x0_batch = np.asarray(x0_batch, float)
            B = x0_batch.shape[0]
            x0 = np.ascontiguousarray(x0_batch.T)
        if p.has_noise and noise_batch is None:
            R = noise_refreshes(nsteps, p.noise_length)
            if noise_rng == "philox":
                noise_batch = eng.noise_fill(B, R, 0 if seed is None else seed, lane_offset, on_device=on_dev)
            elif noise_rng == "numpy":
                rng = np.random.default_rng(seed)
                noise_batch = rng.standard_normal((R, 2, B)) * p.sig[None, :, None]
                if on_dev:
                    import torch
                    noise_batch = torch.from_numpy(noise_batch).to(x0.device)
            else:
                raise ValueError(f"unknown noise_rng {noise_rng!r}")
        if not p.has_noise:
            noise_batch = None
        return eng.simulate_discrete(x0, noise_batch, nsteps, record)
    finally:
        if engine is None:
            eng.close()


def trajectorySimulate(sim_conditions, mpc_params, fail_params, debris):
    """Reference signature, one trajectory, ``SimRun`` result (``src/trajectorySimulate.py:17-26``)."""
    np.random.seed(123)                                            # :28
    nsim = n_control_steps(sim_conditions)
    noise = sim_conditions.noise
    nl = int(noise.noise_length) if noise is not None else 1
    R = noise_refreshes(nsim, nl)
    if noise is not None:
        sig = noise.constructSigMat()
        draws = np.stack([sig @ np.random.normal(0, 1, 4) for _ in range(R)])        # :268, :352
    else:
        draws = np.zeros((R, 4))
    nb = np.ascontiguousarray(draws[:, :2, None]) if noise is not None else None
    x0 = np.asarray(sim_conditions.x0, float).reshape(1, 4)
    b = trajectorySimulateBatch(sim_conditions, mpc_params, fail_params, debris, x0, nb, nsteps=nsim)
    it = int(b.i_term[0])
    noise_hist = np.full((4, nsim + 1), np.nan)
    upto = min(it + 1, nsim + 1)
    noise_hist[:, :upto] = draws[np.arange(upto) // nl].T
    return SimRun(it, bool(b.isSuccess[0]), b.x_true[:, :it, 0].copy(), b.x_est[:, :, 0].copy(),
                  b.ctrl_hist[:, :, 0].copy(), b.ctrlr_seq[:it, 0].astype(float), noise_hist)
