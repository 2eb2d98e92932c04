"""Drop-in for ``src/trajectorySimulateC.py:17-446`` (continuous-time nonlinear plant) plus the
batched entry point.

The plant is the nonlinear planar relative-motion ODE of ``:64-79``; ``scipy.integrate.solve_ivp``
per ``T_cont`` substep (``:373-380``) is replaced by fixed-step RK4 at ``h = T_cont`` (equal to
RK45's result to ~1e-13 on these dynamics, ``tests/test_oracle_golden.py::test_rk4_equals_rk45``).
The controller runs every ``int(T/T_cont)`` substeps with sample-and-hold (``:335-369``).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .engine import Engine, _RECORD_ALL
from .mpcsim import BatchSimRun, SimRun
from .problem import Problem, SolverSettings, build_problem


def continuous_grid(sim_conditions):
    """``(nsimD, nsimC, ratio)`` of ``trajectorySimulateC.py:55-61``."""
    T, Tc, Tf = sim_conditions.time_stp, sim_conditions.T_cont, sim_conditions.T_final
    return int(Tf / T), int(Tf / Tc), int(T / Tc)


def noise_plan(sim_conditions):
    """``(n_refresh, hold_substeps)``: ``ct.white_noise`` samples at ``arange(0, T_final, T*noise_length)``,
    each held ``int(noise_length*T/T_cont)`` substeps (``:296-307``)."""
    noise = sim_conditions.noise
    if noise is None:
        return 0, 1
    T, Tc, Tf = sim_conditions.time_stp, sim_conditions.T_cont, sim_conditions.T_final
    n_refresh = np.arange(0, Tf, T * noise.noise_length).size
    return int(n_refresh), int((noise.noise_length * T) / Tc)


def build_problem_c(sim_conditions, mpc_params, fail_params, debris=None, settings: Optional[SolverSettings] = None) -> Problem:
    """Same tables as the discrete simulator except the UKF disturbance process noise, which the
    continuous simulator scales by ``T*int(T/T_cont)`` (``:310``) and builds from sigma_x for both
    axes only in the plant noise (``:296``), not in ``Qw`` (``:311``)."""
    _, _, ratio = continuous_grid(sim_conditions)
    return build_problem(sim_conditions, mpc_params, fail_params, debris, settings,
                         ukf_interval_scale=sim_conditions.time_stp * ratio)


def trajectorySimulateCBatch(sim_conditions, mpc_params, fail_params, debris, x0_batch, noise_batch=None,
                             seed: Optional[int] = None, record: Sequence[str] = _RECORD_ALL,
                             settings: Optional[SolverSettings] = None, device: int = 0,
                             engine: Optional[Engine] = None) -> BatchSimRun:
    """B trajectories of ``trajectorySimulateC`` on one GPU.

    ``noise_batch[R, 2, B]``: additive per-substep position disturbance, row ``r`` held for
    ``hold`` substeps (see :func:`noise_plan`); drawn as ``sigma_x * N(0,1)`` for both axes when
    omitted (``:296-301``).  Telemetry is recorded at the controller's sample instants: column
    ``j`` of ``x_true`` is the plant state the j-th estimate was formed from (``:384-392``).
    """
    eng = engine or Engine(build_problem_c(sim_conditions, mpc_params, fail_params, debris, settings), device)
    try:
        p: Problem = eng.problem
        nsimD, nsimC, ratio = continuous_grid(sim_conditions)
        n_refresh, hold = noise_plan(sim_conditions)
        on_dev = type(x0_batch).__module__.startswith("torch")
        if on_dev:
            x0 = x0_batch
            B = x0.shape[1]
        else:
            x0_batch = np.asarray(x0_batch, float)
            B = x0_batch.shape[0]
            x0 = np.ascontiguousarray(x0_batch.T)
        if p.has_noise and noise_batch is None:
            rng = np.random.default_rng(seed)
            noise_batch = rng.standard_normal((n_refresh, 2, B)) * p.sig[0]
            if on_dev:
                import torch
                noise_batch = torch.from_numpy(noise_batch).to(x0.device)
        if not p.has_noise:
            noise_batch = None
        return eng.simulate_continuous(x0, noise_batch, nsimC, ratio, float(sim_conditions.T_cont), hold, record)
    finally:
        if engine is None:
            eng.close()


def trajectorySimulateC(sim_conditions, mpc_params, fail_params, debris):
    """Reference signature, one trajectory (``src/trajectorySimulateC.py:17-26``).

    Like the reference this leaves the global numpy RNG unseeded (``:28`` is commented out) and draws
    the disturbance sequence from it in ``ct.white_noise``'s order.  ``x_true_pcw`` (4 x i_term),
    ``ctrl_hist`` (2 x nsimC) and ``ctrlr_seq`` (i_term) are at every ``T_cont`` substep like the reference's
    (``:414-443``); entries the reference leaves uninitialised (``np.empty``) are NaN here.
    """
    nsimD, nsimC, ratio = continuous_grid(sim_conditions)
    n_refresh, hold = noise_plan(sim_conditions)
    noise = sim_conditions.noise
    nb = None
    sum_vec = np.full((4, nsimD), np.nan)
    if noise is not None:
        sx = noise.constructSigMat()[0, 0]
        W = np.array([np.random.normal(0, 1, n_refresh) for _ in range(2)])      # ct.white_noise(...), :301
        V = sx * W
        nb = np.ascontiguousarray(V.T[:, :, None])
        nl = int(noise.noise_length)
        for j in range(n_refresh):                                                # :304-307
            sum_vec[:, j * nl:nl * (1 + j)] = ratio * np.concatenate([V[:, j], np.zeros(2)]).reshape(-1, 1)
    x0 = np.asarray(sim_conditions.x0, float).reshape(1, 4)
    b = trajectorySimulateCBatch(sim_conditions, mpc_params, fail_params, debris, x0, nb,
                                 record=("x_est", "x_true_sub", "ctrl_sub", "ctrlr_sub"))
    it = int(b.i_term[0])
    seq = b.ctrlr_sub[:it, 0].astype(float)
    if it >= 2:
        seq[-1] = seq[-2]                                                         # :443
    return SimRun(it, bool(b.isSuccess[0]), b.x_true_sub[:, :it, 0].copy(), b.x_est[:, :, 0].copy(),
                  b.ctrl_sub[:, :, 0].copy(), seq, sum_vec)
