"""b200-mpc-batch: batched closed-loop MPC engine for B200 (sm_100a).

Drop-in for the hot path of IsaacTroche1/MPC_ARPO_Project
(``trajectorySimulate`` / ``trajectorySimulateC`` + the ``mpcsim`` objects).
The compute path is ``lib/libmpcb.so`` (hand-written CUDA behind the C ABI of
``include/mpcb.h``); importing this package does not need a GPU, running it does.
"""
from .mpcsim import Noise, SimConditions, SimRun, Debris, MPCParams, FailsafeParams, BatchSimRun  # noqa: F401
from .problem import Problem, SolverSettings, build_problem  # noqa: F401
from .engine import Engine  # noqa: F401
from .trajectorySimulate import trajectorySimulate, trajectorySimulateBatch  # noqa: F401
from .trajectorySimulateC import trajectorySimulateC, trajectorySimulateCBatch, build_problem_c  # noqa: F401
from .montecarlo import RatioSweep, final_distance_ratio_sweep, success_rate  # noqa: F401
from . import presets  # noqa: F401
