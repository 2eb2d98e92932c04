"""b200-mpc-batch: batched closed-loop MPC engine for B200 (sm_100a).

Drop-in for the hot path of IsaacTroche1/MPC_ARPO_Project
(``trajectorySimulate`` / ``trajectorySimulateC`` + the ``mpcsim`` objects).
"""
from .mpcsim import Noise, SimConditions, SimRun, Debris, MPCParams, FailsafeParams, BatchSimRun  # noqa: F401
