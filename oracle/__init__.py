"""CPU oracle for the MPC_ARPO_Project hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy *restatement* of the reference's closed-loop MPC
path (``/root/reference/src/trajectorySimulate.py:17-388``,
``src/trajectorySimulateC.py:17-446``, ``src/simhelpers.py:11-189``) and of the
third-party numerics that path calls but that are absent from the reference
checkout and from this image:

* OSQP (C core, QDLDL)            -> ``oracle.osqp_ref``   [3P: osqp 0.6.x, unpinned]
* filterpy.kalman UKF             -> ``oracle.ukf_ref``    [3P: filterpy 1.4.5, unpinned]
* python-control dlqr/acker       -> ``oracle.control_ref``[3P: control >=0.9.2, unpinned]

PARITY UNPINNED: the reference ships no golden vectors, no asserts and no
dependency pins, and none of osqp / filterpy / control can be installed here
(no network).  What *is* pinned: the QP assembly (P, q, A, l, u and the per-step
``configureDynamicConstraints`` output) is checked entry-for-entry against the
reference's own ``src/`` code run in this container (third-party imports
stubbed, see ``oracle/gen_golden.py``) and committed under ``tests/golden/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``mpc_arpo_project_b200``) never does.
"""
