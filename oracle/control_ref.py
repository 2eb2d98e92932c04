"""Restatement of the python-control calls on the reference path (ORACLE -- test infrastructure).

[3P, python-control >=0.9.2, unpinned; source not in /root/reference]
Call sites: ``src/trajectorySimulate.py:185`` (``ct.dlqr(..., integral_action=)``),
``:198`` (``ct.acker``), ``src/trajectorySimulateC.py:212,225,301`` (``ct.white_noise``).
"""
import numpy as np
import scipy.linalg as sla


def dlqr_integral(A, B, Q, R, C_int):
    """``ct.dlqr(A, B, Q, R, integral_action=C_int)[0]``.

    python-control augments the discrete plant with one integrator per row of
    ``C_int``: ``A_aug=[[A,0],[C,I]]``, ``B_aug=[[B],[0]]``, solves the DARE and
    returns ``K=(R+B'XB)^-1 B'XA`` (reference use: ``trajectorySimulate.py:180-187``).
    """
    A = np.asarray(A, float)
    B = np.asarray(B, float)
    C = np.atleast_2d(np.asarray(C_int, float))
    ns, ni = A.shape[0], C.shape[0]
    A_aug = np.block([[A, np.zeros((ns, ni))], [C, np.eye(ni)]])
    B_aug = np.vstack([B, np.zeros((ni, B.shape[1]))])
    X = sla.solve_discrete_are(A_aug, B_aug, np.asarray(Q, float), np.asarray(R, float))
    K = np.linalg.solve(R + B_aug.T @ X @ B_aug, B_aug.T @ X @ A_aug)
    return K


def acker(A, B, poles):
    """``ct.acker(A, B, poles)``: Ackermann pole placement (single input).

    Reference use: deadbeat avoidance gain, ``trajectorySimulate.py:190-203``.
    """
    A = np.asarray(A, float)
    B = np.asarray(B, float).reshape(A.shape[0], 1)
    n = A.shape[0]
    ctrb = np.hstack([np.linalg.matrix_power(A, i) @ B for i in range(n)])
    p = np.real(np.poly(poles))
    npoly = p.size
    pmat = p[npoly - 1] * np.eye(n)
    for i in range(1, npoly):
        pmat = pmat + p[npoly - i - 1] * np.linalg.matrix_power(A, i)
    K = np.linalg.solve(ctrb, pmat)
    return K[-1:, :]


def white_noise(T, Q, rng):
    """``ct.white_noise(T, Q, dt=0.001)`` with dt != 0: unscaled N(0,1) sources
    mixed by ``sqrtm(Q)`` (reference use: ``trajectorySimulateC.py:301``).
    ``rng`` is a callable ``rng(size)`` returning standard normals (the reference
    uses numpy's legacy global RNG)."""
    T = np.atleast_1d(T)
    Q = np.atleast_2d(Q)
    W = np.array([rng(T.size) for _ in range(Q.shape[0])])
    return np.real(sla.sqrtm(Q)) @ W
