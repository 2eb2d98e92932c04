"""TEST INFRASTRUCTURE (oracle): the linear Kalman-filter estimator variant, SURVEY 8(f)-4.

Restates the estimator of the reference's prototype ``misc/MPCrendezKALMANdisturb.py:261-266``::

    xnom = Ao@xestO[:,i] + Bou@ctrl
    Pest = Ao@Pest@Ao.T + Qw
    L    = Pest@Co.T@inv(Co@Pest@Co.T)            # measurement noise R = 0
    xestO[:,i+1] = xnom + L@(ymeas - Co@xnom)
    Pest = (eye(nx+ndi) - L@Co)@Pest

with ``Co = [Cm 0]``, ``Cm = [I2 0]`` (position measurement, ``:98-99``).  The class exposes the ``predict(u)`` /
``update(z)`` protocol of ``ukf_ref.UKFRef`` so ``sim_ref`` can call it where the simulators call the UKF; the
process noise ``Qw`` and the initial covariance are the simulators' own (``trajectorySimulate.py:250, 272-275``).
PARITY UNPINNED: the prototype is a script (plots, no function seam) and is not on the reference's ``src`` path; what
is pinned is this restatement of its six lines.
"""
import numpy as np


class _NoPoints:
    clamped = 0


class LinearKFRef:
    def __init__(self, Ao, Bou, dim_x=6, dim_z=2):
        self.Ao, self.Bou = np.asarray(Ao, float), np.asarray(Bou, float)
        self.Co = np.hstack([np.eye(dim_z), np.zeros((dim_z, dim_x - dim_z))])
        self.x = np.zeros(dim_x)
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.R = np.zeros((dim_z, dim_z))
        self.points = _NoPoints()

    def predict(self, u):
        self.x = self.Ao @ self.x + self.Bou @ np.asarray(u, float)
        self.P = self.Ao @ self.P @ self.Ao.T + self.Q

    def update(self, z):
        Co = self.Co
        L = self.P @ Co.T @ np.linalg.inv(Co @ self.P @ Co.T + self.R)
        self.x = self.x + L @ (np.asarray(z, float) - Co @ self.x)
        self.P = (np.eye(self.x.size) - L @ Co) @ self.P
