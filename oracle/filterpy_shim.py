"""Restatement of ``filterpy.kalman.KalmanFilter`` (filterpy 1.4.5) -- oracle only.

filterpy is a third-party dependency of the reference (requirements.txt:15) that is
neither under /root/reference nor installed in this image.  Only the members the
reference touches are restated: construction (KalmanFilter.py:57), the attribute
assignments at KalmanFilter.py:63-99, ``predict()`` (mainTracking.py:343) and
``update(z)`` (mainTracking.py:400).  PARITY UNPINNED: no reference test pins it.

The dtype behaviour matters and is kept: ``_I`` is a float64 identity, every other
matrix keeps whatever dtype the caller assigned (float32 in the reference), and all
products are plain numpy ``dot`` calls, so the state migrates float32 -> float64
exactly as described in SURVEY.md section 8 row K.
"""
import sys
import types

import numpy as np


class KalmanFilter:
    def __init__(self, dim_x, dim_z, dim_u=0):
        if dim_x < 1 or dim_z < 1 or dim_u < 0:
            raise ValueError("bad dimensions")
        self.dim_x, self.dim_z, self.dim_u = dim_x, dim_z, dim_u
        self.x = np.zeros((dim_x, 1))
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.B = None
        self.F = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self.R = np.eye(dim_z)
        self._alpha_sq = 1.0
        self.M = np.zeros((dim_x, dim_z))
        self.z = np.array([[None] * dim_z]).T
        self.K = np.zeros((dim_x, dim_z))
        self.y = np.zeros((dim_z, 1))
        self.S = np.zeros((dim_z, dim_z))
        self.SI = np.zeros((dim_z, dim_z))
        self._I = np.eye(dim_x)          # float64 on purpose (drives the dtype migration)
        self.x_prior, self.P_prior = self.x.copy(), self.P.copy()
        self.x_post, self.P_post = self.x.copy(), self.P.copy()
        self.inv = np.linalg.inv

    def predict(self, u=None, B=None, F=None, Q=None):
        B = self.B if B is None else B
        F = self.F if F is None else F
        if Q is None:
            Q = self.Q
        elif np.isscalar(Q):
            Q = np.eye(self.dim_x) * Q
        if B is not None and u is not None:
            self.x = np.dot(F, self.x) + np.dot(B, u)
        else:
            self.x = np.dot(F, self.x)
        self.P = self._alpha_sq * np.dot(np.dot(F, self.P), F.T) + Q
        self.x_prior, self.P_prior = self.x.copy(), self.P.copy()

    def update(self, z, R=None, H=None):
        if z is None:
            self.z = np.array([[None] * self.dim_z]).T
            self.x_post, self.P_post = self.x.copy(), self.P.copy()
            self.y = np.zeros((self.dim_z, 1))
            return
        z = np.atleast_2d(z)
        if z.shape[1] == self.dim_z:
            z = z.T
        if z.shape != (self.dim_z, 1):
            raise ValueError("z must be convertible to shape ({}, 1)".format(self.dim_z))
        if R is None:
            R = self.R
        elif np.isscalar(R):
            R = np.eye(self.dim_z) * R
        H = self.H if H is None else H
        self.y = z - np.dot(H, self.x)
        PHT = np.dot(self.P, H.T)
        self.S = np.dot(H, PHT) + R
        self.SI = self.inv(self.S)
        self.K = np.dot(PHT, self.SI)
        self.x = self.x + np.dot(self.K, self.y)
        I_KH = self._I - np.dot(self.K, H)
        self.P = np.dot(np.dot(I_KH, self.P), I_KH.T) + np.dot(np.dot(self.K, R), self.K.T)
        self.z = z.copy()
        self.x_post, self.P_post = self.x.copy(), self.P.copy()


def install():
    """Register this module as ``filterpy.kalman`` so reference code imports unchanged."""
    if "filterpy.kalman" in sys.modules:
        return
    pkg = types.ModuleType("filterpy")
    sub = types.ModuleType("filterpy.kalman")
    sub.KalmanFilter = KalmanFilter
    pkg.kalman = sub
    sys.modules["filterpy"] = pkg
    sys.modules["filterpy.kalman"] = sub
