"""Oracle (test-only): model/utils/costTool/KalmanFilter.py restated on numpy.

The Kalman object is ``oracle.filterpy_shim.KalmanFilter`` (filterpy 1.4.5 semantics).
"""
import numpy as np

from .filterpy_shim import KalmanFilter


def bbox_xyxy_to_z(bbox):
    """KalmanFilter.py:5-16.  Python-float math, result rounded to float32."""
    x1, y1, x2, y2 = (float(v) for v in bbox)
    w = x2 - x1
    if w < 1.0:
        w = 1.0
    h = y2 - y1
    if h < 1.0:
        h = 1.0
    return np.array([x1 + 0.5 * w, y1 + 0.5 * h, w / h, h], dtype=np.float32)


def x_to_bbox_xyxy(x):
    """KalmanFilter.py:19-33.  State -> corner box, with the h>=1, a>=1e-3, w>=1 clamps."""
    cx, cy, a, h = float(x[0]), float(x[1]), float(x[2]), float(x[3])
    if h < 1.0:
        h = 1.0
    if a < 1e-3:
        a = 1e-3
    w = a * h
    if w < 1.0:
        w = 1.0
    return (cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h)


def init_kf_from_bbox(bbox_xyxy, dt=1.0, std_pos=1.0, std_vel=10.0,
                      std_meas_pos=1.0, std_meas_scale=1.0):
    """KalmanFilter.py:36-101.  8-state constant-velocity filter; every matrix float32."""
    kf = KalmanFilter(dim_x=8, dim_z=4)
    F = np.eye(8, dtype=np.float32)
    F[np.arange(4), np.arange(4) + 4] = dt
    kf.F = F
    H = np.zeros((4, 8), dtype=np.float32)
    H[np.arange(4), np.arange(4)] = 1.0
    kf.H = H
    x0 = np.zeros((8, 1), dtype=np.float32)
    x0[:4, 0] = bbox_xyxy_to_z(bbox_xyxy)
    kf.x = x0
    kf.P = np.diag(np.array([10.0] * 4 + [1000.0] * 4, dtype=np.float32))
    q = np.array([std_pos] * 4 + [std_vel] * 4, dtype=np.float32)
    kf.Q = np.diag(q * q)
    r = np.array([std_meas_pos, std_meas_pos, std_meas_scale, std_meas_scale], dtype=np.float32)
    kf.R = np.diag(r * r)
    return kf


def gating_distance_maha(kf, bbox_xyxy):
    """KalmanFilter.py:105-116.  d^2 = y^T (S + 1e-9 I)^-1 y with S = H P H^T + R."""
    z = bbox_xyxy_to_z(bbox_xyxy).reshape(4, 1).astype(np.float32)
    innov = z - kf.H @ kf.x
    S = kf.H @ kf.P @ kf.H.T + kf.R
    Sinv = np.linalg.inv(S + 1e-9 * np.eye(4, dtype=np.float32))
    return float((innov.T @ Sinv @ innov)[0, 0])
