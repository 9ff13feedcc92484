"""Kalman operators (model/utils/costTool/KalmanFilter.py) backed by the batched GPU kernels.

``BatchedKalman`` is the natural GPU object (M tracks, state on the device).  The reference's
per-track helpers are kept as thin wrappers around it: ``init_kf_from_bbox`` returns a one-track
``KalmanState`` with ``predict()`` / ``update(z)`` like the filterpy object it replaces.
"""

import numpy as np
import torch

from . import _lib


def bbox_xyxy_to_z(bbox) -> np.ndarray:
    """KalmanFilter.py:5-16 (host scalar helper: float64 math, float32 result)."""
    x1, y1, x2, y2 = map(float, bbox)
    w = max(1.0, x2 - x1)
    h = max(1.0, y2 - y1)
    return np.array([x1 + 0.5 * w, y1 + 0.5 * h, w / h, h], dtype=np.float32)


def x_to_bbox_xyxy(x):
    """KalmanFilter.py:19-33."""
    cx, cy, a, h = float(x[0]), float(x[1]), float(x[2]), float(x[3])
    h = max(h, 1.0)
    a = max(a, 1e-3)
    w = max(a * h, 1.0)
    return (cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h)


class BatchedKalman:
    """M constant-velocity filters on the device: x [M,8], P [M,8,8] float64, stage [M] uint8."""

    def __init__(self, boxes_xyxy, std_pos=1.0, std_vel=10.0, std_meas_pos=1.0, std_meas_scale=1.0, device=None):
        if not torch.cuda.is_available():
            raise _lib.B200Error("no CUDA device: this package has no CPU path")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        b = torch.as_tensor(np.asarray(boxes_xyxy, dtype=np.float64).reshape(-1, 4)).to(dev)
        M = b.shape[0]
        self.device, self.M = dev, M
        self.x = torch.empty((M, 8), dtype=torch.float64, device=dev)
        self.P = torch.empty((M, 8, 8), dtype=torch.float64, device=dev)
        self.stage = torch.zeros((M,), dtype=torch.uint8, device=dev)
        q = np.array([std_pos] * 4 + [std_vel] * 4, dtype=np.float32)
        r = np.array([std_meas_pos, std_meas_pos, std_meas_scale, std_meas_scale], dtype=np.float32)
        self.q = torch.from_numpy(q * q).to(dev)
        self.r = torch.from_numpy(r * r).to(dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().b200_kalman_init(_lib.ptr(b), M, _lib.ptr(self.x), _lib.ptr(self.P),
                                                   _lib.ptr(self.stage), _lib.stream_ptr(dev)))

    def predict(self, want_boxes=False):
        """filterpy ``KalmanFilter.predict`` for every track (x <- Fx, P <- FPF' + Q) as ``predict_all`` uses it,
        mainTracking.py:340-345; optionally also the predicted boxes (x_to_bbox_xyxy, KalmanFilter.py:19-33)."""
        pb = torch.empty((self.M, 4), dtype=torch.float64, device=self.device) if want_boxes else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200_kalman_predict(_lib.ptr(self.x), _lib.ptr(self.P), _lib.ptr(self.stage), self.M,
                                                      _lib.ptr(self.q), _lib.ptr(pb), _lib.stream_ptr(self.device)))
        return pb

    def update(self, det_of_track, meas, meas_is_z=False):
        """filterpy ``KalmanFilter.update`` (Joseph form) of the tracks with a measurement, mainTracking.py:400;
        ``meas`` holds boxes (converted with bbox_xyxy_to_z, KalmanFilter.py:5-16) or z vectors."""
        d = torch.as_tensor(np.asarray(det_of_track, dtype=np.int32)).to(self.device)
        m = torch.as_tensor(np.asarray(meas, dtype=np.float64).reshape(-1, 4)).to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200_kalman_update(_lib.ptr(self.x), _lib.ptr(self.P), _lib.ptr(self.stage), self.M,
                                                     _lib.ptr(d), _lib.ptr(m), int(bool(meas_is_z)), _lib.ptr(self.r),
                                                     _lib.stream_ptr(self.device)))

    def maha(self, boxes_xyxy, C=None, maha_thr=9.49, INF=1e9):
        """d2 [M,N] float64; gates C (CUDA float32 [M,N]) in place when given (mainTracking.py:306-338)."""
        b = torch.as_tensor(np.asarray(boxes_xyxy, dtype=np.float64).reshape(-1, 4)).to(self.device)
        N = b.shape[0]
        d2 = torch.zeros((self.M, N), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200_maha_gate(_lib.ptr(self.x), _lib.ptr(self.P), _lib.ptr(self.stage), self.M,
                                                 _lib.ptr(b), N, _lib.ptr(self.r), float(maha_thr), float(INF),
                                                 _lib.ptr(C), N, _lib.ptr(d2), N, _lib.stream_ptr(self.device)))
        return d2


class KalmanState(BatchedKalman):
    """One track with the filterpy-like surface the reference uses (kf.predict(), kf.update(z), kf.x, kf.P)."""

    def update(self, z):  # noqa: D102 - z = bbox_xyxy_to_z(bbox), mainTracking.py:400
        super().update([0], np.asarray(z, dtype=np.float64).reshape(1, 4), meas_is_z=True)

    @property
    def x_host(self):
        x = self.x[0].cpu().numpy().reshape(8, 1)
        return x if int(self.stage[0]) >= 2 else x.astype(np.float32)

    @property
    def P_host(self):
        P = self.P[0].cpu().numpy()
        return P if int(self.stage[0]) >= 1 else P.astype(np.float32)


def init_kf_from_bbox(bbox_xyxy, dt=1.0, std_pos=1.0, std_vel=10.0, std_meas_pos=1.0, std_meas_scale=1.0):
    """KalmanFilter.py:36-101.  Only dt == 1 (the value the tracker uses) is supported."""
    if float(dt) != 1.0:
        raise NotImplementedError("the GPU filter hard-wires dt = 1 (KalmanFilter.py:38 default)")
    return KalmanState([list(map(float, bbox_xyxy))], std_pos, std_vel, std_meas_pos, std_meas_scale)


def gating_distance_maha(kf: BatchedKalman, bbox_xyxy) -> float:
    """KalmanFilter.py:105-116 for one (track, box) pair."""
    return float(kf.maha([list(map(float, bbox_xyxy))])[0, 0].item())
