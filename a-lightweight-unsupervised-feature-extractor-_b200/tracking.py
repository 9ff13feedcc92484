"""GPU-resident tracker with the reference's surface (model/mainTracking.py).

* ``Tracking``            - drop-in for the reference class: ``Tracking().update(obj)`` returns
                            ``(matches_tid, unmatched_track_ids, unmatched_dets)`` (mainTracking.py:450-610).
* ``MultiStreamTracker``  - the same association step for S independent video streams per launch
                            (the batching the reference lacks).

All state (Kalman filters, EMA embeddings, history banks, ages) lives on the device inside a
``b200_tracker`` handle; one ``update`` is one host->device copy of the detections, five kernels and
one device->host copy of a small result table.  There is no CPU fallback.
"""
import ctypes
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import yaml

from . import _lib, cost as cost_ops, kalman as kalman_ops

# mainTracking.py:55-96 -- defaults used when a key is missing from the YAML `tracker` block.
CODE_DEFAULTS = dict(
    init_conf_min=0.5, hist_max=10, emb_top_k=5, app_tau=0.07, eps=1e-12, w_app=1.0, w_bbox=0.3, w_conf=0.2,
    alpha=1.0, beta=0.5, unmatch_cost=10.0, cost_max=50.0, max_age=30, ema_alpha=0.9, conf_update_min=0.55,
    cost_update_max=30.0, maha_thr=9.49, lost_reid_after=60, reid_sim_min=0.6)

# model/conf/conf.yaml:2-24 -- the tracker block the reference ships.
SHIPPED_CONF = dict(
    init_conf_min=0.5, hist_max=30, emb_top_k=5, app_tau=0.07, eps=1e-12, w_app=1.0, w_bbox=0.3, w_conf=0.2,
    alpha=1.0, beta=0.5, unmatch_cost=10.0, cost_max=50.0, max_age=120, ema_alpha=0.9, conf_update_min=0.55,
    cost_update_max=30.0, maha_thr=9.49, lost_reid_after=50, reid_sim_min=0.6, reid_only_cost_max=0.4)

R_NMATCH, R_NUT, R_NUD, R_NLIVE, R_NEXT, R_STATUS, R_M1, R_M2, R_HDR = range(9)


def load_conf(path: str) -> Dict[str, Any]:
    """mainTracking.py:11-13."""
    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


def resolve_conf(tcfg: Dict[str, Any]) -> Dict[str, Any]:
    """Applies the reference's per-key defaults and the reid_only_cost_max rule (:92-96)."""
    c = dict(CODE_DEFAULTS)
    c.update({k: v for k, v in tcfg.items() if v is not None})
    if "reid_only_cost_max" not in tcfg:
        c["reid_only_cost_max"] = 1.0 - float(c["reid_sim_min"])
    for k in ("hist_max", "emb_top_k", "max_age", "lost_reid_after"):
        c[k] = int(c[k])
    return c


def _c_conf(c):
    s = _lib.tracker_conf()
    for name, _ in _lib.tracker_conf._fields_:
        setattr(s, name, c[name])
    return s


class MultiStreamTracker:
    """S independent trackers stepped by the same kernel launches.

    Detections are passed as dense host arrays padded to ``max_dets`` per stream:
    ``n_det[S]`` (-1 = stream idle, 0 = empty frame), ``boxes[S,max_dets,4]`` float64 xyxy,
    ``confs[S,max_dets]`` float64, ``embs[S,max_dets,128]`` float32, ``frame_ids[S]``.
    """

    def __init__(self, n_streams: int, conf: Optional[Dict[str, Any]] = None, max_tracks: int = 256,
                 max_dets: int = 128, device=None, auto_grow: bool = True):
        if not torch.cuda.is_available():
            raise _lib.B200Error("no CUDA device: this package has no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.conf = resolve_conf(SHIPPED_CONF if conf is None else conf)
        self.S, self.auto_grow = int(n_streams), bool(auto_grow)
        self._h = ctypes.c_void_p()
        self._create(int(max_tracks), int(max_dets))
        self.n_live = np.zeros(self.S, dtype=np.int64)
        self._n_live_stale = False                 # set by step_device: the host copy of the live counts is out of date
        self._pending = []                         # StepHandles of step_async calls whose result has not been collected
        self._pending_births = np.zeros(self.S, dtype=np.int64)

    def n_live_now(self) -> np.ndarray:
        """Live tracks per stream as of the work queued so far (re-read from the device after device-side steps)."""
        self.drain()
        if self._n_live_stale:
            n = np.zeros(self.S, np.int32)
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().b200_tracker_live_counts(self._h, n.ctypes.data_as(ctypes.c_void_p), None,
                                                               _lib.stream_ptr(self.device)))
            self.n_live[:] = n
            self._n_live_stale = False
        return self.n_live

    def _create(self, max_tracks, max_dets):
        h = ctypes.c_void_p()
        cc = _c_conf(self.conf)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200_tracker_create(ctypes.byref(h), self.S, max_tracks, max_dets, ctypes.byref(cc)))
        self._h, self.max_tracks, self.max_dets = h, max_tracks, max_dets
        self.stride = _lib.lib().b200_tracker_result_stride(h)
        self._res = np.zeros((self.S, self.stride), dtype=np.int32)

    def grow(self, max_tracks: Optional[int] = None, max_dets: Optional[int] = None):
        """Migrates every stream's state into a handle with larger capacities (export -> create -> import)."""
        self.drain()
        snaps = [self.export(s) for s in range(self.S)]
        old = self._h
        self._create(max(self.max_tracks, int(max_tracks or 0)), max(self.max_dets, int(max_dets or 0)))
        _lib.lib().b200_tracker_destroy(old)
        for s, snap in enumerate(snaps):
            self.import_state(s, snap)

    def import_state(self, stream: int, snap: Dict[str, np.ndarray]):
        """Inverse of ``export`` for one stream."""
        n = len(snap["ids"])
        c = lambda k, dt: np.ascontiguousarray(snap[k], dtype=dt)  # noqa: E731
        arrs = [c("ids", np.int32), c("x", np.float64), c("P", np.float64), c("stage", np.uint8), c("ema", np.float32),
                c("bank", np.float32), c("bank_len", np.int32), c("miss", np.int32), c("age", np.int32),
                c("last_bbox", np.float64), c("last_conf", np.float64), c("last_cost", np.float64)]
        with torch.cuda.device(self.device):
            rc = _lib.lib().b200_tracker_import(self._h, int(stream), n, *[a.ctypes.data_as(ctypes.c_void_p) for a in arrs],
                                                int(snap["next_id"]), _lib.stream_ptr(self.device))
        _lib.check(rc)
        self.n_live[stream] = n

    def close(self):
        """Frees the device state (b200_tracker_destroy)."""
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().b200_tracker_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        """Back to the state of ``Tracking.__init__`` (mainTracking.py:45-96): no tracks, next id 0, every stream."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200_tracker_reset(self._h, _lib.stream_ptr(self.device)))
        self.n_live[:] = 0

    # -- one frame-step for every stream, host arrays in, host result table out -------------------
    def step(self, n_det, boxes, confs, embs, frame_ids) -> np.ndarray:
        """One ``Tracking.update`` (mainTracking.py:450-610) for every stream from host arrays padded to ``max_dets``:
        n_det [S], boxes [S,max_dets,4] float64 xyxy, confs [S,max_dets] float64, embs [S,max_dets,128] float32,
        frame_ids [S].  Returns the int32 result table [S, stride] (``decode`` turns a row into the reference's
        return value); raises ValueError where scipy would (NaN / infeasible cost matrix, hung.py:28)."""
        self.drain()                                   # earlier asynchronous steps first (their results stay on their handles)
        n_det, boxes, confs, embs, frame_ids = self._prepare(n_det, boxes, confs, embs, frame_ids, False)
        res = np.zeros((self.S, self.stride), dtype=np.int32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        with torch.cuda.device(self.device):           # synchronous entry point: copies and kernels on one stream
            rc = _lib.lib().b200_tracker_step_host(self._h, p(n_det), p(boxes), p(confs), p(embs), p(frame_ids), p(res),
                                                   _lib.stream_ptr(self.device))
        _lib.check(rc)
        h = StepHandle(self, -1, np.zeros(self.S, np.int64))
        h._finish(res)
        return h.result()

    def step_async(self, n_det, boxes, confs, embs, frame_ids, *, pinned: bool = False) -> "StepHandle":
        """``step`` without the wait: uploads the detections, queues the step and the download of the result table on the
        current CUDA stream and returns at once; ``handle.result()`` blocks until THAT step is on the host and returns
        (or raises) exactly what ``step`` would.  The reference's consumer is a queue (tracking.py:329): a caller can
        queue frame t+1 before it looks at the result of frame t.  At most four steps may be pending; results must be
        collected in order.

        ``pinned=True``: ``boxes`` / ``confs`` / ``embs`` are page-locked arrays of the exact dtype and shape (numpy views
        of ``torch.empty(..., pin_memory=True)``, or tensors) that the caller leaves untouched until the result is
        collected; they are uploaded by DMA from where they are instead of through the handle's staging ring."""
        n_det, boxes, confs, embs, frame_ids = self._prepare(n_det, boxes, confs, embs, frame_ids, pinned)
        if len(self._pending) >= 4:
            raise _lib.B200Error("step_async: four steps are already pending; collect their results first")
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        ticket = ctypes.c_int64(-1)
        fn = _lib.lib().b200_tracker_step_pinned_async if pinned else _lib.lib().b200_tracker_step_host_async
        with torch.cuda.device(self.device):
            rc = fn(self._h, p(n_det), p(boxes), p(confs), p(embs), p(frame_ids), ctypes.byref(ticket),
                    _lib.stream_ptr(self.device))
        _lib.check(rc)
        h = StepHandle(self, int(ticket.value), np.maximum(n_det, 0).astype(np.int64))
        h._keep = (boxes, confs, embs) if pinned else None     # the DMA reads these after this call returns
        self._pending.append(h)
        self._pending_births = self._pending_births + h._births
        return h

    def _prepare(self, n_det, boxes, confs, embs, frame_ids, pinned):
        """Argument conversion and the capacity check shared by ``step`` and ``step_async``."""
        n_det = np.ascontiguousarray(n_det, dtype=np.int32).reshape(self.S)
        frame_ids = np.ascontiguousarray(frame_ids, dtype=np.int32).reshape(self.S)
        if pinned:
            arrs = []
            for a, dt, shape in ((boxes, np.float64, (self.S, self.max_dets, 4)), (confs, np.float64, (self.S, self.max_dets)),
                                 (embs, np.float32, (self.S, self.max_dets, 128))):
                a = a.numpy() if isinstance(a, torch.Tensor) else a
                if not (isinstance(a, np.ndarray) and a.dtype == dt and a.shape == shape and a.flags.c_contiguous):
                    raise TypeError("step_async(pinned=True): arrays must be C-contiguous %s of shape %s" % (np.dtype(dt), shape))
                arrs.append(a)
            boxes, confs, embs = arrs
        else:
            boxes = np.ascontiguousarray(boxes, dtype=np.float64).reshape(self.S, self.max_dets, 4)
            confs = np.ascontiguousarray(confs, dtype=np.float64).reshape(self.S, self.max_dets)
            embs = np.ascontiguousarray(embs, dtype=np.float32).reshape(self.S, self.max_dets, 128)
        # capacity: live tracks are only known up to the last collected result; every pending step may have added
        # all of its detections
        bound = self.n_live + self._pending_births + np.maximum(n_det, 0)
        while int(bound.max()) > self.max_tracks and self._pending:     # collect the oldest results until the bound fits
            self._pending[0]._collect()
            bound = self.n_live + self._pending_births + np.maximum(n_det, 0)
        if int(bound.max()) > self.max_tracks and self._n_live_stale:
            bound = self.n_live_now() + np.maximum(n_det, 0)
        if int(bound.max()) > self.max_tracks:         # the reference is unbounded: migrate to a larger handle
            if not self.auto_grow:
                raise _lib.B200Error("tracker capacity: live tracks + detections could exceed max_tracks=%d; "
                                     "construct the tracker with a larger max_tracks" % self.max_tracks)
            self.grow(max_tracks=max(int(bound.max()), 2 * self.max_tracks))
        return n_det, boxes, confs, embs, frame_ids

    def drain(self):
        """Waits for every pending ``step_async`` (their results stay available on their handles)."""
        for h in list(self._pending):
            h._collect()

    # -- device-resident inputs, asynchronous on the current stream (no read-back) ------------------
    def step_device(self, n_det: torch.Tensor, boxes: torch.Tensor, confs: torch.Tensor, embs: torch.Tensor,
                    frame_ids: torch.Tensor, result: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The same step with device-resident inputs, asynchronous on the current stream, no read-back (SURVEY 8f-1:
        the reference rebuilds host lists every frame, tracking.py:315-324).  The per-stream status is in column
        R_STATUS of the returned table."""
        for t, dt in ((n_det, torch.int32), (boxes, torch.float64), (confs, torch.float64), (embs, torch.float32),
                      (frame_ids, torch.int32)):
            _lib.require_cuda(t, "input")
            if t.dtype != dt or not t.is_contiguous():
                raise TypeError("step_device: wrong dtype or non-contiguous input")
        if result is None:
            result = torch.empty((self.S, self.stride), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().b200_tracker_step(self._h, _lib.ptr(n_det), _lib.ptr(boxes), _lib.ptr(confs), _lib.ptr(embs),
                                              _lib.ptr(frame_ids), _lib.ptr(result), _lib.stream_ptr(self.device))
        _lib.check(rc)
        self._n_live_stale = True
        return result

    def decode(self, row: np.ndarray):
        """Result-table row -> (matches_tid, unmatched_track_ids, unmatched_dets) as Tracking.update returns."""
        MD, MT = self.max_dets, self.max_tracks
        nm, nut, nud = int(row[R_NMATCH]), int(row[R_NUT]), int(row[R_NUD])
        m = row[R_HDR:R_HDR + 2 * nm].reshape(nm, 2)
        ut = row[R_HDR + 2 * MD:R_HDR + 2 * MD + nut]
        ud = row[R_HDR + 2 * MD + MT:R_HDR + 2 * MD + MT + nud]
        return list(map(tuple, m.tolist())), ut.tolist(), ud.tolist()

    def export(self, stream: int = 0) -> Dict[str, np.ndarray]:
        """Copies one stream's live tracks (ascending track id) to host arrays."""
        MT, H = self.max_tracks, self.conf["hist_max"]
        out = dict(ids=np.zeros(MT, np.int32), x=np.zeros((MT, 8)), P=np.zeros((MT, 8, 8)), stage=np.zeros(MT, np.uint8),
                   ema=np.zeros((MT, 128), np.float32), bank=np.zeros((MT, H, 128), np.float32),
                   bank_len=np.zeros(MT, np.int32), miss=np.zeros(MT, np.int32), age=np.zeros(MT, np.int32),
                   last_bbox=np.zeros((MT, 4)), last_conf=np.zeros(MT), last_cost=np.zeros(MT))
        nxt = ctypes.c_int32(0)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        with torch.cuda.device(self.device):
            n = _lib.lib().b200_tracker_export(
                self._h, int(stream), p(out["ids"]), p(out["x"]), p(out["P"]), p(out["stage"]), p(out["ema"]),
                p(out["bank"]), p(out["bank_len"]), p(out["miss"]), p(out["age"]), p(out["last_bbox"]),
                p(out["last_conf"]), p(out["last_cost"]), ctypes.byref(nxt), _lib.stream_ptr(self.device))
        if n < 0:
            _lib.check(n)
        out = {k: v[:n] for k, v in out.items()}
        out["next_id"] = int(nxt.value)
        return out


class StepHandle:
    """A queued ``MultiStreamTracker.step_async``; ``result()`` is what ``step`` would have returned."""

    def __init__(self, ms: MultiStreamTracker, ticket: int, births: np.ndarray):
        self._ms, self._ticket, self._births = ms, ticket, births
        self._res, self._error = None, None
        self.table = None                              # the step's result table once collected, also when result() raises

    def _collect(self):
        ms = self._ms
        if self._res is not None or self._error is not None:
            return
        if not ms._pending or ms._pending[0] is not self:
            for h in list(ms._pending):                # results come back in queue order
                if h is self:
                    break
                h._collect()
        res = np.zeros((ms.S, ms.stride), dtype=np.int32)
        with torch.cuda.device(ms.device):
            rc = _lib.lib().b200_tracker_step_result(ms._h, self._ticket, res.ctypes.data_as(ctypes.c_void_p))
        ms._pending.remove(self)
        ms._pending_births = ms._pending_births - self._births
        try:
            _lib.check(rc)
        except Exception as exc:                       # noqa: BLE001
            self._error = exc
            return
        self._finish(res)

    def _finish(self, res):
        """Takes the step's result table: live counts, last_result, and the error ``result()`` will raise, if any."""
        ms = self._ms
        self._res = self.table = res
        ms._res = res
        ms.n_live[:] = res[:, R_NLIVE]
        # A failing stream does not lose the others: the table (handle.table; ms.last_result = the most recently collected
        # one) is complete before anything is raised, and the failing stream's state is what the reference leaves behind
        # when scipy raises (predict only).
        ms.last_result = res
        bad = np.nonzero(res[:, R_STATUS])[0]
        if len(bad):
            st = int(res[bad[0], R_STATUS])
            if st == _lib.ENUMERIC:
                self._error = ValueError("matrix contains invalid numeric entries (stream %d)" % bad[0])
            elif st == _lib.EINFEASIBLE:
                self._error = ValueError("cost matrix is infeasible (stream %d)" % bad[0])
            elif st == _lib.ECAPACITY:
                self._error = _lib.B200Error(
                    "tracker capacity exceeded on stream %d: births were dropped because live tracks + detections "
                    "exceeded max_tracks=%d (device-side steps do not grow the handle; call grow() or construct with a "
                    "larger max_tracks)" % (bad[0], ms.max_tracks))
            else:
                self._error = _lib.B200Error("tracker step failed on stream %d with status %d" % (bad[0], st))

    def done(self) -> bool:
        return self._res is not None or self._error is not None

    def result(self) -> np.ndarray:
        self._collect()
        if self._error is not None:
            raise self._error
        return self._res


class TrackView:
    """Read-only host snapshot of one track (TrackState + TrackMemory, mainTracking.py:15-42)."""

    def __init__(self, snap, i):
        self.track_id = int(snap["ids"][i])
        self.miss_count = int(snap["miss"][i])
        self.age = int(snap["age"][i])
        self.state = "ACTIVE" if self.miss_count == 0 else "LOST"
        st = int(snap["stage"][i])
        self.x = snap["x"][i].reshape(8, 1) if st >= 2 else snap["x"][i].reshape(8, 1).astype(np.float32)
        self.P = snap["P"][i] if st >= 1 else snap["P"][i].astype(np.float32)
        self.encoder_feat = snap["ema"][i]
        n = int(snap["bank_len"][i])
        self.feat_historical = [snap["bank"][i, t] for t in range(n)]
        self.last_bbox = tuple(float(v) for v in snap["last_bbox"][i])
        self.last_conf = float(snap["last_conf"][i])
        c = float(snap["last_cost"][i])
        self.last_match_cost = None if np.isnan(c) else c


class Tracking:
    """Drop-in for the reference's ``Tracking`` (model/mainTracking.py:45).

    ``Tracking()`` reads ``model/conf/conf.yaml`` relative to the working directory exactly like the
    reference (:47); pass ``conf=`` (a tracker block) to skip the file.  ``max_tracks`` / ``max_dets``
    are initial capacities of the device-side state: like the reference the tracker is unbounded, a step
    that would not fit migrates the state into a larger handle first (``auto_grow=False`` raises instead).
    """

    def __init__(self, conf_path: str = "model/conf/conf.yaml", *, conf: Optional[Dict[str, Any]] = None,
                 max_tracks: int = 512, max_dets: int = 256, device=None, auto_grow: bool = True):
        if conf is None:
            full = load_conf(conf_path)
            if "tracker" not in full:
                raise KeyError("Missing 'tracker' section in YAML config.")
            conf = full["tracker"]
        self._ms = MultiStreamTracker(1, conf, max_tracks, max_dets, device, auto_grow)
        c = self._ms.conf
        for k, v in c.items():                       # same attribute names as the reference (:55-96)
            setattr(self, "tau" if k == "app_tau" else k, v)
        MD = self._ms.max_dets
        self._boxes = np.zeros((1, MD, 4), np.float64)
        self._confs = np.zeros((1, MD), np.float64)
        self._embs = np.zeros((1, MD, 128), np.float32)
        self._snap = None

    @property
    def device(self):
        return self._ms.device

    # -- the call the pipeline makes (tracking.py:326) ---------------------------------------------
    def update(self, obj: Dict):
        """``Tracking.update`` (mainTracking.py:450-610): obj = {embs, bboxes, confs, input_hw, frame_id} ->
        (matches [(track_id, det_idx)], unmatched track ids, unmatched det indices), same order, same ValueErrors
        (:457-462, :267-268)."""
        det_embs = obj.get("embs", []) or []
        det_boxes = obj.get("bboxes", []) or []
        det_confs = obj.get("confs", []) or []
        input_hw = obj.get("input_hw", None)
        frame_id = obj.get("frame_id", None)
        if input_hw is None:
            raise ValueError("obj['input_hw'] is required")
        if frame_id is None:
            raise ValueError("obj['frame_id'] is required")
        if not (len(det_embs) == len(det_boxes) == len(det_confs)):
            raise ValueError("Length mismatch: embs/bboxes/confs must have same length")
        N = len(det_boxes)
        if N:
            try:                                   # equal-length vectors: one C-level conversion
                e = np.asarray(det_embs, dtype=np.float32).reshape(N, -1)
            except ValueError:                     # ragged input: per-vector conversion, then the reference's error
                e = None
            if e is None or e.shape[1] != 128:
                shapes = sorted({np.asarray(v).size for v in det_embs})
                raise ValueError(f"det_embs must be 128D, got sizes {shapes}")
            return self.update_arrays(np.asarray(det_boxes, dtype=np.float64), np.asarray(det_confs, dtype=np.float64),
                                      e, int(frame_id))
        return self.update_arrays(np.zeros((0, 4)), np.zeros((0,)), np.zeros((0, 128), np.float32), int(frame_id))

    def update_arrays(self, boxes: np.ndarray, confs: np.ndarray, embs: np.ndarray, frame_id: int):
        """Array form of ``update``: boxes [N,4] float64 xyxy, confs [N], embs [N,128] float32."""
        N = boxes.shape[0]
        if N > self._ms.max_dets:
            if not self._ms.auto_grow:
                raise ValueError("%d detections exceed max_dets=%d" % (N, self._ms.max_dets))
            self._ms.grow(max_dets=max(N, 2 * self._ms.max_dets))
        MD = self._ms.max_dets
        if self._boxes.shape[1] != MD:                 # the handle has grown (here or in one of the step-by-step methods)
            self._boxes, self._confs = np.zeros((1, MD, 4), np.float64), np.zeros((1, MD), np.float64)
            self._embs = np.zeros((1, MD, 128), np.float32)
        self._boxes[0, :N] = boxes
        self._confs[0, :N] = confs
        self._embs[0, :N] = embs
        res = self._ms.step([N], self._boxes, self._confs, self._embs, [frame_id])
        self._snap = None
        return self._ms.decode(res[0])

    # -- the reference's methods one by one (mainTracking.py:340-448), on the device state -----------
    def _dets(self, det_embs, det_boxes, det_confs):
        n = len(det_boxes)
        if not (len(det_embs) == n == len(det_confs)):
            raise ValueError("Length mismatch: embs/bboxes/confs must have same length")
        if n > self._ms.max_dets:
            self._ms.grow(max_dets=max(n, 2 * self._ms.max_dets))
        e = np.zeros((n, 128), np.float32)
        for j, v in enumerate(det_embs):
            v = np.asarray(v, dtype=np.float32).reshape(-1)
            if v.shape[0] != 128:
                raise ValueError(f"emb must be shape (128,), got {v.shape}")
            e[j] = v
        return (n, np.ascontiguousarray(np.asarray(det_boxes, dtype=np.float64).reshape(n, 4)),
                np.ascontiguousarray(np.asarray(det_confs, dtype=np.float64).reshape(n)), e)

    def _op(self, name, *args):
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if isinstance(a, np.ndarray) else a  # noqa: E731
        with torch.cuda.device(self.device):
            rc = getattr(_lib.lib(), name)(self._ms._h, 0, *[p(a) for a in args], _lib.stream_ptr(self.device))
        self._snap = None
        self._ms._n_live_stale = True
        return rc

    def predict_all(self):
        """mainTracking.py:340-345: Kalman predict of every track; ``last_bbox`` becomes the predicted box."""
        _lib.check(self._op("b200_tracker_predict_all"))

    def mark_missed(self, track_ids: List[int]):
        """mainTracking.py:347-355 (ids that are not live are skipped)."""
        ids = np.ascontiguousarray(np.asarray(list(track_ids), dtype=np.int32))
        _lib.check(self._op("b200_tracker_mark_missed", ids, len(ids)))

    def purge_dead(self):
        """mainTracking.py:357-360: drops tracks whose miss_count exceeds max_age."""
        _lib.check(self._op("b200_tracker_purge_dead"))

    def create_new_tracks(self, det_ids, det_embs, det_boxes, det_confs, frame_id):
        """mainTracking.py:362-373: one new track per listed detection with conf >= init_conf_min, ids in list order."""
        ids = np.ascontiguousarray(np.asarray(list(det_ids), dtype=np.int32))
        if len(ids) == 0:
            return
        n, boxes, confs, embs = self._dets(det_embs, det_boxes, det_confs)
        if int((confs[ids] >= self.init_conf_min).sum()) + int(self._ms.n_live_now()[0]) > self._ms.max_tracks:
            self._ms.grow(max_tracks=2 * self._ms.max_tracks + len(ids))
        rc = self._op("b200_tracker_create_tracks", ids, len(ids), boxes, confs, embs, n, int(frame_id))
        if rc < 0:
            _lib.check(rc)

    def update_matched(self, matches, row_to_tid, det_embs, det_boxes, det_confs, frame_id, C_total_np, *,
                       ema_alpha=0.9, conf_update_min=0.55, cost_update_max=50.0, maha_thr=9.49):
        """mainTracking.py:375-448: Kalman update + bookkeeping for every (row, det) match and, behind the
        confidence / cost / posterior-Mahalanobis gates, the EMA embedding and history-bank update."""
        matches = list(matches)
        if not matches:
            return
        try:
            tids = np.array([row_to_tid[r] for r, _ in matches], dtype=np.int32)
        except IndexError as exc:
            raise IndexError("update_matched: match row outside row_to_tid") from exc
        dets = np.array([j for _, j in matches], dtype=np.int32)
        C = np.asarray(C_total_np)
        costs = np.ascontiguousarray(np.array([C[r, j] for r, j in matches], dtype=np.float64).astype(np.float32))
        if not np.array_equal(costs.astype(np.float64), np.array([float(C[r, j]) for r, j in matches])):
            raise TypeError("update_matched: C_total_np must hold float32-representable costs (the reference passes float32)")
        n, boxes, confs, embs = self._dets(det_embs, det_boxes, det_confs)
        if len(dets) and (dets.min() < 0 or dets.max() >= n):
            raise IndexError("update_matched: detection index out of range")       # the reference: det_boxes[det_j] at :393
        rc = self._op("b200_tracker_update_matched", tids, dets, costs, len(matches), boxes, confs, embs, n, int(frame_id),
                      float(ema_alpha), float(conf_update_min), float(cost_update_max), float(maha_thr))
        _lib.check(rc, {_lib.EINVAL: KeyError})             # a track id that is not live: self.tracks[tid] at :390

    # -- state inspection ---------------------------------------------------------------------------
    def snapshot(self) -> Dict[str, np.ndarray]:
        """Host copy of the live tracks in ascending id order: the fields of TrackMemory / TrackState
        (mainTracking.py:15-42) plus the Kalman state."""
        if self._snap is None:
            self._snap = self._ms.export(0)
        return self._snap

    @property
    def tracks(self) -> Dict[int, TrackView]:
        """Read-only view shaped like the reference's ``self.tracks`` dict (mainTracking.py:45-46, :15-42)."""
        s = self.snapshot()
        return {int(s["ids"][i]): TrackView(s, i) for i in range(len(s["ids"]))}

    @property
    def next_id(self) -> int:
        """The id the next new track gets (mainTracking.py:371-372)."""
        return self.snapshot()["next_id"]

    def _rows(self, row_to_tid):
        s = self.snapshot()
        pos = {int(t): i for i, t in enumerate(s["ids"])}
        return s, [pos[int(t)] for t in row_to_tid]

    # -- operator-level queries over the device state (read-only) -----------------------------------
    @torch.no_grad()
    def build_C_app_topk(self, *, row_to_tid: List[int], det_embs: List[np.ndarray], device=None, topk: int = 5,
                         use_topk_mean: bool = True, fallback_to_ema: bool = True) -> torch.Tensor:
        """mainTracking.py:141-211."""
        M, N = len(row_to_tid), len(det_embs)
        if M == 0 or N == 0:
            return torch.zeros((M, N), device=self.device)
        s, rows = self._rows(row_to_tid)
        det = torch.from_numpy(np.stack([np.asarray(e, dtype=np.float32).reshape(-1) for e in det_embs])).to(self.device)
        bank = torch.from_numpy(s["bank"][rows]).to(self.device)
        lens = torch.from_numpy(s["bank_len"][rows]).to(self.device)
        fb = torch.from_numpy(s["ema"][rows]).to(self.device) if fallback_to_ema else None
        return cost_ops.app_cost_topk(bank, lens, det, topk=topk, use_topk_mean=use_topk_mean, fallback=fb)

    @torch.no_grad()
    def cal_cost(self, *, row_to_tid, det_embs, det_boxes, det_confs, input_hw, device=None, assign=None):
        """mainTracking.py:213-303."""
        s, rows = self._rows(row_to_tid)
        for e in det_embs:
            if np.asarray(e).reshape(-1).shape[0] != 128:
                raise ValueError("det_embs must be 128D")
        C_app = self.build_C_app_topk(row_to_tid=row_to_tid, det_embs=det_embs, topk=self.emb_top_k)
        return cost_ops.cal_cost(C_app=C_app, boxes_prev=s["last_bbox"][rows].tolist(), boxes_cur=det_boxes,
                                 input_hw=input_hw, conf_prev=s["last_conf"][rows].tolist(), conf_cur=det_confs,
                                 w_app=self.w_app, w_bbox=self.w_bbox, w_conf=self.w_conf, alpha=self.alpha,
                                 beta=self.beta, assign=assign, unmatch_cost=self.unmatch_cost)

    def apply_kalman_gating(self, C_total_np: np.ndarray, row_to_tid: List[int], det_boxes, *, maha_thr: float = 13.28,
                            INF: float = 1e9) -> np.ndarray:
        """mainTracking.py:306-338: gates ``C_total_np`` in place and returns it."""
        M, N = C_total_np.shape
        if M == 0 or N == 0:
            return C_total_np
        s, rows = self._rows(row_to_tid)
        bk = kalman_ops.BatchedKalman.__new__(kalman_ops.BatchedKalman)
        bk.device, bk.M = self.device, M
        bk.x = torch.from_numpy(s["x"][rows]).to(self.device)
        bk.P = torch.from_numpy(s["P"][rows]).to(self.device)
        bk.stage = torch.from_numpy(s["stage"][rows]).to(self.device)
        bk.r = torch.ones(4, dtype=torch.float32, device=self.device)
        d2 = bk.maha(det_boxes).cpu().numpy()
        C_total_np[d2 > float(maha_thr)] = float(INF)
        return C_total_np
