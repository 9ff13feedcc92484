#!/usr/bin/env python
"""Headline benchmark: tracked frames/sec of the per-frame hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

The frame is BASELINE.json config 2 (the config the metric is quoted on): ROI Align of a
[1,512,40,40] map at 64 boxes -> [64,512,10,10], then the association step (predict, cost + gate,
assignment, update, births/purge) for 64 detections against ~64 tracks.  A GPU owns a GROUP of
--streams independent video streams (BASELINE.json north_star: "one stream group per GPU") and one
step advances every stream of the group by one frame with batched launches, so frames per step =
streams.  extra.single_stream reports the same path with a single stream (one frame in flight).
Synthetic inputs follow SURVEY.md section 8d.  Prints ONE JSON line (rank 0).

* value     : frames/s with every input already in HBM (device-resident maps, rois, detections),
              timed with CUDA events over exactly K steps, max over ranks.
* e2e       : frames/s through the public Python API with HOST (pinned) buffers: per step the map,
              rois and detections are copied up and the match table is copied back.
* roofline  : ROI Align, algorithmic bytes / average launch duration measured with CUDA events
              inside the timed region, against the measured copy bandwidth (MEASURED_PEAKS.json).
* cpu_baseline : the oracle port of the reference path (torchvision CPU roi_align +
              Tracking.update restated, single-threaded Python like the reference) on a bounded
              sample of the same frames, on this box's host cores.
--impl reference times that CPU path alone.  N > 1 (torchrun): one stream group per GPU (weak
scaling), per-step NCCL all-gather of the result tables, value = total frames / max time.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tracked frames/sec at N=64 boxes; ROIAlign achieved HBM GB/s vs B200 peak"
WORKLOAD = ("c2: 1280x1280 frame, map [1,512,40,40], 64 detections vs ~64 tracks, roi_align 10x10 + "
            "cost + gate + Kalman + assignment per frame")
C, HF, WF, H_IN, W_IN, NBOX, PS = 512, 40, 40, 1280, 1280, 64, 10
ROI_ALG_BYTES = NBOX * C * PS * PS * 4 + C * HF * WF * 4 + NBOX * 20
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (default)
    "c2": (40, 40, 1280, 1280, 64, "c2: 1280x1280 frame, map [1,512,40,40], 64 detections vs ~64 tracks, roi_align 10x10 + "
                                   "cost + gate + Kalman + assignment per frame"),
    # BASELINE.json configs[4]: the streams of the multi-GPU configuration
    "c5": (34, 60, 1088, 1920, 128, "c5: 1088x1920 frame, map [1,512,34,60], 128 detections vs ~128 tracks, roi_align "
                                    "10x10 + cost + gate + Kalman + assignment per frame"),
}


def select_workload(name):
    """Sets the module-level shape constants; everything below reads them at call time."""
    global HF, WF, H_IN, W_IN, NBOX, WORKLOAD, ROI_ALG_BYTES
    HF, WF, H_IN, W_IN, NBOX, WORKLOAD = WORKLOADS[name]
    ROI_ALG_BYTES = NBOX * C * PS * PS * 4 + C * HF * WF * 4 + NBOX * 20


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_frames(seed, n_frames):
    """Detections of n_frames consecutive frames as dense arrays + the reference's obj dicts."""
    from alufe_b200 import synth
    scene = synth.Scene(seed, NBOX, H_IN, W_IN)
    objs = [scene.step() for _ in range(n_frames)]
    boxes = np.array([o["bboxes"] for o in objs], dtype=np.float64)
    confs = np.array([o["confs"] for o in objs], dtype=np.float64)
    embs = np.array([np.stack(o["embs"]) for o in objs], dtype=np.float32)
    rois = np.concatenate([np.zeros((n_frames, NBOX, 1)), boxes], axis=2).astype(np.float32)
    return objs, boxes, confs, embs, rois


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                     "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.01)

    def summary(self):
        med = float(np.median(self.sm)) if self.sm else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm)}


def _cpu_worker(args):
    """One host process = one video stream through the reference path (module-level for spawn)."""
    seed, warm, max_frames, budget_s, workload = args
    select_workload(workload)                # spawned processes start from the defaults
    import torch
    torch.set_num_threads(1)
    import alufe_b200  # noqa: F401
    from alufe_b200 import synth
    from oracle import tracker_ref, native
    try:
        from torchvision.ops import roi_align as tv_roi
    except Exception:                                              # noqa: BLE001
        tv_roi = None
    objs, *_ = make_frames(seed, warm + max_frames)
    feat_cpu = synth.feature_map(seed, 1, C, HF, WF)
    f_t = torch.from_numpy(feat_cpu)
    ref = tracker_ref.TrackerRef(tracker_ref.SHIPPED_CONF)
    for obj in objs[:warm]:
        ref.update(obj)                       # banks fill and Kalman dtypes reach steady state
    n, t_roi, t_upd, t_start = 0, 0.0, 0.0, time.perf_counter()
    for obj in objs[warm:]:
        t0 = time.perf_counter()
        rois = np.array([[0.0] + list(b) for b in obj["bboxes"]], dtype=np.float32)
        if tv_roi is not None:
            tv_roi(f_t, torch.from_numpy(rois), (PS, PS), HF / float(H_IN), 2, True)
        else:
            native.roi_align(feat_cpu, rois, (PS, PS), HF / float(H_IN), 2, True)
        t1 = time.perf_counter()
        ref.update(obj)
        t2 = time.perf_counter()
        t_roi, t_upd, n = t_roi + (t1 - t0), t_upd + (t2 - t1), n + 1
        if t2 - t_start > budget_s:
            break
    return n, t_roi, t_upd, time.perf_counter() - t_start, ("torchvision" if tv_roi else "oracle C")


def cpu_group_fps(warm, max_frames, budget_s, workers=None):
    """The reference path on the host cores: the reference is single-threaded Python per stream
    (torchvision's CPU roi_align does not scale with threads either), so independent
    streams are spread over one process per core.  Returns a cpu_baseline dict."""
    import multiprocessing as mp
    workers = workers or max(1, os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        name = [k for k, v in WORKLOADS.items() if v[5] == WORKLOAD][0]
        out = pool.map(_cpu_worker, [(50000 + w, warm, max_frames, budget_s, name) for w in range(workers)])
    fps = sum(n / wall for n, _, _, wall, _ in out)
    n_tot = sum(o[0] for o in out)
    ms_roi = 1e3 * sum(o[1] for o in out) / n_tot
    ms_upd = 1e3 * sum(o[2] for o in out) / n_tot
    return {"value": fps, "unit": "frames/s", "cores": workers, "kind": "port",
            "sample": ("%d independent " + name + " streams, one host process each (%d logical CPUs), %d frames in total "
                       "after %d warm-up frames per stream: roi_align (%s CPU) %.1f ms + Tracking.update port %.1f ms per "
                       "frame per core; the association step is single-threaded Python as in the reference")
                      % (workers, os.cpu_count() or 0, n_tot, warm, out[0][4], ms_roi, ms_upd)}


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL) may print to stdout; the driver wants exactly one JSON line there.  Everything
    else is sent to stderr and emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference is pure Python and its arithmetic lives in torchvision / filterpy / scipy), on all
    host cores, one stream per core, for the same per-frame workload as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # like the GPU arm: the workload is defined with full history banks (30 frames); warm-up steps the caller
    # did not ask for run as untimed set-up before the W warm-up steps
    setup = max(0, 30 - args.warmup)
    warm = setup + args.warmup
    frames = max(1, min(args.steps, 400))
    cpu = cpu_group_fps(warm, frames, budget_s=120.0)
    fps = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": frames,
        "warmup": args.warmup, "setup_steps": setup, "ms_per_step": 1e3 * cpu["cores"] / fps, "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d independent streams (one per host core), each " % cpu["cores"] + WORKLOAD,
                   "frames_per_step": cpu["cores"]},
        "cpu_baseline": cpu,
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


class StreamGroup:
    """S independent config-2 streams on one GPU with every input resident in HBM.

    One step = one frame of every stream: a single ROI Align launch over the S maps (K = 64*S ROIs,
    batch index = stream) on stream A, and one MultiStreamTracker step on stream B that waits for
    that frame's ROI launch (in the real pipeline the encoder sits between them), so ROI Align of
    frame t+1 overlaps the association of frame t."""

    def __init__(self, S, n_frames, rank, dev, channels_last=False):
        import torch
        import alufe_b200
        from alufe_b200 import _lib
        self.torch, self.lib, self._lib, self.S, self.F, self.dev = torch, _lib.lib(), _lib, S, n_frames, dev
        per = [make_frames(1000 * rank + s, n_frames) for s in range(S)]
        self.objs0 = per[0][0]
        self.boxes = np.stack([p[1] for p in per], axis=1)          # [F, S, 64, 4]
        self.confs = np.stack([p[2] for p in per], axis=1)
        self.embs = np.stack([p[3] for p in per], axis=1)
        rois = np.stack([p[4] for p in per], axis=1)                # [F, S, 64, 5]
        rois[..., 0] = np.arange(S, dtype=np.float32)[None, :, None]
        self.rois = rois.reshape(n_frames, S * NBOX, 5)
        self.d_boxes, self.d_confs = torch.from_numpy(self.boxes).to(dev), torch.from_numpy(self.confs).to(dev)
        self.d_embs, self.d_rois = torch.from_numpy(self.embs).to(dev), torch.from_numpy(self.rois).to(dev)
        self.d_ndet = torch.full((n_frames, S), NBOX, dtype=torch.int32, device=dev)
        self.d_frame = torch.arange(n_frames, dtype=torch.int32, device=dev)[:, None].repeat(1, S).contiguous()
        # enough distinct maps / output buffers that nothing is served from the 126 MB L2
        self.nmap = max(2, -(-160 // max(1, (S * C * HF * WF * 4) // 1000000)))
        self.nout = max(2, -(-260 // max(1, (S * NBOX * C * PS * PS * 4) // 1000000)))
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        self.nhwc = 1 if channels_last else 0          # same values; only the memory order of each map differs
        self.maps = torch.randn((self.nmap, S, HF, WF, C) if channels_last else (self.nmap, S, C, HF, WF),
                                device=dev, generator=gen)
        self.outs = torch.empty((self.nout, S * NBOX, C, PS, PS), device=dev)
        self.trk = alufe_b200.MultiStreamTracker(S, alufe_b200.SHIPPED_CONF, max_tracks=256, max_dets=NBOX, device=dev)
        self.results = torch.zeros((n_frames, S, self.trk.stride), dtype=torch.int32, device=dev)
        # the association chain is short and latency-bound: it gets the high-priority stream so its CTAs are
        # scheduled ahead of the queued ROI Align tiles of the next frame
        self.sA, self.sB = torch.cuda.Stream(dev, priority=0), torch.cuda.Stream(dev, priority=-1)
        self.roi_done = [torch.cuda.Event() for _ in range(8)]

        self.map_b, self.out_b = S * C * HF * WF * 4, S * NBOX * C * PS * PS * 4
        self.roi_alg_bytes = S * NBOX * C * PS * PS * 4 + S * C * HF * WF * 4 + S * NBOX * 20
        self.ptr = {k: getattr(self, k).data_ptr() for k in
                    ("maps", "outs", "d_rois", "d_ndet", "d_boxes", "d_confs", "d_embs", "d_frame", "results")}

    def roi(self, i):
        P, S, p = ctypes.c_void_p, self.S, self.ptr
        rc = self.lib.b200_roi_align_fwd_f32(P(p["maps"] + (i % self.nmap) * self.map_b), self.nhwc, S, C, HF, WF,
                                             P(p["d_rois"] + i * S * NBOX * 20), S * NBOX, PS, PS, HF / float(H_IN), 2, 1,
                                             P(p["outs"] + (i % self.nout) * self.out_b), P(self.sA.cuda_stream))
        if rc:
            self._lib.check(rc)

    def assoc(self, i):
        P, S, p = ctypes.c_void_p, self.S, self.ptr
        rc = self.lib.b200_tracker_step(self.trk._h, P(p["d_ndet"] + 4 * i * S), P(p["d_boxes"] + i * S * NBOX * 32),
                                        P(p["d_confs"] + i * S * NBOX * 8), P(p["d_embs"] + i * S * NBOX * 512),
                                        P(p["d_frame"] + 4 * i * S), P(p["results"] + i * S * self.trk.stride * 4),
                                        P(self.sB.cuda_stream))
        if rc:
            self._lib.check(rc)

    def step(self, i, probe=None):
        if probe is not None:
            # roofline probe: this ROI Align launch runs with no association kernel beside it, so the
            # events bracket the kernel alone (the other steps overlap it with the previous frame)
            self.sA.wait_stream(self.sB)
            probe[0].record(self.sA)
        self.roi(i)
        if probe is not None:
            probe[1].record(self.sA)
        ev = self.roi_done[i % len(self.roi_done)]
        ev.record(self.sA)
        self.sB.wait_event(ev)
        self.assoc(i)

    def run(self, first, count, n_probe=0, after_step=None):
        """Runs `count` steps starting at frame `first`; returns (elapsed ms on the device, probe times us)."""
        torch = self.torch
        main = torch.cuda.current_stream(self.dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        probes = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_probe)]
        every = max(1, count // max(1, n_probe))
        ev0.record(main)
        self.sA.wait_event(ev0)
        self.sB.wait_event(ev0)
        for k in range(count):
            q = k // every
            self.step(first + k, probes[q] if (n_probe and k % every == 0 and q < n_probe) else None)
            if after_step is not None:
                after_step(first + k)
        main.wait_stream(self.sA)
        main.wait_stream(self.sB)
        ev1.record(main)
        torch.cuda.synchronize(self.dev)
        return ev0.elapsed_time(ev1), [a.elapsed_time(b) * 1e3 for a, b in probes]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--streams", type=int, default=64, help="config-2 streams per GPU stepped together")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2",
                    help="per-stream frame shape: c2 (BASELINE configs[1], the headline) or c5 (configs[4])")
    ap.add_argument("--gather-every", type=int, default=8,
                    help="N > 1: all-gather the result tables every this many frames")
    ap.add_argument("--no-extra", action="store_true", help="skip the single-stream side measurement")
    args = ap.parse_args()
    select_workload(args.workload)
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import alufe_b200
    from alufe_b200 import _lib, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    # Run this rank's host side on the CPUs next to its GPU (NVML's ideal affinity): the e2e leg copies 212 MB per
    # step from pinned host memory, and pinned pages allocated on the far socket halve that copy rate.
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    near_cpus = None
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        near_cpus = len(os.sched_getaffinity(0))
    except Exception:                                               # noqa: BLE001
        pass
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    K, W, S = args.steps, args.warmup, args.streams
    peak, peak_src = measured_peaks()

    # ---- stream group of S config-2 streams on this GPU --------------------------------------------
    # The workload is defined with full history banks (hist_max = 30) and steady-state Kalman dtypes, which
    # takes 35 frames.  If the caller asks for fewer warm-up steps, the difference is run as untimed set-up
    # before the W warm-up steps, so the timed K steps always see the same workload.
    pre = max(0, 35 - W)
    grp = StreamGroup(S, pre + W + K, rank, dev)
    # The only inter-GPU traffic: the per-stream result tables, all-gathered for the consumer of tracking.py:329.
    # They are gathered GATHER_EVERY frames at a time ("optionally gather every K frames"):
    # a NCCL kernel per step holds SM slots while it waits for the slowest rank, which cost 17 % at 8 GPUs.
    GATHER_EVERY = max(1, args.gather_every)
    n_total = pre + W + K
    gathered = (torch.zeros((2, world, GATHER_EVERY, S, grp.trk.stride), dtype=torch.int32, device=dev)
                if world > 1 else None)
    pending = []

    def gather_after(i):
        if world == 1:
            return
        full = (i + 1) % GATHER_EVERY == 0
        if not (full or i == n_total - 1 or i == pre + W - 1):     # also flush at the end of each run() call
            return
        i0 = (i // GATHER_EVERY) * GATHER_EVERY
        n = i + 1 - i0
        torch.cuda.current_stream(dev).wait_stream(grp.sB)         # NCCL orders after the current stream
        dst = gathered[(i // GATHER_EVERY) % 2][:, :n] if n == GATHER_EVERY else \
            torch.empty((world, n, S, grp.trk.stride), dtype=torch.int32, device=dev)
        pending.append(dist.all_gather_into_tensor(dst.reshape(-1) if n == GATHER_EVERY else dst.view(-1),
                                                   grp.results[i0:i0 + n].reshape(-1), async_op=True))
        if len(pending) > 1:
            pending.pop(0).wait()

    grp.run(0, pre + W, after_step=gather_after)
    for h in pending:
        h.wait()
    pending.clear()

    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = lib.b200_launch_count()
    sampler.sample()
    sampler.start()
    elapsed_ms, roi_us = grp.run(pre + W, K, n_probe=min(K, 16), after_step=gather_after)
    for h in pending:
        h.wait()
    torch.cuda.synchronize()
    sampler.sample()
    sampler.stop_flag = True
    launches = lib.b200_launch_count() - launches0
    if world > 1:
        dist.barrier()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    roi_us_avg = float(np.mean(roi_us))
    last = grp.results[pre + W + K - 1].cpu().numpy()
    assert (last[:, 5] == 0).all() and (last[:, 0] > 0).all(), "device path produced no matches"

    # ---- end to end through the public API with host (pinned) buffers --------------------------------
    ms2 = alufe_b200.MultiStreamTracker(S, alufe_b200.SHIPPED_CONF, max_tracks=256, max_dets=NBOX, device=dev)
    NPIN = 2
    pin_maps = torch.randn((NPIN, S, C, HF, WF), generator=torch.Generator().manual_seed(99 + rank)).pin_memory()
    pin_rois = torch.from_numpy(grp.rois).pin_memory()
    feat_dev = [torch.empty((S, C, HF, WF), device=dev) for _ in range(2)]     # double-buffered upload target
    rois_dev = [torch.empty((S * NBOX, 5), device=dev) for _ in range(2)]
    n_det = np.full(S, NBOX, np.int32)
    n_e2e = min(K, 60)
    scale = HF / float(H_IN)
    copy_stream = torch.cuda.Stream(dev)
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):                                   # frame i's map and boxes, host (pinned) -> device
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])      # the ROI launch that read this buffer two frames ago
            feat_dev[b].copy_(pin_maps[i % NPIN], non_blocking=True)
            rois_dev[b].copy_(pin_rois[i % len(pin_rois)], non_blocking=True)
            uploaded[b].record(copy_stream)

    def e2e_step(i):
        """One frame of every stream through the public API.  The upload of frame i+1 is queued on a copy stream
        before the (synchronising) tracker call of frame i, as a caller feeding frames from the host would do."""
        b = i & 1
        main = torch.cuda.current_stream(dev)
        main.wait_event(uploaded[b])
        patches = alufe_b200.roi_align(feat_dev[b], rois_dev[b], (PS, PS), scale, 2, True)
        consumed[b].record(main)
        upload(i + 1)
        return patches, ms2.step(n_det, grp.boxes[i], grp.confs[i], grp.embs[i], np.full(S, i, np.int32))

    for b in range(2):
        consumed[b].record(torch.cuda.current_stream(dev))
    upload(0)
    for i in range(pre + W):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        _, res = e2e_step(pre + W + k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert (res[:, 0] > 0).all()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d = grp.map_b + S * NBOX * 20 + S * (8 + NBOX * (32 + 8 + 512))
    d2h = S * grp.trk.stride * 4
    extra = {}
    # ---- the same public-API loop with the maps already on the device (how the reference's CUDA deployment calls
    # roi_align: the detector's map never leaves the GPU, only boxes / confidences / embeddings come from the host) ----
    if rank == 0 and not args.no_extra and K - n_e2e > 0:          # needs frames beyond those of the e2e leg
        try:
            def api_step(i):
                rois_dev[i & 1].copy_(pin_rois[i % len(pin_rois)], non_blocking=True)
                patches = alufe_b200.roi_align(feat_dev[i & 1], rois_dev[i & 1], (PS, PS), scale, 2, True)
                return patches, ms2.step(n_det, grp.boxes[i], grp.confs[i], grp.embs[i], np.full(S, i, np.int32))
            torch.cuda.synchronize()
            n_api = min(60, K - n_e2e)                 # the frames after those of the e2e leg
            t0 = time.perf_counter()
            for k in range(n_api):
                _, res = api_step(pre + W + n_e2e + k)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert (res[:, 0] > 0).all()
            extra["api_device_maps"] = {"value": S * n_api / dt, "unit": "frames/s", "ms_per_step": 1e3 * dt / n_api,
                                        "h2d_bytes_per_step": S * NBOX * 20 + S * (8 + NBOX * (32 + 8 + 512)),
                                        "note": "alufe_b200.roi_align + MultiStreamTracker.step with host boxes / "
                                                "confidences / embeddings and device-resident maps"}
        except Exception as exc:                                    # noqa: BLE001
            extra["api_device_maps_error"] = repr(exc)
    del pin_maps, feat_dev

    # ---- single-stream latency mode (one frame in flight; the shape the reference runs) --------------
    if rank == 0 and not args.no_extra:
        try:
            K1 = 400
            g1 = StreamGroup(1, pre + W + K1, 7000 + rank, dev)
            g1.run(0, pre + W)
            ms1, roi1 = g1.run(pre + W, K1, n_probe=32)
            extra["single_stream"] = {"value": K1 / (ms1 * 1e-3), "unit": "frames/s", "ms_per_frame": ms1 / K1,
                                      "roi_us_per_launch": float(np.mean(roi1)),
                                      "roi_frac_of_peak": g1.roi_alg_bytes / float(np.mean(roi1)) / 1e3 / peak}
            del g1
        except Exception as exc:                                    # noqa: BLE001
            extra["single_stream_error"] = repr(exc)

    # ---- the same stream group fed channels-last maps (what a channels_last detector would hand over) ----
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            del grp.maps, grp.outs
            torch.cuda.empty_cache()
            Kc = min(K, 100)
            gc = StreamGroup(S, pre + W + Kc, 9000 + rank, dev, channels_last=True)
            gc.run(0, pre + W)
            msc, roic = gc.run(pre + W, Kc, n_probe=min(Kc, 16))
            extra["channels_last_maps"] = {"value": S * Kc / (msc * 1e-3), "unit": "frames/s", "ms_per_step": msc / Kc,
                                           "roi_us_per_launch": float(np.mean(roic)),
                                           "roi_frac_of_peak": gc.roi_alg_bytes / float(np.mean(roic)) / 1e3 / peak,
                                           "kernel": "roi_prep_kernel + roi_align_pipe_kernel<10,10,NHWC>"}
            del gc
        except Exception as exc:                                    # noqa: BLE001
            extra["channels_last_error"] = repr(exc)

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only) ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if affinity0 is not None:
            os.sched_setaffinity(0, affinity0)                       # the CPU baseline uses every host core
        cpu = cpu_group_fps(warm=30, max_frames=60, budget_s=12.0)

    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roi_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch_%d_streams" % S)
        frames = world * S * K
        line = {
            "metric": METRIC, "value": frames / (elapsed_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": W, "setup_steps": pre, "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d independent streams per GPU, each " % S + WORKLOAD, "streams_per_gpu": S,
                       "frames_per_step": S, "roi_out": [PS, PS], "layout": "nchw",
                       "l2": "inputs larger than L2: each step reads %d maps (%.0f MB) and writes %.0f MB; %d map sets and %d "
                             "output buffers rotate" % (S, grp.map_b / 1e6, grp.out_b / 1e6, grp.nmap, grp.nout),
                       "pipeline": "ROI Align of frame t+1 (stream A) overlaps the association of frame t (stream B)",
                       "gather": ("result tables of all ranks all-gathered with NCCL every %d frames" % GATHER_EVERY)
                       if world > 1 else "single GPU: none",
                       "state_dtype": "f64 Kalman/assignment duals, f32 ROI/cost"},
            "clocks": sampler.summary(),
            "e2e": {"value": world * S * n_e2e / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": n_e2e, "api": "alufe_b200.roi_align + MultiStreamTracker.step (pinned host buffers)",
                    "h2d_gbps": h2d * n_e2e / e2e_s / 1e9, "host_cpus_near_gpu": near_cpus},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "roi_prep_kernel + roi_align_multi_kernel<10,10,NCHW> (one ROI Align launch)", "bound": "hbm",
                         "achieved": grp.roi_alg_bytes / roi_us_avg / 1e3, "peak": peak, "unit": "GB/s",
                         "frac": grp.roi_alg_bytes / roi_us_avg / 1e3 / peak, "traffic": traffic,
                         "alg_bytes_per_launch": grp.roi_alg_bytes, "us_per_launch": roi_us_avg, "launches_timed": len(roi_us),
                         "peak_source": peak_src,
                         "note": "the timed launches are %d of the K ROI Align launches of the timed region, bracketed by CUDA "
                                 "events on their stream and run without a concurrent association kernel" % len(roi_us)},
            "cpu_baseline": cpu,
            "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
