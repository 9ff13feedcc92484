"""Host-side cost of the public API calls of one 64-stream step (wall microseconds per call while the GPU runs behind)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import alufe_b200

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sh = bench.WORKLOADS["c2"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
try:                                     # host side on the CPUs next to the GPU, like bench.py
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(0))
except Exception as exc:
    print("no affinity:", exc)
F = 90
grp = bench.StreamGroup(sh, S, F, 0, dev)
ms = alufe_b200.MultiStreamTracker(S, alufe_b200.SHIPPED_CONF, max_tracks=512, max_dets=sh.NBOX, device=dev)
hb = torch.from_numpy(grp.boxes).pin_memory().numpy()
hc = torch.from_numpy(grp.confs).pin_memory().numpy()
he = torch.from_numpy(grp.embs).pin_memory().numpy()
pin_rois = torch.from_numpy(grp.rois).pin_memory()
feat = grp.maps[0]
rois_dev = [torch.empty((S * sh.NBOX, 5), device=dev) for _ in range(2)]
n_det = np.full(S, sh.NBOX, np.int32)
sA, sB = torch.cuda.Stream(dev, priority=0), torch.cuda.Stream(dev, priority=-1)
evs = [torch.cuda.Event() for _ in range(4)]
acc = {k: [] for k in ("copy_rois", "roi_align", "events", "step_async", "result", "total")}
prev = None
# ---- wall time per step of reduced loops: which part of the step sets the period? ----
def loop(roi, trk, lag, copy_rois=True, n=60):
    import collections
    q = collections.deque()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        if roi:
            with torch.cuda.stream(sA):
                if copy_rois:
                    rois_dev[i & 1].copy_(pin_rois[i], non_blocking=True)
                p = alufe_b200.roi_align(grp.maps[i % grp.nmap], rois_dev[i & 1], (10, 10), sh.HF / float(sh.H_IN), 2, True)
                evs[i & 3].record(sA)
        if trk:
            with torch.cuda.stream(sB):
                if roi:
                    sB.wait_event(evs[i & 3])
                q.append(ms.step_async(n_det, hb[i], hc[i], he[i], np.full(S, 1000 + i, np.int32), pinned=True))
            if len(q) > lag:
                q.popleft().result()
    while q:
        q.popleft().result()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
for i in range(40):                      # banks full
    ms.step(n_det, hb[i], hc[i], he[i], np.full(S, i, np.int32))
rois_dev[0].copy_(pin_rois[0]); rois_dev[1].copy_(pin_rois[1])
for name, kw in (("roi only", dict(roi=True, trk=False, lag=0)), ("roi only, rois resident", dict(roi=True, trk=False, lag=0, copy_rois=False)),
                 ("tracker only, lag 0", dict(roi=False, trk=True, lag=0)), ("tracker only, lag 2", dict(roi=False, trk=True, lag=2)),
                 ("both, lag 1", dict(roi=True, trk=True, lag=1)), ("both, lag 2", dict(roi=True, trk=True, lag=2)),
                 ("both, lag 3", dict(roi=True, trk=True, lag=3))):
    loop(**kw, n=10)
    print("  %-28s %8.1f us per step" % (name, loop(**kw)))
ms.reset()
for i in range(F):
    t0 = time.perf_counter()
    with torch.cuda.stream(sA):
        rois_dev[i & 1].copy_(pin_rois[i], non_blocking=True)
        t1 = time.perf_counter()
        patches = alufe_b200.roi_align(feat, rois_dev[i & 1], (10, 10), sh.HF / float(sh.H_IN), 2, True)
        t2 = time.perf_counter()
        evs[i & 3].record(sA)
    with torch.cuda.stream(sB):
        sB.wait_event(evs[i & 3])
        t3 = time.perf_counter()
        h = ms.step_async(n_det, hb[i], hc[i], he[i], np.full(S, i, np.int32), pinned=True)
        t4 = time.perf_counter()
    if prev is not None:
        prev.result()
    t5 = time.perf_counter()
    prev = h
    if i >= 40:
        for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)):
            acc[k].append(v * 1e6)
prev.result()
torch.cuda.synchronize()
print("streams %d: host microseconds per call (median / max over %d steps)" % (S, len(acc["total"])))
for k, v in acc.items():
    print("  %-12s %8.1f %8.1f" % (k, np.median(v), np.max(v)))
