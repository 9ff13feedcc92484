"""Where a host-buffer tracker step spends its time (64 c2 streams): raw DMA times, the kernels behind step_async on the
caller's stream (CUDA events), host time of the call, wall time until the result is on the host."""
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
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(0))
except Exception as exc:
    print("no affinity:", exc)
F = 100
grp = bench.StreamGroup(sh, S, F, 0, dev, with_roi=False)
ms = alufe_b200.MultiStreamTracker(S, alufe_b200.SHIPPED_CONF, max_tracks=256, max_dets=sh.NBOX, device=dev)
hb = torch.from_numpy(grp.boxes).pin_memory().numpy()
hc = torch.from_numpy(grp.confs).pin_memory().numpy()
he_t = torch.from_numpy(grp.embs).pin_memory()
he = he_t.numpy()
n_det = np.full(S, sh.NBOX, np.int32)
st = torch.cuda.Stream(dev)
# raw DMA
dst = torch.empty_like(he_t[0], device=dev)
res_d = torch.zeros((S, ms.stride), dtype=torch.int32, device=dev)
res_h = torch.zeros((S, ms.stride), dtype=torch.int32).pin_memory()
for name, fn in (("H2D embeddings %.2f MB" % (he_t[0].numel() * 4 / 1e6), lambda i: dst.copy_(he_t[i], non_blocking=True)),
                 ("D2H result table %.0f KB" % (res_d.numel() * 4 / 1e3), lambda i: res_h.copy_(res_d, non_blocking=True))):
    ts = []
    with torch.cuda.stream(st):
        for i in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); fn(i); b.record(st)
            st.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
    print("  %-34s %7.1f us (median, events)" % (name, np.median(ts[5:])))
rows = []
with torch.cuda.stream(st):
    for i in range(F):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(st)
        h = ms.step_async(n_det, hb[i], hc[i], he[i], np.full(S, i, np.int32), pinned=True)
        b.record(st)
        t1 = time.perf_counter()
        b.synchronize()
        t2 = time.perf_counter()
        h.result()
        t3 = time.perf_counter()
        if i >= 40:
            rows.append([(t1 - t0) * 1e6, a.elapsed_time(b) * 1e3, (t2 - t0) * 1e6, (t3 - t0) * 1e6])
m = np.median(np.array(rows), axis=0)
print("  synchronous step, %d streams (medians of %d):" % (S, len(rows)))
for n, v in zip(["host time inside step_async", "caller's stream: event before -> kernels done", "wall until kernels done",
                 "wall until result() returned"], m):
    print("    %-48s %7.1f us" % (n, v))
