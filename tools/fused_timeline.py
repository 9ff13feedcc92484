"""Debug (B200TRACK_LIB = a -DB200_TRK_TIMING build): global-timer stamps of the phases of the fused two-launch step
(front_kernel / back_kernel, CTA 0 of stream 0), median over steady-state frames; association only, one sync per frame."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from alufe_b200 import _lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
WL = sys.argv[2] if len(sys.argv) > 2 else "c2"
W, K = 40, 60
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.lib()
g = bench.StreamGroup(bench.WORKLOADS[WL], S, W + K, 0, dev, with_roi=False)
g.timed(0, W)
rows, spans = [], []
sp = (ctypes.c_ulonglong * 128)()
for k in range(K):
    lib.b200_debug_spans_trk(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(g.sB)
    g.assoc(W + k)
    e1.record(g.sB)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    lib.b200_debug_timing(buf)
    lib.b200_debug_spans_trk(sp, 0)
    a = np.array(sp[:96], dtype=np.float64).reshape(8, 6, 2)[(W + k) & 7]
    t = np.array(buf[16:25], dtype=np.float64)
    rows.append(np.concatenate([(t - t[0]) / 1e3, [(a[0, 1] - t[0]) / 1e3, (a[2, 0] - t[0]) / 1e3, (a[2, 1] - t[0]) / 1e3,
                                                   e0.elapsed_time(e1) * 1e3]]))
m = np.median(np.array(rows), axis=0)
names = ["front start", "front: rows split", "front: dets prepped + predicted", "front: CTA 0 cost rows done",
         "back: CTA 0 starts work", "back: stage-1 assignment done", "back: ReID cost done", "back: stage 2 / births / purge done",
         "back: updates done", "front_kernel ends (all CTAs)", "back_kernel first CTA resident", "back_kernel ends",
         "events around the step (us)"]
print("streams %d workload %s (us after front_kernel's first instruction, medians of %d frames)" % (S, WL, K))
for n, v in zip(names, m):
    print("  %-40s %8.2f" % (n, v))
