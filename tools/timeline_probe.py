"""Debug (build with EXTRA=-DB200_TRK_TIMING): global-timer spans of every kernel over a few steady-state steps."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from alufe_b200 import _lib

S, W = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 40
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = ctypes.CDLL(_lib.LIB_PATH)
names = ["roi", "begin", "cost1", "assign1", "cost2", "assign2", "update"]
for mode in ("serial", "overlap"):
    g = bench.StreamGroup(bench.WORKLOADS["c2"], S, W + 12, 0, dev)
    if mode == "serial":
        g.sB = g.sA
    g.timed(0, W)
    print("==", mode)
    b1, b2 = (ctypes.c_ulonglong * 128)(), (ctypes.c_ulonglong * 128)()
    lib.b200_debug_spans_roi(None, 1)
    lib.b200_debug_spans_trk(None, 1)
    first = W + (8 - W % 8) % 8          # frame index that is a multiple of 8
    g.timed(W, first - W) if first > W else None
    lib.b200_debug_spans_roi(None, 1)
    lib.b200_debug_spans_trk(None, 1)
    g.timed(first, 6)                      # six consecutive steady-state steps, one span slot each
    lib.b200_debug_spans_roi(b1, 0)
    lib.b200_debug_spans_trk(b2, 0)
    roi = np.array(b1[:16], dtype=np.float64).reshape(8, 2)
    trk = np.array(b2[:96], dtype=np.float64).reshape(8, 6, 2)
    # ROI slot order is only known up to a rotation: sort the used slots by start time
    used = sorted([r for r in roi if r[1] > 0], key=lambda r: r[0])
    t0 = min(used[0][0], trk[0, 0, 0])
    for k in range(6):
        r = used[k] if k < len(used) else (0, 0)
        print("  step %d  roi[%7.1f,%7.1f]  " % (k, (r[0] - t0) / 1e3, (r[1] - t0) / 1e3) +
              "  ".join("%s[%7.1f,%7.1f]" % (n, (trk[k, i, 0] - t0) / 1e3, (trk[k, i, 1] - t0) / 1e3)
                        for i, n in enumerate(names[1:])))
    del g
    torch.cuda.empty_cache()
