"""Debug: phase timing of assign_kernel<1> for one stream of n detections (default 64).
Needs a build with `make -C <pkg>/csrc EXTRA=-DB200_TRK_TIMING` (SM-clock stamps at the phase boundaries)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: F401
import alufe_b200
from alufe_b200 import synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
trk = alufe_b200.Tracking(conf=alufe_b200.SHIPPED_CONF, max_tracks=max(256, 2 * n + 64), max_dets=n)
scene = synth.Scene(0, n, 1280, 1280)
acc = []
stats = (ctypes.c_ulonglong * 4)()
for f in range(60):
    if f == 40 and hasattr(_lib.lib(), "b200_debug_lsap_stats"):
        torch.cuda.synchronize()
        _lib.lib().b200_debug_lsap_stats(None, 1)
    trk.update(scene.step())
    if f >= 40:
        buf = (ctypes.c_longlong * 32)()
        _lib.lib().b200_debug_timing(buf)
        t = np.array([buf[0], buf[1], buf[5]], dtype=np.int64)
        d = list(np.diff(t))
        if hasattr(_lib.lib(), "b200_debug_lsap_clk"):        # stamps of the LAST solve_block of the step (stage 2 if it ran)
            c = (ctypes.c_longlong * 8)()
            _lib.lib().b200_debug_lsap_clk(c)
            d += [c[1] - c[0], c[2] - c[1]]
        acc.append(d)
a = np.array(acc)
for name, v in zip(["validate + stage + LSAP", "match lists, misses, leftover dets",
                    "  last solve_block: validate/stage/row minima", "  last solve_block: solver"], np.median(a, axis=0)):
    print("n=%d  %-36s %9.0f cycles  %7.1f us @1.965GHz" % (n, name, v, v / 1965.0))
if hasattr(_lib.lib(), "b200_debug_lsap_stats"):
    _lib.lib().b200_debug_lsap_stats(stats, 0)
    print("n=%d  per frame (both stages): %.1f rows by the known-first-step rule, %.1f full searches, %.1f Dijkstra steps"
          % (n, stats[0] / 20.0, stats[1] / 20.0, stats[2] / 20.0))
