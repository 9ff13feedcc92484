"""Debug: phase timing of assign_kernel<1> (needs a build with EXTRA=-DB200_TRK_TIMING)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import alufe_b200
from alufe_b200 import synth, _lib
trk = alufe_b200.Tracking(conf=alufe_b200.SHIPPED_CONF, max_tracks=256, max_dets=64)
scene = synth.Scene(0, 64, 1280, 1280)
acc = []
for f in range(60):
    trk.update(scene.step())
    if f >= 40:
        buf = (ctypes.c_longlong * 32)()
        _lib.lib().b200_debug_timing(buf)
        t = np.array(buf[:6], dtype=np.int64)
        acc.append(np.diff(t))
a = np.array(acc)
names = ["stage+LSAP", "compactions", "KF update (phase A)", "EMA/bank (phase B)", "tail"]
for n, v in zip(names, np.median(a, axis=0)):
    print("%-22s %8.0f cycles  %6.1f us @1.965GHz" % (n, v, v / 1965.0))
