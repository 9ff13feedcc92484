"""BASELINE config 4 (dense crowd: 512 detections x 512 tracks, no ROI): association step time per frame."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import alufe_b200
from alufe_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W, K = 35, 30
trk = alufe_b200.MultiStreamTracker(1, alufe_b200.SHIPPED_CONF, max_tracks=2 * n + 64, max_dets=n)
scene = synth.Scene(0, n, 1280, 1280)
frames = [scene.step() for _ in range(W + K)]
boxes = np.zeros((1, n, 4)); confs = np.zeros((1, n)); embs = np.zeros((1, n, 128), np.float32)
ts = []
for f, o in enumerate(frames):
    boxes[0] = o["bboxes"]; confs[0] = o["confs"]; embs[0] = np.stack(o["embs"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = trk.step([n], boxes, confs, embs, [f])
    ts.append(time.perf_counter() - t0)
print(json.dumps({"n": n, "ms_per_frame_host_api_median": round(1e3 * float(np.median(ts[W:])), 3),
                  "matches_last": int(res[0, 0]), "live": int(res[0, 3])}))
