"""Per-frame wall time of the drop-in API on one stream of config 2 (64 detections): Tracking.update(obj) with the
reference's list-of-arrays obj, Tracking.update_arrays with numpy arrays, and roi_align_from_input_boxes + update."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import alufe_b200
from alufe_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene = synth.Scene(0, n, 1280, 1280)
frames = [scene.step() for _ in range(300)]
feat = torch.randn((1, 512, 40, 40), device="cuda")
out = {}
for mode in ("update", "update_arrays", "roi+update"):
    trk = alufe_b200.Tracking(conf=alufe_b200.SHIPPED_CONF, max_tracks=256, max_dets=max(64, n))
    ts = []
    for f, o in enumerate(frames):
        if mode == "update_arrays":
            b, c, e = np.asarray(o["bboxes"], np.float64), np.asarray(o["confs"], np.float64), np.stack(o["embs"]).astype(np.float32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "update":
            trk.update(o)
        elif mode == "update_arrays":
            trk.update_arrays(b, c, e, f)
        else:
            p = alufe_b200.roi_align_from_input_boxes(feat, o["bboxes"], (1280, 1280), out_size=(10, 10))
            trk.update(o)
            torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    out[mode] = round(1e6 * float(np.median(ts[60:])), 1)
print(json.dumps({"detections": n, "us_per_frame_median": out}))
