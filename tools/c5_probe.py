"""BASELINE config 5 on one GPU: S streams (default 64) of 1088x1920 frames, map [1,512,34,60], 128 detections
per frame, through the same StreamGroup as bench.py (ROI Align on one stream, tracker step on the other).
    python tools/c5_probe.py [S] [cl]      cl = channels-last maps"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

bench.HF, bench.WF, bench.H_IN, bench.W_IN, bench.NBOX = 34, 60, 1088, 1920, 128
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
CL = len(sys.argv) > 2 and sys.argv[2] == "cl"
W, K = 40, 60
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = {}
for mode in ("roi_only", "assoc_only", "overlap"):
    g = bench.StreamGroup(S, W + K, 0, dev, channels_last=CL)
    if mode == "roi_only":
        g.assoc = lambda i: None
    if mode == "assoc_only":
        g.roi = lambda i: None
    g.run(0, W)
    ms, _ = g.run(W, K)
    out[mode] = round(ms / K * 1e3, 1)
    if mode == "overlap":
        last = g.results[W + K - 1].cpu().numpy()
        out["matches_per_stream_last_frame"] = float(last[:, 0].mean())
        out["frames_per_s"] = round(S * K / (ms * 1e-3), 1)
    del g
    torch.cuda.empty_cache()
print(json.dumps({"config": "c5", "streams": S, "channels_last_maps": CL, "us_per_step": out}))
