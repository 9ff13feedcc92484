"""Experiment: 64-stream group step decomposed -- ROI alone, association alone, serial, overlapped."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
CL = len(sys.argv) > 2 and sys.argv[2] == "cl"      # channels-last maps
WL = sys.argv[3] if len(sys.argv) > 3 else "c2"
W, K = 40, 100
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = {}
for mode in os.environ.get("MODES", "roi_only,assoc_only,serial,overlap_prio").split(","):
    g = bench.StreamGroup(bench.WORKLOADS[WL], S, W + K, 0, dev, channels_last=CL)
    if mode == "serial":
        g.sB = g.sA
    if mode == "roi_only":
        g.assoc = lambda i: None
    if mode == "assoc_only":
        g.roi = lambda i: None
    g.timed(0, W)
    ms, _ = g.timed(W, K)
    out[mode] = round(ms / K * 1e3, 1)
    del g
    torch.cuda.empty_cache()
print(json.dumps({"workload": WL, "streams": S, "channels_last_maps": CL, "us_per_step": out}))
