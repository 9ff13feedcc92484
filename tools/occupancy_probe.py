"""Experiment: persistent ROI Align grid with N CTAs per SM -- ROI alone and overlapped with the association chain."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from alufe_b200 import _lib

S, W, K = 64, 40, 100
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = {}
for cps in (0, 6, 5, 4, 3):
    _lib.lib().b200_roi_align_set_ctas_per_sm(cps)
    row = {}
    for mode in ("roi_only", "overlap"):
        g = bench.StreamGroup(S, W + K, 0, dev)
        if mode == "roi_only":
            g.assoc = lambda i: None
        g.run(0, W)
        ms, _ = g.run(W, K)
        row[mode] = round(ms / K * 1e3, 1)
        del g
        torch.cuda.empty_cache()
    out["ctas_per_sm=%d" % cps] = row
print(json.dumps(out))
