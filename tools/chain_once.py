"""Forty association steps of a 64-stream c2 group (no ROI Align), for ncu:
    ncu --set full --clock-control none --import-source on -k regex:"begin_kernel|cost1_|assign_kernel|cost2_kernel|update_kernel" -s 228 -c 6 ... same command"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
g = bench.StreamGroup(bench.WORKLOADS["c2"], S, 40, 0, dev, with_roi=False)
ms, _ = g.timed(0, 40)
last = g.results[39].cpu().numpy()
print("ok", round(ms / 40 * 1e3, 1), "us per step,", int(last[:, 0].sum()), "matches in the last frame")
