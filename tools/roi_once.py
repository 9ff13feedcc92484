"""Three ROI Align launches of the bench.py stream-group shape (64 maps [512,40,40], 4 096 ROIs, 10x10) for ncu:
    python tools/roi_once.py nchw|nhwc [cl_out]
    ncu --set full --clock-control none --import-source on -k regex:roi_align_ --launch-skip 2 -c 1 ... same command"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import alufe_b200
from alufe_b200 import roi, synth

nhwc = len(sys.argv) > 1 and sys.argv[1] == "nhwc"
cl_out = len(sys.argv) > 2 and sys.argv[2] == "cl_out"
rng = np.random.default_rng(0)
boxes = np.concatenate([synth.random_boxes(rng, 64, 1280, 1280) for _ in range(64)])
rois = np.concatenate([np.repeat(np.arange(64), 64)[:, None].astype(np.float64), boxes], 1).astype(np.float32)
r = torch.from_numpy(rois).cuda()
feat = torch.randn((64, 512, 40, 40), device="cuda")
if nhwc:
    feat = feat.contiguous(memory_format=torch.channels_last)
for _ in range(3):
    out = roi.roi_align(feat, r, (10, 10), 40 / 1280.0, 2, True, out_channels_last=cl_out)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.float().abs().mean()))
