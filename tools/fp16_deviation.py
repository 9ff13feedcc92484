"""How far the float16 ROI Align path (float32 sampling of the half-stored map, one rounding at the end) is from
torchvision's own half kernel (which also computes sample positions in half precision), and how far each is from the float32
result of the half-rounded inputs.  BASELINE config 2 shape, 10x10, aligned, sampling_ratio 2."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torchvision.ops import roi_align as tv_roi
import alufe_b200
from alufe_b200 import synth

torch.cuda.set_device(0)
rng = np.random.default_rng(5)
out = {}
for name, (Hf, Wf, H_in, W_in, n) in (("c2", synth.CONFIGS["c2"]), ("c5", synth.CONFIGS["c5"])):
    feat = torch.from_numpy(synth.feature_map(11, 1, 512, Hf, Wf)).cuda().half()
    boxes = synth.random_boxes(rng, n, H_in, W_in).astype(np.float16)
    rois16 = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.float16), boxes], 1)).cuda()
    scale = Hf / float(H_in)
    ours = alufe_b200.roi_align(feat, rois16.float(), (10, 10), scale, 2, True).float()
    tv16 = tv_roi(feat, rois16, (10, 10), scale, 2, True).float()
    ref32 = tv_roi(feat.float(), rois16.float(), (10, 10), scale, 2, True)        # float32 arithmetic on the same half-rounded inputs
    ulp = lambda a, b: ((a - b).abs() / torch.clamp(ref32.abs(), min=2.0 ** -14) / 2.0 ** -10)   # in half ulps of the reference value
    out[name] = {
        "max_abs_ours_vs_tv_half": float((ours - tv16).abs().max()),
        "max_abs_ours_vs_f32": float((ours - ref32).abs().max()),
        "max_abs_tv_half_vs_f32": float((tv16 - ref32).abs().max()),
        "rel_ulps_ours_vs_f32_p999": float(torch.quantile(ulp(ours, ref32).flatten()[:4_000_000], 0.999)),
        "rel_ulps_tv_half_vs_f32_p999": float(torch.quantile(ulp(tv16, ref32).flatten()[:4_000_000], 0.999)),
        "value_scale_max_abs": float(ref32.abs().max()),
    }
print(json.dumps(out, indent=1))
