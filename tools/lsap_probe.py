"""Times the assignment operator (b200_lsap_f32) with CUDA events: batch of 64 problems per launch.
Matrix kinds: 'planted' (one clearly best column per row: every search is the known-first-step case),
'clash' (30 % of the rows want an already wanted column), 'random' (U[0,2)), 'gated' (random, 90 % = 1e9)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import alufe_b200
from alufe_b200 import hung, synth

def make(kind, n, rng):
    if kind in ("planted", "clash"):
        C = np.full((n, n), 1e9, np.float32)
        mask = rng.random((n, n)) < 0.03
        C[mask] = rng.uniform(0.5, 2.0, int(mask.sum()))
        perm = rng.permutation(n)
        if kind == "clash":
            bad = rng.random(n) < 0.3
            perm[bad] = rng.choice(perm, int(bad.sum()))
        C[np.arange(n), perm] = rng.uniform(0.05, 0.3, n)
        return C
    return synth.lsap_matrix(rng, n, n, 0.9 if kind == "gated" else 0.0)

rng = np.random.default_rng(0)
for n in (64, 128, 256, 512):
    B = 64 if n <= 128 else 8
    for kind in ("planted", "clash", "gated", "random"):
        C = torch.from_numpy(np.stack([make(kind, n, rng) for _ in range(B)])).cuda()
        for _ in range(3):
            hung.lsap_batched(C, 50.0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(10):
            hung.lsap_batched(C, 50.0)
        ev[1].record()
        torch.cuda.synchronize()
        print(json.dumps({"n": n, "batch": B, "kind": kind, "us_per_launch": round(ev[0].elapsed_time(ev[1]) * 100, 1)}), flush=True)
