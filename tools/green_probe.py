"""Experiment: stream-group step with the association chain on a private SM slice (green context)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import alufe_b200
from alufe_b200 import sched

S, W, K = 64, 40, 100
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = {}
for n_small in (0, 8, 16, 24, 32):
    g = bench.StreamGroup(S, W + K, 0, dev)
    if n_small:
        part = sched.SmPartition(n_small)
        g.sA, g.sB = part.big, part.small
        label = "small=%d big=%d" % (part.n_small, part.n_big)
    else:
        label = "two priority streams"
    g.run(0, W)
    ms, _ = g.run(W, K)
    last = g.results[W + K - 1].cpu().numpy()
    out[label] = {"us_per_step": round(ms / K * 1e3, 1), "ok": bool((last[:, 5] == 0).all() and (last[:, 0] > 0).all())}
    del g
    torch.cuda.empty_cache()
print(json.dumps(out))
