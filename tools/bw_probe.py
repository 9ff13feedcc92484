"""Write-only / read-only / copy bandwidth of this GPU (context for the ROI Align roofline)."""
import json
import torch

n = 1 << 30  # bytes
a = torch.empty(n // 4, device="cuda")
b = torch.empty(n // 4, device="cuda")


def t(fn, it=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    return min(s.elapsed_time(e) for s, e in ev) * 1e-3


out = {"write_only_GBps": n / t(lambda: a.fill_(1.0)) / 1e9,
       "copy_rw_GBps": 2 * n / t(lambda: b.copy_(a)) / 1e9,
       "read_only_GBps": n / t(lambda: a.sum()) / 1e9}
print(json.dumps(out))
