"""Experiment: ROI Align of a 64-stream group on stream A; on the high-priority stream B, behind each frame's ROI launch, a
kernel that only WAITS for a fixed time with a chosen footprint (tools/sleeper.cu) in place of the association chain.  Shows
what co-residency itself costs: period of the combined loop vs ROI Align alone."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
sl = ctypes.CDLL(os.path.join(ROOT, "tools", "build", "libsleeper.so"))
sl.sleeper_launch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
W, K = 20, 80
g = bench.StreamGroup(bench.WORKLOADS["c2"], 64, W + K, 0, dev)
cases = [("no kernel on stream B", None),
         ("148 x 32 thr, 60 us", [(148, 32, 0, 60000)]),
         ("148 x 1024 thr (32K regs/SM), 60 us", [(148, 1024, 0, 60000)]),
         ("296 x 1024 thr (all regs), 60 us", [(296, 1024, 0, 60000)]),
         ("148 x 32 thr + 100 KB smem, 60 us", [(148, 32, 100 * 1024, 60000)]),
         ("148 x 32 thr + 20 KB smem, 60 us", [(148, 32, 20 * 1024, 60000)]),
         ("6 kernels of 148 x 32 thr, 10 us each", [(148, 32, 0, 10000)] * 6),
         ("6 kernels of 296 x 1024 thr, 10 us each", [(296, 1024, 0, 10000)] * 6),
         ("148 x 32 thr, 150 us", [(148, 32, 0, 150000)]),
         ("148 x 1024 thr, 150 us", [(148, 1024, 0, 150000)])]
cases = [(n + "  [default carve-out]", k, -1) for n, k in cases[:4]] + \
        [(n + "  [max-shared carve-out]", k, 100) for n, k in cases[1:4]] + \
        [(n + "  [default carve-out]", k, -1) for n, k in cases[4:]] + \
        [("296 x 128 thr, 60 us  [max-shared carve-out]", [(296, 128, 0, 60000)], 100),
         ("148 x 256 thr, 120 us  [max-shared carve-out]", [(148, 256, 0, 120000)], 100)]
for name, ks, carve in cases:
    sl.sleeper_carveout(carve)
    def assoc(i, ks=ks):
        if ks is None:
            return
        for (grid, block, smem, ns) in ks:
            sl.sleeper_launch(grid, block, smem, ns, ctypes.c_void_p(g.sB.cuda_stream))
    g.assoc = assoc
    g.timed(0, W)
    ms, _ = g.timed(W, K)
    print("  %-72s %7.1f us per step" % (name, ms / K * 1e3), flush=True)
