"""SM partitioning for the stream-group pipeline (CUDA green contexts, driver API via cuda-python).

The association chain of a stream group is a handful of small, latency-bound kernels; ROI Align of
the next frame is one bulk kernel that fills every SM's register file.  Launched on two ordinary
streams the small kernels starve behind the bulk kernel's CTAs (measured: no overlap at all).  A
green context gives the association chain a private slice of SMs and ROI Align the rest, so both
run truly concurrently.  Only launch plumbing lives here; nothing in this file computes anything.
"""
from typing import Optional, Tuple

import torch


class SmPartition:
    """Two CUDA streams bound to disjoint SM sets: ``small`` (n_small SMs) and ``big`` (the rest)."""

    def __init__(self, n_small: int = 16, device: Optional[int] = None):
        from cuda.bindings import driver as cu
        self._cu = cu
        dev_index = torch.cuda.current_device() if device is None else int(device)
        torch.cuda.init()
        with torch.cuda.device(dev_index):
            torch.zeros(1, device="cuda")          # make sure the primary context exists and is current

            def ck(res):
                err, *rest = res
                if err != cu.CUresult.CUDA_SUCCESS:
                    raise RuntimeError("CUDA driver error %s" % err)
                return rest[0] if len(rest) == 1 else rest

            ck(cu.cuInit(0))
            dev = ck(cu.cuDeviceGet(dev_index))
            sm = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
            groups, _, remaining = ck(cu.cuDevSmResourceSplitByCount(1, sm, 0, n_small))
            self.n_small = int(groups[0].sm.smCount)
            self.n_big = int(remaining.sm.smCount)
            self._ctx, self._streams = [], []
            for res in (groups[0], remaining):
                desc = ck(cu.cuDevResourceGenerateDesc([res], 1))
                g = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
                s = ck(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
                self._ctx.append(g)
                self._streams.append(s)
            self.small = torch.cuda.ExternalStream(int(self._streams[0]), device=dev_index)
            self.big = torch.cuda.ExternalStream(int(self._streams[1]), device=dev_index)

    def close(self):
        cu = self._cu
        for s in self._streams:
            cu.cuStreamDestroy(s)
        for g in self._ctx:
            cu.cuGreenCtxDestroy(g)
        self._streams, self._ctx = [], []
