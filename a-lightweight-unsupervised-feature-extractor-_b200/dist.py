"""Multi-GPU plumbing: independent video streams are partitioned across ranks (one process per
GPU); the only traffic is a gather of the small per-stream result tables.

The association step itself never crosses GPUs -- a stream's state lives on exactly one device --
so there is no data-path collective; ``ResultGatherer`` exists because a consumer (the display /
sink process of tracking.py:329) wants every stream's matches in one place.
"""
from typing import List, Optional

import torch
import torch.distributed as dist


def init_process_group_small_footprint(device, **kw):
    """``dist.init_process_group("nccl")`` with collectives limited to ONE CTA (ncclConfig minCTAs = maxCTAs = 1).

    The only collective of this path moves a few hundred kilobytes of result tables; NCCL's default channel count
    would park many CTAs on SMs that ROI Align and the association kernels want, spinning while they wait for the
    slowest rank.  One CTA is plenty for the payload and leaves the SMs to the hot path."""
    opts = None
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.min_ctas = 1
        opts.config.max_ctas = 1
    except Exception:                                           # noqa: BLE001  (older torch: default channels)
        opts = None
    if opts is not None:
        dist.init_process_group("nccl", device_id=device, pg_options=opts, **kw)
    else:
        dist.init_process_group("nccl", device_id=device, **kw)


def stream_owner(stream: int, world: int) -> int:
    """Static placement: stream s lives on rank s mod world."""
    return stream % world


def local_streams(n_streams: int, rank: int, world: int) -> List[int]:
    return [s for s in range(n_streams) if stream_owner(s, world) == rank]


class ResultGatherer:
    """All-gathers per-stream result tables (int32 [n_local, stride]) into global stream order."""

    def __init__(self, n_streams: int, stride: int, device, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_streams, self.stride = n_streams, stride
        self.local = local_streams(n_streams, self.rank, self.world)
        self.per_rank = (n_streams + self.world - 1) // self.world
        self._send = torch.zeros((self.per_rank, stride), dtype=torch.int32, device=device)
        self._recv = torch.zeros((self.world, self.per_rank, stride), dtype=torch.int32, device=device)
        # global stream s sits at [s % world, s // world] of the gathered buffer
        idx = torch.tensor([(s % self.world) * self.per_rank + s // self.world for s in range(n_streams)],
                           dtype=torch.long, device=device)
        self._index = idx
        # CUDA: collectives are enqueued from a private side stream, so the caller's streams never wait for a peer
        self.stream = torch.cuda.Stream(device) if torch.device(device).type == "cuda" else None

    def wait(self, handle):
        """Completes an ``async_op=True`` gather: ``handle`` is the (work, finish) pair; returns the gathered tensor.
        On CUDA the wait is a stream-level dependency of the side stream, not a host block."""
        work, finish = handle
        if work is not None:
            if self.stream is not None:
                with torch.cuda.stream(self.stream):
                    work.wait()
            else:
                work.wait()
        return finish()

    def gather(self, local_results: torch.Tensor, async_op: bool = False):
        """Returns Tensor[n_streams, stride] (or (work, finish) when async_op)."""
        n = len(self.local)
        if local_results.shape != (n, self.stride):
            raise ValueError("expected local results of shape %s" % ((n, self.stride),))
        self._send[:n].copy_(local_results)
        if self.world == 1:
            out = self._send[:n].clone()
            return (None, lambda: out) if async_op else out
        work = dist.all_gather_into_tensor(self._recv.view(-1), self._send.view(-1), group=self.group,
                                           async_op=async_op)
        finish = lambda: self._recv.view(-1, self.stride).index_select(0, self._index)  # noqa: E731
        if async_op:
            return work, finish
        return finish()

    def gather_frames(self, local_results: torch.Tensor, async_op: bool = False, after=None):
        """Several frames per call: ``local_results`` is int32 [F, n_local, stride]; returns
        Tensor[F, n_streams, stride] (or (work, finish) when async_op).  One collective per F frames keeps NCCL
        kernels -- which hold SM slots while they wait for the slowest rank -- off most steps.

        On CUDA the copy into the send buffer and the collective are enqueued on the gatherer's side stream, which
        first waits for ``after`` (a CUDA event recorded where the tables were produced) or, without one, for the
        caller's current stream."""
        n = len(self.local)
        if local_results.dim() != 3 or local_results.shape[1:] != (n, self.stride):
            raise ValueError("expected local results of shape [F, %d, %d]" % (n, self.stride))
        F = local_results.shape[0]
        if self.stream is not None:
            if after is not None:
                self.stream.wait_event(after)
            else:
                self.stream.wait_stream(torch.cuda.current_stream(self._send.device))
            ctx = torch.cuda.stream(self.stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            send = torch.zeros((F, self.per_rank, self.stride), dtype=torch.int32, device=self._send.device)
            send[:, :n].copy_(local_results)
            if self.stream is not None:
                local_results.record_stream(self.stream)
            if self.world == 1:
                out = send[:, :n].clone()
                return (None, lambda: out) if async_op else out
            recv = torch.empty((self.world, F, self.per_rank, self.stride), dtype=torch.int32, device=self._send.device)
            work = dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group, async_op=True)
            if not async_op:
                work.wait()

        def finish():
            flat = recv.permute(1, 0, 2, 3).reshape(F, self.world * self.per_rank, self.stride)
            return flat.index_select(1, self._index)

        if async_op:
            return work, finish
        return finish()
