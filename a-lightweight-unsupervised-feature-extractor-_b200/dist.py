"""Multi-GPU plumbing: independent video streams are partitioned across ranks (one process per
GPU); the only traffic is a gather of the small per-stream result tables.

The association step itself never crosses GPUs -- a stream's state lives on exactly one device --
so there is no data-path collective; ``ResultGatherer`` exists because a consumer (the display /
sink process of tracking.py:329) wants every stream's matches in one place.
"""
import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib


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


class PeerResultGatherer:
    """``ResultGatherer.gather_frames`` without a collective: every rank pushes its tables straight into every peer's
    receive ring over NVLink peer memory (``b200_peer_gather_*``: CUDA IPC mappings, 16-byte stores, a flag per cell) and
    collects a sequence number only when it wants to look at it.  No kernel of the producing side ever waits for another
    rank, so nothing sits on SMs while ranks drift, and the transfer runs at link speed (a one-CTA NCCL all-gather moved
    the same 15 MB in ~0.6 ms at 8 GPUs).  One process per GPU on one node; ``torch.distributed`` is used once, to
    exchange the IPC handles.

    ``push_frames(tables, after=event)`` -> sequence number; ``collect(seq)`` -> Tensor[F, n_streams, stride] in global
    stream order (stream s lives on rank s mod world).  Every rank must push every sequence number with the same F."""

    def __init__(self, n_streams: int, stride: int, max_frames: int, device, group: Optional[dist.ProcessGroup] = None,
                 n_slots: int = 4):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device)
        self.n_streams, self.stride, self.max_frames, self.n_slots = n_streams, stride, int(max_frames), n_slots
        self.local = local_streams(n_streams, self.rank, self.world)
        self.per_rank = (n_streams + self.world - 1) // self.world
        self.bytes_per_frame = self.per_rank * stride * 4
        cap = (self.max_frames * self.bytes_per_frame + 15) // 16 * 16
        self._h = ctypes.c_void_p()
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(lib.b200_peer_gather_create(ctypes.byref(self._h), self.rank, self.world, cap, n_slots))
            if self.world > 1:
                mine = (ctypes.c_ubyte * 64)()
                _lib.check(lib.b200_peer_gather_handle(self._h, mine))
                send = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
                recv = torch.empty(64 * self.world, dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(recv, send, group=group)
                handles = bytes(recv.cpu().tolist())
                rc = lib.b200_peer_gather_connect(self._h, handles)
                ok = torch.tensor([1 if rc == 0 else 0], device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all ranks agree before anyone pushes
                if int(ok.item()) == 0:
                    msg = lib.b200_last_error().decode("utf-8", "replace") if rc else "a peer could not map this rank's ring"
                    lib.b200_peer_gather_destroy(self._h)
                    self._h = ctypes.c_void_p()
                    raise _lib.B200Error("peer memory exchange unavailable: " + msg)
        idx = torch.tensor([(s % self.world) * self.per_rank + s // self.world for s in range(n_streams)],
                           dtype=torch.long, device=self.device)
        self._index = idx
        self.stream = torch.cuda.Stream(self.device)
        self._seq = 0
        self._frames = {}

    def push_frames(self, local_results: torch.Tensor, after=None) -> int:
        """local_results: int32 [F, n_local, stride] (F <= max_frames).  Queued on the gatherer's side stream behind
        ``after`` (a CUDA event recorded where the tables were produced) or the caller's current stream."""
        n = len(self.local)
        if local_results.dim() != 3 or local_results.shape[1:] != (n, self.stride) or local_results.shape[0] > self.max_frames:
            raise ValueError("expected local results of shape [F <= %d, %d, %d]" % (self.max_frames, n, self.stride))
        F = local_results.shape[0]
        if after is not None:
            self.stream.wait_event(after)
        else:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            nbytes = (F * self.bytes_per_frame + 15) // 16 * 16
            if n == self.per_rank and local_results.is_contiguous() and local_results.data_ptr() % 16 == 0 \
                    and nbytes == F * self.bytes_per_frame:
                send = local_results                        # pushed from where the tracker wrote it
            else:
                send = torch.zeros(nbytes // 4, dtype=torch.int32, device=self.device)
                send[:F * self.per_rank * self.stride].view(F, self.per_rank, self.stride)[:, :n].copy_(local_results)
            send.record_stream(self.stream)
            seq = self._seq
            _lib.check(_lib.lib().b200_peer_gather_push(self._h, _lib.ptr(send), nbytes, seq,
                                                        ctypes.c_void_p(self.stream.cuda_stream)))
        self._seq += 1
        self._frames[seq] = (F, nbytes)
        return seq

    def collect(self, seq: int) -> torch.Tensor:
        """Tensor[F, n_streams, stride] of sequence number ``seq`` (asynchronous on the side stream; a consumer on another
        stream waits for ``gatherer.stream``).  Sequence numbers must be collected in order, each once."""
        F, nbytes = self._frames.pop(seq)
        with torch.cuda.stream(self.stream):
            recv = torch.empty((self.world, nbytes // 4), dtype=torch.int32, device=self.device)
            _lib.check(_lib.lib().b200_peer_gather_collect(self._h, _lib.ptr(recv), nbytes, seq,
                                                           ctypes.c_void_p(self.stream.cuda_stream)))
            flat = recv[:, :F * self.per_rank * self.stride].view(self.world, F, self.per_rank, self.stride)
            flat = flat.permute(1, 0, 2, 3).reshape(F, self.world * self.per_rank, self.stride)
            return flat.index_select(1, self._index)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            torch.cuda.synchronize(self.device)
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)              # no peer is still storing into this rank's ring
            _lib.lib().b200_peer_gather_destroy(self._h)
            self._h = ctypes.c_void_p()
