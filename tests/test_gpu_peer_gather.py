"""Result-table exchange over peer memory (b200_peer_gather_*, dist.PeerResultGatherer): one rank on one GPU, and two
ranks on two GPUs (skipped on a single-GPU box) against ResultGatherer's layout contract."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alufe_b200  # noqa: E402,F401
from alufe_b200 import dist as bdist  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_gatherer_single_rank_ring_reuse():
    """world = 1: push -> collect through the ring, more sequence numbers than slots, ragged frame counts."""
    n_streams, stride = 5, 52
    gat = bdist.PeerResultGatherer(n_streams, stride, max_frames=6, device="cuda:0", n_slots=3)
    rng = np.random.default_rng(0)
    kept = []
    for k in range(11):
        F = 1 + k % 6
        t = torch.from_numpy(rng.integers(-5, 1 << 20, (F, n_streams, stride)).astype(np.int32)).cuda()
        kept.append((gat.push_frames(t), t))
        if len(kept) > 2:
            seq, want = kept.pop(0)
            got = gat.collect(seq)
            gat.stream.synchronize()
            assert torch.equal(got, want), seq
    for seq, want in kept:
        got = gat.collect(seq)
        gat.stream.synchronize()
        assert torch.equal(got, want), seq
    with pytest.raises(ValueError):
        gat.push_frames(torch.zeros((7, n_streams, stride), dtype=torch.int32, device="cuda"))
    gat.close()


def _rank_main(rank, world, port, n_streams, stride, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    bdist.init_process_group_small_footprint(dev, rank=rank, world_size=world)
    gat = bdist.PeerResultGatherer(n_streams, stride, max_frames=4, device=dev, n_slots=2)
    ok = True
    pend = []
    for k in range(9):                                     # more pushes than slots: the acknowledgement path runs
        F = 1 + k % 4
        local = torch.empty((F, len(gat.local), stride), dtype=torch.int32, device=dev)
        for i, s in enumerate(gat.local):                  # value encodes (sequence, frame, stream, column)
            for f in range(F):
                local[f, i] = torch.arange(stride, device=dev, dtype=torch.int32) + 1000 * s + 100000 * f + 10000000 * k
        pend.append((gat.push_frames(local), F, k))
        if len(pend) > 1:
            seq, Fq, kq = pend.pop(0)
            got = gat.collect(seq)
            gat.stream.synchronize()
            for s in range(n_streams):
                for f in range(Fq):
                    want = torch.arange(stride, device=dev, dtype=torch.int32) + 1000 * s + 100000 * f + 10000000 * kq
                    ok = ok and bool(torch.equal(got[f, s], want))
    while pend:
        seq, Fq, kq = pend.pop(0)
        got = gat.collect(seq)
        gat.stream.synchronize()
        ok = ok and got.shape == (Fq, n_streams, stride)
    gat.close()
    dist.destroy_process_group()
    out[rank] = ok


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_peer_gatherer_two_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, 29533, 5, 52, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(0) is True and out.get(1) is True
