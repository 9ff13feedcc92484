"""CPU, world_size 2 over gloo: stream partitioning and the result gather used on N > 1 GPUs."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

import alufe_b200  # noqa: E402,F401
from alufe_b200 import dist as bdist  # noqa: E402


def test_partition_is_a_disjoint_cover():
    for world in (1, 2, 3, 8):
        for n in (1, 7, 64):
            owned = [bdist.local_streams(n, r, world) for r in range(world)]
            assert sorted(s for o in owned for s in o) == list(range(n))
            assert all(bdist.stream_owner(s, world) == r for r, o in enumerate(owned) for s in o)
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def _worker(rank, world, port, n_streams, stride, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = bdist.ResultGatherer(n_streams, stride, "cpu")
        local = torch.stack([torch.full((stride,), 100 * s, dtype=torch.int32) + torch.arange(stride, dtype=torch.int32)
                             for s in g.local]) if g.local else torch.zeros((0, stride), dtype=torch.int32)
        out = g.gather(local)
        work, finish = g.gather(local, async_op=True)
        work.wait()
        out2 = finish()
        ok = all(int(out[s, 0]) == 100 * s and int(out[s, stride - 1]) == 100 * s + stride - 1 for s in range(n_streams))
        # three frames per collective (what bench.py does eight at a time): frame f adds 7 * f to every entry
        frames = torch.stack([local + 7 * f for f in range(3)])
        out3 = g.gather_frames(frames)
        work, finish = g.gather_frames(frames, async_op=True)
        work.wait()
        ok3 = tuple(out3.shape) == (3, n_streams, stride) and torch.equal(out3, finish()) and \
            all(torch.equal(out3[f], out + 7 * f) for f in range(3))
        q.put((rank, ok and ok3 and torch.equal(out, out2), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [4, 5])
def test_result_gather_world2_gloo(n_streams):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_streams
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _, _ in got) == [0, 1]
    assert all(ok and shape == (n_streams, 11) for _, ok, shape in got)


def test_reference_arm_is_silent_on_nonzero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "2", "--warmup", "3"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """bench.py --impl reference on rank 0: exactly one JSON line on stdout with the keys the driver reads."""
    import json
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "3"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 3 and d["setup_steps"] == 27 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert "workload" in d["config"] and d["metric"].startswith("tracked frames/sec")
