"""CPU, world_size=2 over gloo: the N>1 host logic (batch sharding, max-over-ranks timing)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from automoe_b200._shard import job_throughput, shard_batch, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 11
        batch = {"image": torch.arange(n).float().view(n, 1, 1, 1).expand(n, 3, 2, 2), "speed": torch.arange(n).float().view(n, 1),
                 "meta": "keep"}
        mine = shard_batch(batch, world, rank)
        s, c = shard_range(n, world, rank)
        ids = mine["speed"].flatten().tolist()
        # every frame is owned by exactly one rank
        gathered = [None] * world
        dist.all_gather_object(gathered, ids)
        # weak-scaling throughput: rank r takes (r+1)*10 ms for 256 frames -> job = 512 frames / 20 ms
        fps = job_throughput(256, 10.0 * (rank + 1))
        q.put((rank, s, c, ids, gathered, fps, mine["meta"]))
    finally:
        dist.destroy_process_group()


def test_shard_and_timing_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_ids = sorted(i for r in res for i in r[3])
    assert all_ids == [float(i) for i in range(11)]
    assert [r[2] for r in res] == [6, 5]
    for r in res:
        assert r[4] == [res[0][3], res[1][3]]
        assert abs(r[5] - 512 / 0.020) < 1e-6
        assert r[6] == "keep"


def test_shard_range_properties():
    from automoe_b200._shard import shard_range
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from automoe_b200.training.train_gating_network import allreduce_flat_
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)        # rank 0: i, rank 1: 2i
        scale = allreduce_flat_(flat)
        q.put((rank, flat.tolist(), scale))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    """The training step's one exchange: SUM all-reduce of the flat gradient buffer, averaged by 1/world
    inside the fused optimizer kernel (here: the host logic over gloo)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, flat, scale in res:
        assert flat == [3.0 * i for i in range(10)] and scale == 0.5


def test_allreduce_flat_single_process_is_identity():
    from automoe_b200.training.train_gating_network import allreduce_flat_
    t = torch.ones(4)
    assert allreduce_flat_(t) == 1.0 and t.tolist() == [1.0] * 4


def _buffer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import torch.nn as nn
    from automoe_b200.training.train_gating_network import broadcast_buffers_
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Toy(nn.Module):
            def __init__(self):
                super().__init__()
                self.experts = nn.ModuleList([nn.BatchNorm2d(4)])
                self.policy = nn.BatchNorm2d(4)
        out = []
        for frozen_eval in (False, True):
            m = Toy()
            m.frozen_experts_eval = frozen_eval
            with torch.no_grad():
                for b in m.buffers():
                    b.fill_(rank + 1)                   # replicas drifted apart: rank 0 holds 1, rank 1 holds 2
            v0 = m.experts[0].running_mean._version
            broadcast_buffers_(m)
            out.append((m.policy.running_mean.tolist(), int(m.policy.num_batches_tracked), m.experts[0].running_mean.tolist(),
                        m.experts[0].running_mean._version - v0))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_broadcast_buffers_world2():
    """DDP-style per-step buffer sync: every rank takes rank 0's BatchNorm statistics and counters.  With the frozen experts on
    running statistics (frozen_experts_eval) their buffers are constants: they are left alone, so their Tensor._version - the
    key of the cached inference packs - does not move every step."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_buffer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, (plain, frozen) in res:
        assert plain[0] == [1.0] * 4 and plain[1] == 1 and plain[2] == [1.0] * 4 and plain[3] >= 1
        assert frozen[0] == [1.0] * 4 and frozen[1] == 1
        assert frozen[2] == [float(rank + 1)] * 4 and frozen[3] == 0       # untouched, version unchanged
