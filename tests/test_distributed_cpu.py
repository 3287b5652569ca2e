"""World-size-2 gloo test of the sharded-sampling host logic: contiguous shards,
rank-distinct seeds, and one all_gather that reproduces the single-rank concatenation."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lion_b200.utils.dist_sampling import gather_samples, rank_seed, shard_sizes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_sample(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 16, 3, generator=g)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = shard_sizes(total, world)
    local = _fake_sample(sizes[rank], rank_seed(5, rank))
    full = gather_samples(local)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def _run(total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    sizes = shard_sizes(total, world)
    expect = torch.cat([_fake_sample(sizes[r], rank_seed(5, r)) for r in range(world)], dim=0)
    for r in range(world):
        assert torch.equal(got[r], expect)


def test_even_shards():
    _run(8)


def test_ragged_shards():
    assert shard_sizes(7, 2) == [4, 3]
    _run(7)


def test_rank_seeds():
    assert rank_seed(3, 0) != rank_seed(3, 1)
    assert rank_seed(3, 1, reference_behaviour=True) == 3
