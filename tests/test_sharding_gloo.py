"""world_size-2 CPU test (gloo) of the multi-GPU host logic: global env id
layout and the end-of-run statistics all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mdp_playground_b200 import sharding


def test_group_id_bases_tile_the_global_id_space():
    sizes = [5, 3, 7]
    world = 4
    seen = []
    for rank in range(world):
        bases = sharding.group_id_bases(sizes, rank, world)
        for b, n in zip(bases, sizes):
            seen.extend(range(b, b + n))
    assert sorted(seen) == list(range(world * sum(sizes)))
    # a group's ids are contiguous over ranks
    b0 = sharding.group_id_bases(sizes, 0, world)
    b1 = sharding.group_id_bases(sizes, 1, world)
    assert [y - x for x, y in zip(b0, b1)] == sizes
    assert sharding.even_group_sizes(10, 4) == [3, 3, 2, 2]
    # whole CTAs per group when the batch allows it
    assert sharding.even_group_sizes(64 * 10, 4) == [192, 192, 128, 128]
    assert sum(sharding.even_group_sizes(1 << 20, 1000)) == 1 << 20
    assert all(n % 64 == 0 for n in sharding.even_group_sizes(1 << 20, 1000))
    assert sharding.even_group_sizes(64 * 3, 4) == [48] * 4
    assert [s.start for s in sharding.local_slices(sizes)] == [0, 5, 8]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = torch.zeros((3, 8), dtype=torch.float64)
    stats[:, 0] = torch.tensor([2.0, 4.0, 0.0]) * (rank + 1)     # episodes
    stats[:, 1] = torch.tensor([20.0, 40.0, 10.0]) * (rank + 1)  # transitions
    stats[:, 2] = torch.tensor([6.0, 2.0, 0.0]) * (rank + 1)     # reward
    red = sharding.reduce_stats(stats)
    summ = sharding.summarize_stats(red)
    if rank == 0:
        out.put({k: v.tolist() for k, v in summ.items()})
    dist.destroy_process_group()


def test_stats_allreduce_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got["episodes"] == [6.0, 12.0, 0.0]
    assert got["transitions"] == [60.0, 120.0, 30.0]
    assert got["episode_reward_mean"] == [3.0, 0.5, 0.0]
    assert got["episode_len_mean"] == [10.0, 10.0, 30.0]
