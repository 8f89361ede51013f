"""world_size-2 gloo test of the N>1 host logic: contiguous world shards, shard-exact synthetic
inputs, and the SUM all-reduce of the 8-double statistics vector."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crux_b200 import scenes, shard


def test_shard_ranges_partition_the_batch():
    for total in (4096, 4097, 7, 1):
        for ws in (1, 2, 3, 8):
            seen = []
            for r in range(ws):
                first, count = shard.shard_range(total, r, ws)
                seen.extend(range(first, first + count))
            assert seen == list(range(total))
    assert shard.weak_range(4096, 3) == (12288, 4096)
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _worker(rank, world_size, port, total_worlds, n_spheres, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    first, count = shard.shard_range(total_worlds, rank, world_size)
    pos, vel = scenes.sphere_worlds(count, n_spheres, 10.0, first_world=first)
    local = torch.from_numpy(shard.stats_from_state(vel))
    shard.all_reduce_stats(local)
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
    dist.all_gather(gathered, torch.tensor([count]))
    if rank == 0:
        out["stats"] = local.numpy().copy()
        out["counts"] = [int(g.item()) for g in gathered]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stats_allreduce_equals_single_process():
    total, n = 10, 32
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, total, n, out), nprocs=2, join=True)
    pos, vel = scenes.sphere_worlds(total, n, 10.0)
    want = shard.stats_from_state(vel)
    assert out["counts"] == [5, 5]
    assert np.allclose(out["stats"], want, rtol=1e-12, atol=1e-9)
    assert out["stats"][0] == total * n
