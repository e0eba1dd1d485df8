"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding and the record all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from junction_mpc.distributed import allgather_records, shard_bounds, shard_sizes


def test_shard_bounds_partition_the_batch():
    for B in [0, 1, 7, 4096, 4097, 1048576]:
        for world in [1, 2, 3, 4, 8]:
            cuts = [shard_bounds(B, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = shard_sizes(B, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == B
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(B * 8, dtype=torch.float64).reshape(B, 8)      # global record table
        lo, hi = shard_bounds(B, world, rank)
        got = allgather_records(full[lo:hi].clone(), B)
        results[rank] = bool(torch.equal(got, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4096, 4097, 5])
def test_record_allgather_gloo_world2(B):
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, B, results), nprocs=world, join=True)
        assert dict(results) == {0: True, 1: True}
