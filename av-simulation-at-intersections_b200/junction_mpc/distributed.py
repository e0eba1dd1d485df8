"""Multi-GPU plumbing: one process per GPU, the batch axis sharded in contiguous slices, no communication on the
solve path, one all-gather of the packed per-instance result record afterwards.

Only `torch.distributed` is used (NCCL over NVLink on the GPU box, gloo in the CPU tests); the send buffer is the
record buffer the step kernel's epilogue wrote, so there is no staging copy.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of B instances owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(B: int, world: int) -> List[int]:
    return [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)]


def allgather_records(local, B: Optional[int] = None, group=None):
    """All-gather per-instance records.  `local` is this rank's [b_r, W] tensor (CUDA for NCCL, CPU for gloo).
    With equal shard sizes this is a single `all_gather_into_tensor` straight out of `local`; ragged shards are
    padded to the largest shard first.  Returns the [B, W] tensor in global instance order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    b_r, W = local.shape
    if B is None:
        B = b_r * world
    sizes = shard_sizes(B, world)
    if sizes[dist.get_rank(group)] != b_r:
        raise ValueError("local shard size does not match shard_bounds()")
    m = max(sizes)
    if all(s == m for s in sizes):
        out = torch.empty(world * m, W, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    send = local if b_r == m else torch.cat([local, local.new_zeros(m - b_r, W)], 0)
    buf = torch.empty(world * m, W, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, send.contiguous(), group=group)
    return torch.cat([buf[r * m:r * m + sizes[r]] for r in range(world)], 0)


class FusedRecordGather:
    """All-gather of the result records fused into the step kernel (no collective call on the data path).

    Every rank allocates the gathered table [world * B, RECORD_LEN] in symmetric memory; after the rendezvous each
    GPU holds NVLink peer pointers to all tables, and `BatchedMPC.set_record_peers` makes the step kernel's epilogue
    store each instance's record straight into row `rank * B + b` of every table.  What is left after the launch is
    a cross-GPU barrier (symmetric-memory signal pads) so that readers see the peers' stores."""

    def __init__(self, engine, B_local: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _cabi
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("fused record gather addresses the GPUs of one NVSwitch box (<= 8)")
        dev = torch.device("cuda", engine.device)
        self.table = symm_mem.empty(self.world * B_local, _cabi.RECORD_LEN, dtype=torch.float64, device=dev)
        self.table.zero_()
        self.handle = symm_mem.rendezvous(self.table, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric-memory rendezvous returned no peer pointers")
        self.engine = engine
        engine.set_record_peers(ptrs, self.rank * B_local)

    def finish(self):
        """Call after `engine.step(...)` (same stream): returns the gathered table, valid once the stream has passed
        the barrier."""
        self.handle.barrier(channel=0)
        return self.table

    def close(self):
        self.engine.set_record_peers([], 0)


class ShardedMPC:
    """Runs this rank's slice of a global batch and gathers the result records.

    engine: a BatchedMPC on this rank's GPU.  Inputs to `step` are this rank's slices (device tensors)."""

    def __init__(self, engine, B_global: int, group=None):
        import torch.distributed as dist
        self.engine, self.B, self.group = engine, int(B_global), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(self.B, self.world, self.rank)

    def step(self, state, target_ind, oa, od, out, **kw):
        out = self.engine.step(state, target_ind, oa, od, out, **kw)
        if self.world > 1:
            return out, allgather_records(out.record, self.B, self.group)
        return out, out.record
