"""Batch sharding of the MSDeformAttn path over ranks.

Every output row [b, q, m, :] and every grad_value[b] depends on sample b only (all addressing in
the reference is offset by b*S*M*D, ms_deform_im2col_cuda.cuh:266-269, 278), and the reference
already cuts the batch into independent chunks (ms_deform_attn_cuda.cu:61-75).  So the path shards
over the batch (frames of a clip) with NO exchange step: one process per GPU, each rank runs the
operator on its own slice, and the only collectives are the barrier and the max-over-ranks of the
timing.  Nothing here touches tensors on the data path.
"""
from __future__ import annotations

import os
from typing import Sequence, Tuple


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the batch owned by `rank`: contiguous, balanced to within one sample,
    earlier ranks take the remainder (the split torch.utils.data.DistributedSampler produces for
    consecutive indices, util/settings.py:599-607 in the reference)."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} world={world} rank={rank}")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(tensors: dict, world: int, rank: int, batched=("value", "loc", "aw", "grad_out")) -> dict:
    """Slice the batched operands of the operator (dim 0) for `rank`; level tables are shared."""
    n = next(tensors[k] for k in batched if k in tensors).shape[0]
    a, b = shard_range(n, world, rank)
    return {k: (v[a:b] if k in batched else v) for k, v in tensors.items()}


def env_world() -> Tuple[int, int, int]:
    """(world_size, rank, local_rank) as torchrun exports them; (1, 0, 0) when launched plainly."""
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str, device=None):
    """Rendezvous on 127.0.0.1 (the container hostname may not resolve).  Returns torch.distributed
    or None for a single process."""
    world, _, _ = env_world()
    if world <= 1:
        return None
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if not dist.is_initialized():
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, **kw)
    return dist


def max_over_ranks(values: Sequence[float], dist=None, device="cpu"):
    """Element-wise MAX of per-rank timings (a multi-GPU number is the slowest rank's)."""
    vals = [float(v) for v in values]
    if dist is None:
        return vals
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def aggregate_throughput(items_per_rank: int, world: int, seconds_max: float) -> float:
    """Whole-job units/s = the units ALL ranks processed / the slowest rank's time."""
    return world * items_per_rank / seconds_max
