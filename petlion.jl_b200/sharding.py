"""Batch sharding across the GPUs of one box.

Simulations are independent, so the path shards with NO data-path collective: rank r owns the
contiguous block [r*B/G, (r+1)*B/G) of systems.  The only collective is one all-gather of the
fixed-size (80-byte) per-system summary records after the integrate kernel -- `nccl` on GPUs,
`gloo` in the CPU tests.
"""
import numpy as np


def shard_bounds(B, world, rank):
    """contiguous block of systems owned by `rank` (the first B % world ranks get one extra)"""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_summaries(local, B, group=None):
    """all-gather variable-size blocks of the [n_local, 10] float64 summary records into [B, 10]"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)]
    nmax = max(sizes)
    buf = torch.zeros(nmax, local.shape[1], dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


def summaries_as_f64(summ):
    """structured numpy summary array -> [B, 10] float64 view (80-byte records)"""
    return np.ascontiguousarray(summ).view(np.float64).reshape(len(summ), 10)
