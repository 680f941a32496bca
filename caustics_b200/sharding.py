"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), contiguous slices of the
flattened source-position / polynomial axis, NO collective on the data path, one final gather.

Every polynomial (cpu_ops.cc:45-72) and every source position (lightcurve.py:245-254) is an
independent unit, so the only communication is the result gather."""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "sharded_apply", "balanced_order"]


def shard_bounds(n, world, rank):
    """contiguous ceil(n / world) split; the last ranks may get shorter (or empty) slices"""
    per = -(-n // world) if world > 0 else n
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def balanced_order(cost, world):
    """Permutation that deals the units round-robin in decreasing-cost order, so that gated workloads
    (a full contour integration costs ~10^3 hexadecapole points) spread evenly over contiguous
    shards.  Returns (perm, inverse): apply fn to x[perm] sharded, then out[inverse]."""
    cost = np.asarray(cost)
    order = np.argsort(-cost, kind="stable")
    n = len(order)
    per = -(-n // world)
    # position k of rank r's shard <- the (k*world + r)-th most expensive unit
    perm = np.full(per * world, -1, dtype=np.int64)
    for r in range(world):
        sel = order[r::world]
        perm[r * per:r * per + len(sel)] = sel
    perm = perm[perm >= 0]
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    return perm, inv


def sharded_apply(fn, x, group=None, gather=True):
    """Apply `fn` (rows in -> rows out, same leading length) to this rank's slice of `x` and, if
    `gather`, return the full result on every rank (one all_gather; padded to equal shard sizes).
    Works without an initialised process group (single process)."""
    if not (dist.is_available() and dist.is_initialized()):
        return fn(x)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    local = fn(x[lo:hi])
    if not gather:
        return local
    per = -(-n // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n]
