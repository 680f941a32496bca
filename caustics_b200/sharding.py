"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), contiguous slices of the
flattened source-position / polynomial axis, NO collective on the data path, one final gather.

Every polynomial (cpu_ops.cc:45-72) and every source position (lightcurve.py:245-254) is an
independent unit, so the only communication is the result gather."""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "sharded_apply", "balanced_order", "row_block", "sharded_rows"]


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


def row_block(nrows, world, rank, align=1):
    """Row block [lo, hi) of rank `rank` for a map of `nrows` rows: contiguous, equal to within one
    `align`-row unit.  With align = 32 (the walk length of csrc/ps_walk.cuh) every rank's block starts on a
    walk boundary of the unsharded map, so a map computed with `walk=True` does not depend on the number of
    ranks as long as every block is large enough to be given full 32-row walks (the launcher shortens the
    walks of maps too small to fill the GPU); each column segment is then the same walk wherever it runs.
    The default (align = 1) balances the rows exactly instead."""
    units = -(-nrows // align)
    per = -(-units // world) if world > 0 else units
    lo = min(rank * per * align, nrows)
    return lo, min(lo + per * align, nrows)


def sharded_rows(fn, nrows, group=None, gather=True, align=1):
    """Row-sharded evaluation of a map (BASELINE config C5): `fn(lo, hi)` returns this rank's rows as a
    (hi - lo, ...) tensor; no data-path collective, one all_gather at the end if `gather`.  Example:

        mag = sharded_rows(lambda lo, hi: mag_point_source_map(x0, y0, dx, dy, nx, ny, rows=(lo, hi), s=s, q=q),
                           ny, align=32)

    Works without an initialised process group (single process)."""
    if not (dist.is_available() and dist.is_initialized()):
        return fn(0, nrows)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = row_block(nrows, world, rank, align)
    local = fn(lo, hi)
    if not gather:
        return local
    per = row_block(nrows, world, 0, align)[1]
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:nrows]
