"""CPU tier: the N>1 host logic (contiguous sharding, cost balancing, final gather) with
world_size-2 gloo process groups.  The per-shard function here is plain torch arithmetic -- the
kernels themselves are covered by the GPU tier; what is tested is that sharded == unsharded."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    sys.path.insert(0, ROOT)
    from caustics_b200.sharding import sharded_apply, shard_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.arange(n, dtype=torch.float64)
    calls = []

    def fn(rows):
        calls.append(len(rows))
        return torch.stack([rows * rows, rows + 0.5], dim=1)

    out = sharded_apply(fn, x)
    lo, hi = shard_bounds(n, world, rank)
    ok = bool(torch.equal(out, torch.stack([x * x, x + 0.5], dim=1))) and calls == [hi - lo]
    local = sharded_apply(fn, x, gather=False)
    ok = ok and local.shape[0] == hi - lo
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 1])
def test_sharded_apply_gloo_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True), (1, True)]


def test_shard_bounds_and_balance():
    from caustics_b200.sharding import shard_bounds, balanced_order
    for n in (0, 1, 5, 8, 1001):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    cost = np.array([1, 1000, 1, 1, 1000, 1, 1000, 1, 1, 1000])
    perm, inv = balanced_order(cost, 2)
    assert sorted(perm) == list(range(10)) and (perm[inv] == np.arange(10)).all()
    per = 5
    loads = [cost[perm[r * per:(r + 1) * per]].sum() for r in range(2)]
    assert max(loads) - min(loads) <= 1


def _rows_worker(rank, world, port, nrows, align, q):
    import sys
    sys.path.insert(0, ROOT)
    from caustics_b200.sharding import sharded_rows, row_block
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx = 5
    full = torch.arange(nrows, dtype=torch.float64)[:, None] * 10 + torch.arange(nx, dtype=torch.float64)[None, :]
    out = sharded_rows(lambda lo, hi: full[lo:hi].clone(), nrows, align=align)
    lo, hi = row_block(nrows, world, rank, align)
    ok = bool(torch.equal(out, full)) and (lo % align == 0 or lo == nrows)
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nrows,align", [(100, 32), (7, 1), (64, 32), (1, 32)])
def test_sharded_rows_gloo_world2(nrows, align):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, nrows, align, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True), (1, True)]


def test_row_block_alignment():
    from caustics_b200.sharding import row_block
    for nrows in (0, 1, 31, 32, 33, 1000, 10_000):
        for world in (1, 2, 4, 8):
            for align in (1, 32):
                spans = [row_block(nrows, world, r, align) for r in range(world)]
                assert spans[0][0] == 0 and spans[-1][1] == nrows
                assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
                assert all(lo % align == 0 or lo == nrows for lo, _ in spans)


def _host_gather_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from caustics_b200.sharding import HostGather, shard_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1001
    hg = HostGather(n * 8, dst=0)
    v = hg.view(np.float64, (n,))
    lo, hi = shard_bounds(n, world, rank)
    v[lo:hi] = np.arange(lo, hi) * 0.5          # what a *_host entry point's D2H copies do
    hg.finish()
    ok = bool(np.array_equal(v, np.arange(n) * 0.5))   # every rank sees the assembled result
    del v
    hg.close()
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_host_gather_gloo_world2():
    """HostGather: the ranks' slices land in ONE shared host buffer without a collective (the end-to-end
    path of a one-process-per-GPU driver); page-locking is skipped where there is no CUDA device"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_host_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True), (1, True)]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_peer_gather_needs_a_device():
    from caustics_b200._lib import CausticsError
    from caustics_b200.sharding import PeerGather
    with pytest.raises(CausticsError):
        PeerGather(1 << 20)
