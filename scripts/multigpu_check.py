#!/usr/bin/env python
"""Multi-GPU check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/multigpu_check.py

Asserts that sharded results are BITWISE equal to single-GPU results for
 - config 5: a binary point-source magnification map (row blocks per rank, final gather),
 - the primitive on a polynomial batch,
 - a gated light curve (`mag`) with cost-balanced dealing of the full-integration points,
and prints the aggregate map throughput (device-timed, max over ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import caustics_b200 as cb  # noqa: E402
from caustics_b200 import _lib  # noqa: E402
from caustics_b200.sharding import sharded_apply, shard_bounds, balanced_order  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hp = dict(s=0.9, q=0.2)
    ok = True
    # ---- point-source map, 2000 x 2000 slice of the C5 grid
    nx = ny = 2000
    x = torch.linspace(-1.5, 1.5, nx, dtype=torch.float64, device="cuda")
    w = (x[None, :] + 1j * x[:, None]).reshape(-1)
    full = sharded_apply(lambda rows: cb.mag_point_source(rows, nlenses=2, **hp), w)
    if rank == 0:
        single = cb.mag_point_source(w, nlenses=2, **hp)
        ok &= bool(torch.equal(full, single))
    # ---- primitive
    rng = np.random.default_rng(0)
    c = torch.from_numpy(rng.standard_normal((100001, 11)) + 1j * rng.standard_normal((100001, 11))).cuda()
    roots = sharded_apply(lambda rows: cb.poly_roots(rows, itmax=2500), c)
    if rank == 0:
        ok &= bool(torch.equal(roots, cb.poly_roots(c, itmax=2500)))
    # ---- gated light curve with cost balancing: gate everywhere (cheap), then deal the expensive points
    wl = torch.from_numpy(np.linspace(-2, 2, 4000) + 0.1j).cuda()
    m0, used = cb.mag(wl, 1e-2, nlenses=2, npts_limb=200, return_test=True, **hp) if rank == 0 else (None, None)
    _, hexa = cb.mag_gate(wl, 1e-2, **hp)                                          # gate decisions (cheap pass)
    perm, inv = balanced_order((~hexa).cpu().numpy() * 1000 + 1, world)
    perm_t, inv_t = torch.from_numpy(perm).cuda(), torch.from_numpy(inv).cuda()
    out = sharded_apply(lambda rows: cb.mag(rows, 1e-2, nlenses=2, npts_limb=200, **hp), wl[perm_t])[inv_t]
    if rank == 0:
        ok &= bool(torch.equal(out, m0))
    # ---- aggregate throughput of the map (device time, max over ranks)
    lo, hi = shard_bounds(w.numel(), world, rank)
    mine = w[lo:hi].contiguous()
    for _ in range(3):
        cb.mag_point_source(mine, nlenses=2, **hp)
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        cb.mag_point_source(mine, nlenses=2, **hp)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 10], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        print(f"multigpu_check world={world} bitwise_equal={ok} map_evals_per_s={w.numel() / (t.item() * 1e-3):.4e}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
