#!/usr/bin/env python
"""BASELINE config C5: the 10^4 x 10^4 binary-lens magnification map (10^8 point-source evaluations) sharded by
row blocks over the ranks (strong scaling: the map is fixed, each rank computes ceil(10^4 / N) rows into its own
buffer, no data-path collective).  Device-timed, max over ranks.  One rank per GPU:

    python scripts/bench_c5_map.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29544 scripts/bench_c5_map.py

Prints one JSON line per mode: per-pixel cold solves (bit-identical to mag_point_source on the explicit grid
whatever the sharding) and warm-started column walks (CAUSTICS_FLAG_GRID_WALK)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import caustics_b200 as cb  # noqa: E402
from caustics_b200.sharding import shard_bounds  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 10_000
    dx = 3.0 / (n - 1)
    r0, r1 = shard_bounds(n, world, rank)
    for walk in (False, True):
        fn = lambda: cb.mag_point_source_map(-1.5, -1.5, dx, dx, n, n, rows=(r0, r1), walk=walk, s=0.9, q=0.2)
        for _ in range(2):
            m = fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            m = fn()
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
        chk = torch.stack([m.sum(), torch.isfinite(m).all().double()])
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(chk, op=dist.ReduceOp.SUM)
        if rank == 0:
            print(json.dumps({"config": "C5 binary magnification map 10^4 x 10^4, rows sharded", "mode": "walk" if walk else "cold",
                              "n_gpus": world, "ms": t.item(), "evals_per_s": n * n / (t.item() * 1e-3),
                              "scaling": "strong", "map_sum": chk[0].item(), "all_finite": chk[1].item() == world}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
