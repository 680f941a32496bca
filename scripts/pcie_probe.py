#!/usr/bin/env python
"""Why does the END-TO-END headline stop scaling at 8 GPUs?  (VERDICT r1, weak point 3.)

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/pcie_probe.py

Every rank owns one GPU and moves exactly the bytes of one headline step between PINNED host memory and
the device -- 176 MB host->device, 160 MB device->host -- with NO kernel in between, first alone (the other
ranks idle), then k = 2, 4, ... ranks at once.  Prints one JSON line: per-rank and aggregate GB/s for
H2D alone, D2H alone and both directions at once, as a function of the number of active ranks, plus the
box's PCIe / NUMA topology.  If the aggregate stops growing with k, the limit is the host side
(memory controller / PCIe root), not anything the library does.
"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

H2D_BYTES, D2H_BYTES = 176_000_000, 160_000_000


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hin = torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory()
    hout = torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory()
    din = torch.empty(H2D_BYTES, dtype=torch.uint8, device="cuda")
    dout = torch.zeros(D2H_BYTES, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(mode, active, reps=6):
        """seconds per repetition on this rank (0 if idle)"""
        barrier()
        t0 = time.perf_counter()
        if rank < active:
            for _ in range(reps):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        din.copy_(hin, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        hout.copy_(dout, non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps if rank < active else 0.0
        barrier()
        return dt

    res = {}
    ks = [k for k in (1, 2, 4, 8) if k <= world]
    for mode in ("h2d", "d2h", "both"):
        nbytes = {"h2d": H2D_BYTES, "d2h": D2H_BYTES, "both": H2D_BYTES + D2H_BYTES}[mode]
        for k in ks:
            run(mode, k, reps=2)
            dt = run(mode, k)
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            worst = t.item()
            res[f"{mode}_k{k}"] = {"per_rank_gbs": nbytes / worst / 1e9, "aggregate_gbs": k * nbytes / worst / 1e9,
                                  "ms": worst * 1e3}
    if rank == 0:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        print(json.dumps({"world": world, "bytes": {"h2d": H2D_BYTES, "d2h": D2H_BYTES}, "results": res,
                          "cpu_count": os.cpu_count(), "topo": topo.splitlines()[:14]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
