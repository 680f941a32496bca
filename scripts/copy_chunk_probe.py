#!/usr/bin/env python
"""One headline step's bytes (176 MB in, 160 MB out) moved between pinned host memory and the device as ONE copy per
direction or cut into pieces of 2^k polynomials, both directions at once on two streams, no kernel: what does the
cutting itself cost?   python scripts/copy_chunk_probe.py"""
import json
import time

import torch

NP_, BC, BR = 1_000_000, 176, 160
hin = torch.empty(NP_ * BC, dtype=torch.uint8).pin_memory()
hout = torch.empty(NP_ * BR, dtype=torch.uint8).pin_memory()
din = torch.empty(NP_ * BC, dtype=torch.uint8, device="cuda")
dout = torch.zeros(NP_ * BR, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(chunk, mode, reps=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for off in range(0, NP_, chunk):
            m = min(chunk, NP_ - off)
            if mode != "d2h":
                with torch.cuda.stream(s1):
                    din[off * BC:(off + m) * BC].copy_(hin[off * BC:(off + m) * BC], non_blocking=True)
            if mode != "h2d":
                with torch.cuda.stream(s2):
                    hout[off * BR:(off + m) * BR].copy_(dout[off * BR:(off + m) * BR], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for chunk in (NP_, 1 << 18, 1 << 17, 1 << 16, 1 << 15, 1 << 14):
    run(chunk, "both", 2)
    print(json.dumps({"polys_per_piece": chunk, "h2d_ms": round(run(chunk, "h2d"), 3), "d2h_ms": round(run(chunk, "d2h"), 3),
                      "both_ms": round(run(chunk, "both"), 3)}), flush=True)
