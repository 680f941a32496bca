#!/usr/bin/env python
"""End-to-end headline step (C2 from pinned host buffers through caustics_b200.poly_roots) for several shapes of the
host pipeline: slots in flight x chunk length.   python scripts/e2e_probe.py"""
import json
import os
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import caustics_b200 as cb  # noqa: E402
from caustics_b200 import _lib  # noqa: E402

L = _lib.lib()
coeffs = bench.make_coeffs(0, 1)
pin_in = torch.from_numpy(coeffs).pin_memory()
pin_out = torch.empty((bench.N_POLY, bench.DEG), dtype=torch.complex128).pin_memory()
ref = None
for slots, lg in [tuple(int(v) for v in x.split(':')) for x in os.environ.get('SHAPES', '3:15,4:15,5:15,6:15,8:15,4:14,8:14,4:16,6:16,4:15').split(',')]:
    L.caustics_set_tuning(b"host_slots", slots)
    L.caustics_set_tuning(b"host_chunk_log2", lg)
    for _ in range(3):
        cb.poly_roots(pin_in.numpy(), itmax=2500, out=pin_out.numpy())
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(20):
            cb.poly_roots(pin_in.numpy(), itmax=2500, out=pin_out.numpy())
        best = min(best, (time.perf_counter() - t0) / 20 * 1e3)
    if ref is None:
        ref = pin_out.clone()
    print(json.dumps({"slots": slots, "chunk_log2": lg, "ms": round(best, 3), "same": bool(torch.equal(ref, pin_out))}), flush=True)
