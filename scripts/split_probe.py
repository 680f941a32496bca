#!/usr/bin/env python
"""Two windows on two streams (caustics_set_tuning("ext_split", nA)) against one window: C4 and the binary lens at 10^5
sources, plain and with gradients; bitwise comparison of the results.   python scripts/split_probe.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import caustics_b200 as cb  # noqa: E402
from caustics_b200 import _lib  # noqa: E402

L = _lib.lib()


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for nl, n in [tuple(int(v) for v in x.split(':')) for x in os.environ.get('CASES', '3:100000,3:200000,2:100000,3:50000,2:50000,3:25000').split(',')]:
    if nl == 3:
        lens = cb.point_source._c_lens(3, 0.0, **bench.LENS)
    else:
        p, x_cm = cb.lens_params(2, s=0.9, q=0.2)
        lens = cb.point_source._c_lens(2, x_cm, **p)
    w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
    nb = L.caustics_ext_workspace_bytes(n, nl, 200, 0, 100)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    mag = torch.empty(n, dtype=torch.float64, device="cuda")
    grad = torch.empty((8, n), dtype=torch.float64, device="cuda")
    for what in ("plain", "grad"):
        if what == "plain":
            fn = lambda: _lib.check(L.caustics_mag_extended_source(w.data_ptr(), mag.data_ptr(), n, 1e-2, lens, 200, 0, 0.0, 100, 2500, 0, ws.data_ptr(), nb, None))
        else:
            fn = lambda: _lib.check(L.caustics_mag_extended_source_grad(w.data_ptr(), mag.data_ptr(), grad.data_ptr(), n, 1e-2, lens, 200, 2500, 0, ws.data_ptr(), nb, None))
        rec = {"nlenses": nl, "n": n, "what": what}
        ref = None
        for conf in os.environ.get("CONFIGS", "0:1,-1:2,-1:3,-1:4,65536:2").split(","):
            split, K = (int(v) for v in conf.split(":"))
            if split >= n:
                continue
            L.caustics_set_tuning(b"ext_split", split)
            L.caustics_set_tuning(b"ext_windows", K)
            mag.zero_(); grad.zero_()
            tag = f"s{split}k{K}"
            rec[tag] = round(timeit(fn), 3)
            cur = (mag.clone(), grad.clone())
            if ref is None:
                ref = cur
            else:
                rec["same_" + tag] = bool(torch.equal(ref[0], cur[0]) and (what == "plain" or torch.equal(ref[1], cur[1])))
        L.caustics_set_tuning(b"ext_windows", -1)
        L.caustics_set_tuning(b"ext_split", -1)
        print(json.dumps(rec), flush=True)
    del ws
