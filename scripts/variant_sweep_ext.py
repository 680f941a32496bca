#!/usr/bin/env python
"""Which phase variants of kernel family 3 win at which batch size?  (the shards of a strong-scaled C4 run:
10^5 / N sources per GPU)

    python scripts/variant_sweep_ext.py            # one JSON line per (lens, batch size): ms per variant mask

Variant mask bits (csrc/extended.cu): 1 warp-per-source selection, 2 staged stitching, 4 warp-per-source
limb-darkened sum, 8 lane-per-root limb walk, 16 lane-per-root refinement solves; -1 = the library's rule.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import caustics_b200 as cb  # noqa: E402
from caustics_b200 import _lib  # noqa: E402

L = _lib.lib()
if os.environ.get("EXT_WINDOWS"):
    L.caustics_set_tuning(b"ext_windows", int(os.environ["EXT_WINDOWS"]))
C2P = dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    if os.environ.get("OPEN_WSMALL"):
        L.caustics_set_tuning(b"open_wsmall", int(os.environ["OPEN_WSMALL"]))
    masks = [int(x) for x in os.environ.get("MASKS", "-1,0,2,8,10").split(",")]
    for nl in (3, 2):
        if nl == 3:
            lens = cb.point_source._c_lens(3, 0.0, **C2P)
        else:
            p, x_cm = cb.lens_params(2, s=0.9, q=0.2)
            lens = cb.point_source._c_lens(2, x_cm, **p)
        for n in [int(x) for x in os.environ.get("SIZES", "100000,50000,25000,12500,6250,3000").split(",")]:
            w = torch.from_numpy(np.linspace(-2, 2, 100_000)[:n] * 1.0 + 0.1j).cuda()
            w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
            mag = torch.empty(n, dtype=torch.float64, device="cuda")
            nb = L.caustics_ext_workspace_bytes(n, nl, 200, 0, 100)
            ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
            rec = {"nlenses": nl, "n": n}
            ref = None
            for m in masks:
                L.caustics_set_tuning(b"ext_variants", m)
                fn = lambda: _lib.check(L.caustics_mag_extended_source(w.data_ptr(), mag.data_ptr(), n, 1e-2, lens, 200, 0, 0.0,
                                                                       100, 2500, 0, ws.data_ptr(), nb, None))
                rec[f"mask{m}"] = round(timeit(fn), 3)
                if ref is None:
                    ref = mag.clone()
                    if os.environ.get("DUMP"):  # bitwise comparisons between library builds
                        np.save(os.path.join(ROOT, "gpurun_out", f"{os.environ['DUMP']}_{nl}_{n}.npy"), ref.cpu().numpy())
                else:
                    rec[f"dev{m}"] = float(((mag - ref).abs() / ref).max().item())
            L.caustics_set_tuning(b"ext_variants", -1)
            print(json.dumps(rec), flush=True)
            del ws


if __name__ == "__main__":
    main()
