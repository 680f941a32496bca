#!/usr/bin/env python
"""The kernels round 2 profiles, each launched a few times on its BASELINE workload (run under ncu by
scripts/gpu_profile_r2.sh):  python scripts/profile_targets.py <target>
  plain10  ea_kernel<10, false>  the headline: plain degree-10 solve of the C2 batch
  comp10   ea_kernel<10, true>   compensated degree-10 solve of the C2 batch (kernel (1) of north_star)
  deg5     ea_kernel<5, false>   10^6 degree-5 polynomials (C1 at GPU-filling size)
  map      ps_kernel<2, 0, 0>    2000 rows of the C5 map, per-pixel cold solves
  c4       kernel family 3 on 10^5 triple-lens sources (k_limb_walk, k_refine_select, k_refine_solve, k_tracks, k_contours)
  c4grad   the same with the fused tangent (k_contours<10, true>)
  c3       the gated limb-darkened light curve (k_gate, k_ld_pq, ...)
  c3x100   the same with 10^6 points
"""
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
if os.environ.get("EXT_WINDOWS"):
    L.caustics_set_tuning(b"ext_windows", int(os.environ["EXT_WINDOWS"]))
target = sys.argv[1]
reps = 3
if target == "plain10":
    c = torch.from_numpy(bench.make_coeffs(0, 1)).cuda()
    for _ in range(reps):
        cb.poly_roots(c, itmax=2500)
elif target == "comp10":
    c = torch.from_numpy(bench.make_coeffs(0, 1)).cuda()
    for _ in range(reps):
        cb.poly_roots(c, itmax=2500, compensated=True)
elif target == "deg5":
    from caustics_b200.point_source import _poly_coeffs_torch, lens_params
    p, x_cm = lens_params(2, **bench.HP2)
    c = _poly_coeffs_torch(torch.from_numpy(np.linspace(-2, 2, 1_000_000) + 0.1j + x_cm).cuda(), 2, **p)
    for _ in range(reps):
        cb.poly_roots(c, itmax=2500)
elif target == "map":
    for _ in range(reps):
        cb.mag_point_source_map(-1.5, -1.5, 3.0 / 9999, 3.0 / 9999, 10_000, 10_000, rows=(4000, 6000), walk=False, **bench.HP2)
elif target in ("c4", "c4grad"):
    n = 100_000
    if os.environ.get("OPEN_WSMALL"):
        L.caustics_set_tuning(b"open_wsmall", int(os.environ["OPEN_WSMALL"]))
    if os.environ.get("EXT_MASK"):
        L.caustics_set_tuning(b"ext_variants", int(os.environ["EXT_MASK"]))
    w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
    lens3 = cb.point_source._c_lens(3, 0.0, **bench.LENS)
    nb = L.caustics_ext_workspace_bytes(n, 3, 200, 0, 100)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    mag = torch.empty(n, dtype=torch.float64, device="cuda")
    grad = torch.empty((8, n), dtype=torch.float64, device="cuda")
    for _ in range(2):
        if target == "c4":
            _lib.check(L.caustics_mag_extended_source(w.data_ptr(), mag.data_ptr(), n, 1e-2, lens3, 200, 0, 0.0, 100, 2500, 0, ws.data_ptr(), nb, None))
        else:
            _lib.check(L.caustics_mag_extended_source_grad(w.data_ptr(), mag.data_ptr(), grad.data_ptr(), n, 1e-2, lens3, 200, 2500, 0, ws.data_ptr(), nb, None))
elif target == "c3x100":
    w = torch.from_numpy(np.linspace(-2, 2, 1_000_000) + 0.1j).cuda()
    for _ in range(2):
        cb.mag(w, 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100, **bench.HP2)
elif target == "c3":
    w = torch.from_numpy(np.linspace(-2, 2, 10_000) + 0.1j).cuda()
    for _ in range(reps):
        cb.mag(w, 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100, **bench.HP2)
else:
    raise SystemExit(__doc__)
torch.cuda.synchronize()
