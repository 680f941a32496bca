#!/usr/bin/env python
"""Small workload for compute-sanitizer over the kernels changed at the end of round 2: large-batch forms forced
(k_sweep with its shared-memory column, k_open_compact, sorted k_round_select), caustic-crossing sources of a triple and
a binary lens, plain / gradient / limb-darkened calls.   compute-sanitizer --tool X python scripts/sanitize_target.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import caustics_b200 as cb  # noqa: E402
from caustics_b200 import _lib  # noqa: E402

L = _lib.lib()
L.caustics_set_tuning(b"ext_variants", 0)     # thread-per-source limb walk, sweep, compact open pass, unstaged contours
rng = np.random.default_rng(3)
n = int(os.environ.get("N", 400))
for nl, hp in ((3, dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0)), (2, dict(s=0.9, q=0.2))):
    _, ca = cb.critical_and_caustic_curves(npts=100, nlenses=nl, **hp)
    ca = ca.reshape(-1).cpu().numpy()
    ca = ca[rng.choice(len(ca), n, replace=True)]
    for rho in (1e-2, 2e-1):
        w = ca + rng.uniform(0, 2 * rho, n) * np.exp(1j * rng.uniform(-np.pi, np.pi, n))
        m = cb.mag_extended_source(w, rho, nlenses=nl, npts_limb=200, **hp)
        print("plain", nl, rho, bool(np.isfinite(np.asarray(m)).all()))
    wt = torch.from_numpy(w).cuda()
    s = torch.tensor(0.9, dtype=torch.float64, device="cuda", requires_grad=True)
    hp2 = dict(hp); hp2["s"] = s
    mg = cb.mag_extended_source(wt, 1e-2, nlenses=nl, npts_limb=200, **hp2)
    mg.sum().backward()
    print("grad", nl, bool(torch.isfinite(s.grad)))
    ml = cb.mag_extended_source(w, 1e-2, nlenses=nl, npts_limb=200, limb_darkening=True, u1=0.5, npts_ld=50, **hp)
    print("ld", nl, bool(np.isfinite(np.asarray(ml)).all()))
torch.cuda.synchronize()
