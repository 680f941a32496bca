#!/usr/bin/env python
"""Randomised differential test (not collected by pytest; run on a GPU box):
GPU kernels vs the CPU oracle over random lens geometries, source radii and positions near caustics.

    python tests/fuzz_extended.py [n_configs] [points_per_config] [seed] [mode]
"""
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import caustics_b200 as cb  # noqa: E402
from oracle import extended  # noqa: E402


def main():
    ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    npt = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0     # 1: also vary sampling, radii up to 1, field points, gate
    worst = 0.0
    nbad = ntot = 0
    t0 = time.time()
    for k in range(ncfg):
        nl = 2 if k % 3 else 3
        s, q = float(rng.uniform(0.4, 2.0)), float(10 ** rng.uniform(-3, 0))
        hp = dict(s=s, q=q) if nl == 2 else dict(s=s, q=q, q3=float(10 ** rng.uniform(-2, 0)),
                                                   r3=float(rng.uniform(0.3, 1.5)), psi=float(rng.uniform(0, 6.28)))
        rho = float(10 ** rng.uniform(-3, -0.5 if mode == 0 else 0.0))
        _, ca = cb.critical_and_caustic_curves(npts=80, nlenses=nl, **hp)
        ca = ca.reshape(-1).cpu().numpy()
        ca = ca[rng.choice(len(ca), npt, replace=False)]
        w = ca + rng.uniform(0, 2 * rho, npt) * np.exp(1j * rng.uniform(-np.pi, np.pi, npt))
        if mode and k % 4 == 3:      # anywhere in the field, not only at the caustics
            w = rng.uniform(-2, 2, npt) + 1j * rng.uniform(-2, 2, npt)
        ld = bool(k % 2)
        u1 = float(rng.uniform(0, 1)) if ld else 0.0
        N = 200 if mode == 0 else int(rng.choice([100, 150, 200, 300, 400]))
        nld = 60 if mode == 0 else int(rng.choice([20, 51, 100]))
        if mode and nl == 2 and k % 5 == 0:   # the gated light-curve entry, decisions included
            got, gt = cb.mag(w, rho, nlenses=2, npts_limb=N, limb_darkening=ld, u1=u1, npts_ld=nld, return_test=True, **hp)
            want, wt = extended.mag(w, rho, 2, N, ld, u1, nld, return_test=True, **hp)
            if not np.array_equal(np.asarray(gt), np.asarray(wt)):
                print(f"cfg {k}: gate decisions differ at", np.flatnonzero(np.asarray(gt) != np.asarray(wt)))
        else:
            got = cb.mag_extended_source(w, rho, nlenses=nl, npts_limb=N, limb_darkening=ld, u1=u1, npts_ld=nld, **hp)
            want = np.array([extended.mag_extended_source(x, rho, nl, N, ld, u1, nld, **hp) for x in w])
        rel = np.abs(got / want - 1)
        worst = max(worst, rel.max())
        nbad += int((rel > 1e-4).sum())
        ntot += npt
        flag = "  <-- > 1e-4" if rel.max() > 1e-4 else ""
        print(f"cfg {k:3d} nl={nl} s={s:.3f} q={q:.2e} rho={rho:.2e} N={N} ld={int(ld)} u1={u1:.2f}: max rel {rel.max():.2e} "
              f"(>1e-8: {(rel > 1e-8).sum()}/{npt}) mags {want.min():.2f}..{want.max():.2f}{flag}", flush=True)
        if rel.max() > 1e-4:
            i = int(np.argmax(rel))
            print("      worst point", w[i], "gpu", got[i], "oracle", want[i], "params", hp)
    print(f"fuzz: {ntot} points, worst {worst:.2e}, {nbad} beyond 1e-4, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
