#!/usr/bin/env python
"""Randomised differential test (not collected by pytest; run on a GPU box): kernels 1 and 2 vs the
CPU oracle over random lens geometries / random polynomials.

    python tests/fuzz_point_source.py [n_configs] [points_per_config] [seed]
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
import caustics_b200 as cb  # noqa: E402
from conftest import set_distance  # noqa: E402
from oracle import lens, solver  # noqa: E402


def main():
    ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    npt = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    # ---- kernel 2: point-source magnification ----
    for k in range(ncfg):
        nl = 2 if k % 2 else 3
        s, q = float(rng.uniform(0.3, 2.5)), float(10 ** rng.uniform(-4, 0))
        hp = dict(s=s, q=q) if nl == 2 else dict(s=s, q=q, q3=float(10 ** rng.uniform(-2, 0)),
                                                   r3=float(rng.uniform(0.3, 1.5)), psi=float(rng.uniform(0, 6.28)))
        w = rng.uniform(-2, 2, npt) + 1j * rng.uniform(-2, 2, npt)
        got = cb.mag_point_source(w, nlenses=nl, **hp)
        want = lens.mag_point_source(w, nl, **hp)
        rel = np.abs(got / want - 1)
        bad = np.flatnonzero(rel > 1e-10)
        # classify the outliers: an image whose lens-equation residual sits on the 1e-6 filter, or |det J| ~ 0
        p, x_cm = lens.lens_params(nl, **hp)
        expl = 0
        for i in bad:
            z, m = lens.images_point_source(np.array([w[i] + x_cm]), nl, roots_compensated=True, **p)
            res = np.abs(lens.lens_eq(z, nl, **p) - (w[i] + x_cm))[:, 0]
            detj = np.abs(lens.lens_eq_det_jac(z, nl, **p))[:, 0]
            # the oracle's own sensitivity: plain vs compensated roots on the same point
            sens = abs(lens.mag_point_source(w[i:i + 1], nl, roots_compensated=True, **hp)[0] / want[i] - 1)
            if np.any((res > 1e-8) & (res < 1e-4)) or np.any(detj[m[:, 0]] < 1e-4) or rel[i] < 20 * sens:
                expl += 1
            else:
                print(f"      w={w[i]} gpu {got[i]!r} oracle {want[i]!r} rel {rel[i]:.2e} oracle plain-vs-compensated {sens:.2e} "
                      f"min|detJ| {detj[m[:, 0]].min():.2e} residuals {np.sort(res)[:6]}")
        print(f"PS cfg {k:3d} nl={nl} s={s:.3f} q={q:.2e}: max rel {rel.max():.2e}, {len(bad)}/{npt} beyond 1e-10 "
              f"({expl} at the image filter / on a caustic / within 20x the oracle's own plain-vs-compensated spread)" + ("" if expl == len(bad) else "   <-- UNEXPLAINED"), flush=True)
    # ---- warm-started walks (CAUSTICS_FLAG_GRID_WALK / _PATH_WALK) vs the cold kernels: random lenses, random
    # map patches with steps 1e-5 .. 3e-3, and random (non-path) arrays through the path walk
    for k in range(ncfg):
        nl = 2 if k % 2 else 3
        s, q = float(rng.uniform(0.3, 2.5)), float(10 ** rng.uniform(-4, 0))
        hp = dict(s=s, q=q) if nl == 2 else dict(s=s, q=q, q3=float(10 ** rng.uniform(-2, 0)),
                                                   r3=float(rng.uniform(0.3, 1.5)), psi=float(rng.uniform(0, 6.28)))
        step = float(10 ** rng.uniform(-5, -2.5))
        nx, ny = 1500, 160
        x0, y0 = float(rng.uniform(-1, 1)) - nx * step / 2, float(rng.uniform(-1, 1)) - ny * step / 2
        cold = cb.mag_point_source_map(x0, y0, step, step, nx, ny, nlenses=nl, walk=False, **hp)
        walk = cb.mag_point_source_map(x0, y0, step, step, nx, ny, nlenses=nl, walk=True, **hp)
        rel = ((walk - cold).abs() / cold).reshape(-1)
        # yardstick: how far two COLD solves of the same pixels are apart when only the starting values differ
        # (the fused kernel starts from Bini estimates; the images entry with flags=0 from the reference's)
        p_, xcm_ = cb.lens_params(nl, **hp)
        ix = torch.arange(nx, device="cuda", dtype=torch.float64)
        iy = torch.arange(ny, device="cuda", dtype=torch.float64)
        wm = ((x0 + ix * step)[None, :] + 1j * (y0 + iy * step)[:, None]).reshape(-1)

        def cold_ref_start(wv):
            z, m = cb.point_source._images_point_source(wv + xcm_, nl, flags=0, **p_)
            return ((1 / cb.lens_eq_det_jac(z, nl, **p_).abs()) * m).sum(0)
        own = ((cold_ref_start(wm) - cold.reshape(-1)).abs() / cold.reshape(-1))
        wr = torch.from_numpy(rng.uniform(-2, 2, 200_000) + 1j * rng.uniform(-2, 2, 200_000)).cuda()
        cr = cb.mag_point_source(wr, nlenses=nl, **hp)
        pr = (cb.mag_point_source(wr, nlenses=nl, flags=8, **hp) / cr - 1).abs()
        ownr = (cold_ref_start(wr) / cr - 1).abs()
        n = lambda t: int((t > 1e-9).sum().item())
        flag = "" if (n(rel) <= 2 * n(own) + 10 and n(pr) <= 2 * n(ownr) + 10) else "   <-- MORE THAN TWICE THE COLD-VS-COLD COUNT"
        print(f"WALK cfg {k:3d} nl={nl} s={s:.3f} q={q:.2e} step={step:.1e}: map walk vs cold: median {rel.median().item():.1e} max {rel.max().item():.1e} "
              f"{n(rel)}/{nx * ny} beyond 1e-9 [cold vs cold (other start): max {own.max().item():.1e}, {n(own)} beyond 1e-9]; "
              f"random array, path walk vs cold: max {pr.max().item():.1e} {n(pr)}/200000 beyond 1e-9 "
              f"[cold vs cold: max {ownr.max().item():.1e}, {n(ownr)}]{flag}", flush=True)
    # ---- kernel 1: random polynomials, all supported degrees, wild scales ----
    for deg in range(2, 17):
        n = 2000
        c = rng.standard_normal((n, deg + 1)) + 1j * rng.standard_normal((n, deg + 1))
        c *= 10.0 ** rng.uniform(-30, 30, (n, 1))                        # overall scale
        c *= 10.0 ** (rng.uniform(-1, 1, (n, 1)) * np.arange(deg + 1))   # root-radius scale
        want = solver.solve(np.ascontiguousarray(c), itmax=2500, compensated=True)
        for comp in (False, True):
            got = cb.primitive._solve_flat(torch.from_numpy(c).cuda(), None, 2500, comp, False, 0).cpu().numpy()
            d = set_distance(got, want)
            print(f"EA deg {deg:2d} comp={int(comp)}: max set distance {d.max():.2e}, {(d > 1e-12).sum()}/{n} beyond 1e-12", flush=True)


if __name__ == "__main__":
    main()
