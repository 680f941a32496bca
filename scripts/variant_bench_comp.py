#!/usr/bin/env python
"""Compensated-mode timing and parity for every library build in build_variants/ (experiments)."""
import glob, json, os, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
import caustics_b200 as cb
from caustics_b200.point_source import _poly_coeffs_torch, lens_params
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n = 1000000
P3 = dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)
c10 = _poly_coeffs_torch(torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda(), 3, **P3)
p, x_cm = lens_params(2, s=0.9, q=0.2)
c5 = _poly_coeffs_torch(torch.from_numpy(np.linspace(-2, 2, n) + 0.1j + x_cm).cuda(), 2, **p)
out = {"deg10_comp_ms": t(lambda: cb.poly_roots(c10, itmax=2500, compensated=True)),
       "deg5_comp_ms": t(lambda: cb.poly_roots(c5, itmax=2500, compensated=True))}
g = np.load(%r)
worst = 0.0
for name in ("c1", "c2", "rand5", "rand10"):
    c = g[name + "_coeffs"]; c = c.reshape(-1, c.shape[-1])
    z = cb.poly_roots(torch.from_numpy(c).cuda(), itmax=2500, compensated=True).cpu().numpy()
    worst = max(worst, float(np.abs(z - g[name + "_roots_comp"]).max()))
out["max_ordered_diff_vs_reference_comp"] = worst
_, sw = cb.primitive._solve_flat(c10[:200000], None, 2500, True, False, 2, return_sweeps=True)
out["mean_sweeps"] = float(sw.abs().double().mean().item()); out["notconv"] = int((sw < 0).sum().item())
print(json.dumps(out))
''' % (ROOT, os.path.join(ROOT, "tests", "golden", "ea_golden.npz"))
for lib in sorted(glob.glob(os.path.join(ROOT, "build_variants", "*.so"))):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, CAUSTICS_B200_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-400:], flush=True)
