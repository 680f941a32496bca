#!/usr/bin/env python
"""Times kernel 1 for every library build in build_variants/ (experiments with compile-time knobs).
Each variant runs in its own process (CAUSTICS_B200_LIB selects the .so)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
import caustics_b200 as cb
from oracle import lens
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n = 1000000
P3 = dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)
c10 = torch.from_numpy(lens.poly_coeffs(np.linspace(-2, 2, n) + 0.1j, 3, **P3)).cuda()
p, x_cm = lens.lens_params(2, s=0.9, q=0.2)
c5 = torch.from_numpy(lens.poly_coeffs(np.linspace(-2, 2, n) + 0.1j + x_cm, 2, **p)).cuda()
rng = np.random.default_rng(0)
r10 = torch.from_numpy(rng.standard_normal((n, 11)) + 1j * rng.standard_normal((n, 11))).cuda()
out = {}
out["deg10_ms"] = t(lambda: cb.poly_roots(c10, itmax=2500))
out["deg10_bini_ms"] = t(lambda: cb.poly_roots(c10, itmax=2500, flags=1))
out["deg5_ms"] = t(lambda: cb.poly_roots(c5, itmax=2500))
out["rand10_ms"] = t(lambda: cb.poly_roots(r10, itmax=2500))
out["deg10_comp_ms"] = t(lambda: cb.poly_roots(c10, itmax=2500, compensated=True), reps=5)
dx = 3.0 / 9999
for walk in (False, True):
    out["c5_walk_ms" if walk else "c5_cold_ms"] = t(lambda: cb.mag_point_source_map(-1.5, -1.5, dx, dx, 10000, 10000, rows=(4000, 6000), walk=walk, s=0.9, q=0.2), reps=5)
z = cb.poly_roots(c10[:20000], itmax=2500, compensated=True).cpu().numpy()
np.save(sys.argv[1], z)
_, sw = cb.primitive._solve_flat(c10[:100000], None, 2500, False, False, 2, return_sweeps=True)
out["mean_sweeps"] = float(sw.abs().double().mean().item()); out["notconv"] = int((sw < 0).sum().item())
print(json.dumps(out))
''' % ROOT

libs = sorted(glob.glob(os.path.join(ROOT, "build_variants", "*.so")))
ref = None
import numpy as np
for lib in libs:
    env = dict(os.environ, CAUSTICS_B200_LIB=lib)
    npy = lib + ".npy"
    r = subprocess.run([sys.executable, "-c", CHILD, npy], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        print(os.path.basename(lib), "FAILED", r.stderr[-400:])
        continue
    z = np.sort_complex(np.load(npy))
    if ref is None:
        ref = z
    print(os.path.basename(lib), r.stdout.strip().splitlines()[-1], "max|dz| vs first", float(np.abs(z - ref).max()), flush=True)
