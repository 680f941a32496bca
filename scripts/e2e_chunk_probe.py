import os, sys, subprocess, json
CHILD = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import caustics_b200 as cb
from caustics_b200.point_source import _poly_coeffs_torch
LENS = dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)
n = 1000000
c = _poly_coeffs_torch(torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda(), 3, **LENS).cpu()
pin_in = c.pin_memory(); pin_out = torch.empty((n, 10), dtype=torch.complex128).pin_memory()
for _ in range(3): cb.poly_roots(pin_in.numpy(), itmax=2500, out=pin_out.numpy())
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): cb.poly_roots(pin_in.numpy(), itmax=2500, out=pin_out.numpy())
print((time.perf_counter() - t0) / 20 * 1e3)
'''
for chunk in (8192, 16384, 32768, 65536, 131072, 262144):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, CAUSTICS_B200_CHUNK=str(chunk)), capture_output=True, text=True)
    print(chunk, r.stdout.strip() or r.stderr[-300:], "ms", flush=True)
