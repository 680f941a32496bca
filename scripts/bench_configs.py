#!/usr/bin/env python
"""Secondary measurements: every BASELINE.json config on one GPU, next to the CPU oracle.
(bench.py carries the headline config C2 and the JSON contract; this script prints one JSON line
per config and is what profiles/ cites for C1/C3/C4/C5.)

    python scripts/bench_configs.py [--cpu] [--only C3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import caustics_b200 as cb  # noqa: E402
from caustics_b200 import _lib  # noqa: E402

HP2 = dict(s=0.9, q=0.2)
C2_HP = None  # the C2 lens is given in low-level parameters
C2P = dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on a small sample")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    L = _lib.lib()
    out = []

    def emit(name, unit, n, t, **kw):
        rec = {"config": name, "value": n / t, "unit": unit, "n": n, "ms": t * 1e3}
        rec.update(kw)
        out.append(rec)
        print(json.dumps(rec), flush=True)

    want = lambda k: not args.only or k in args.only.split(",")
    from caustics_b200.point_source import _poly_coeffs_torch, lens_params

    if want("C1"):
        # C1: 10^4 degree-5 polynomials (binary trajectory); too small to fill a B200 -> also 10^6
        for n in (10_000, 1_000_000):
            p, x_cm = lens_params(2, **HP2)
            c = _poly_coeffs_torch(torch.from_numpy(np.linspace(-2, 2, n) + 0.1j + x_cm).cuda(), 2, **p)
            for comp in (False, True):
                t = timeit(lambda: cb.poly_roots(c, itmax=2500, compensated=comp))
                emit(f"C1 ehrlich_aberth deg5 n={n} compensated={comp}", "roots/s", 5 * n, t)
            w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
            t = timeit(lambda: cb.mag_point_source(w, nlenses=2, **HP2))
            emit(f"C1 mag_point_source binary n={n}", "evals/s", n, t)
    if want("C2"):
        n = 1_000_000
        w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
        lens_c = cb.point_source._c_lens(3, 0.0, **C2P)
        mag = torch.empty(n, dtype=torch.float64, device="cuda")
        for flags in (0, 1, 4):      # 4 = CAUSTICS_FLAG_GRID_WALK (warm-started column walks)
            t = timeit(lambda: _lib.check(L.caustics_mag_point_source(w.data_ptr(), mag.data_ptr(), None, n, lens_c, 2500, 0, flags, None)))
            emit(f"C2 fused mag_point_source triple n={n} flags={flags}", "evals/s", n, t)
        c = _poly_coeffs_torch(torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda(), 3, **C2P)
        for comp, flags in ((False, 0), (False, 1), (True, 0)):
            t = timeit(lambda: cb.poly_roots(c, itmax=2500, compensated=comp, flags=flags))
            emit(f"C2 ehrlich_aberth deg10 n={n} compensated={comp} flags={flags}", "roots/s", 10 * n, t)
    if want("C3"):
        n = 10_000
        w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
        res = {}
        def run():
            res["m"], res["t"] = cb.mag(w, 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100, return_test=True, **HP2)
        t = timeit(run, reps=3, warm=1)
        nfull = int((~res["t"]).sum().item())
        emit("C3 mag binary LD light curve n=10^4 (gate on)", "evals/s", n, t, full_integrations=nfull)
        wf = w[~res["t"]]
        t = timeit(lambda: cb.mag_extended_source(wf, 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100, **HP2), reps=3, warm=1)
        emit("C3 full LD contour integrations only", "evals/s", nfull, t)
        t = timeit(lambda: cb.mag_extended_source(wf, 1e-2, nlenses=2, npts_limb=200, **HP2), reps=3, warm=1)
        emit("C3 same points, uniform disk", "evals/s", nfull, t)
        # f3: one likelihood evaluation = trajectory -> mag -> marginalised likelihood, wall clock incl.
        # the single 24-byte read-back (what an HMC step calls)
        tt = torch.linspace(-40.0, 40.0, n, dtype=torch.float64, device="cuda")
        # psi = atan2(piEE, piEN) = pi/2: u = tau - 0.1i, the mirror image of the C3 trajectory (no parallax tables)
        tp = dict(t0=0.0, tE=20.0, u0=0.1, piEE=1e-12, piEN=0.0)
        traj = cb.AnnualParallaxTrajectory()
        A0 = cb.mag(traj.compute(tt, **tp), 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100, **HP2)
        fobs = 2.0 * A0 + 0.5
        cinv = torch.full_like(fobs, 1e4)
        def like():
            return cb.light_curve_log_likelihood(tt, fobs, cinv, traj, 1e-2, tp, HP2, npts_limb=200,
                                                 limb_darkening=True, u1=0.7, npts_ld=100)
        like(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            like()
        dt = (time.perf_counter() - t0) / 5
        emit("C3 likelihood closure: trajectory -> mag (LD, gated) -> marginalised log-likelihood, n=10^4, wall clock", "evals/s", n, dt)
        if args.cpu:
            from oracle import extended      # CPU baseline leg only
            sub = np.linspace(-2, 2, n)[::100] + 0.1j
            t0 = time.perf_counter(); extended.mag(sub, 1e-2, 2, 200, True, 0.7, 100, **HP2); dt = time.perf_counter() - t0
            emit("C3 CPU oracle (NumPy restatement + reference solver), 1 core, every 100th point", "evals/s", len(sub), dt)
    if want("C4"):
        n = 100_000
        # C4 lens is given in low-level parameters -> call the C ABI directly
        w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
        lens_c = cb.point_source._c_lens(3, 0.0, **C2P)
        mag = torch.empty(n, dtype=torch.float64, device="cuda")
        nbytes = L.caustics_ext_workspace_bytes(n, 3, 200, 0, 100)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        t = timeit(lambda: _lib.check(L.caustics_mag_extended_source(w.data_ptr(), mag.data_ptr(), n, 1e-2, lens_c, 200, 0, 0.0, 100, 2500, 0, ws.data_ptr(), nbytes, None)), reps=3, warm=1)
        emit("C4 mag_extended_source triple uniform n=10^5", "evals/s", n, t, workspace_gb=nbytes / 1e9,
             finite=bool(torch.isfinite(mag).all().item()))
    if want("C5"):
        nx, rows = 10_000, 2_000     # 2*10^7 of the 10^8 grid points per timing
        p, x_cm = lens_params(2, **HP2)
        lens_c = cb.point_source._c_lens(2, x_cm, **p)
        mag = torch.empty(nx * rows, dtype=torch.float64, device="cuda")
        for flags in (0, 1, 4):      # 4 = CAUSTICS_FLAG_GRID_WALK (warm-started column walks)
            t = timeit(lambda: _lib.check(L.caustics_mag_point_source_grid(-1.5, -1.5, 3.0 / 9999, 3.0 / 9999, nx, 4000, 4000 + rows, mag.data_ptr(), lens_c, 2500, 0, flags, None)), reps=3, warm=1)
            emit(f"C5 magnification map rows 4000-6000 of 10^4x10^4 flags={flags}", "evals/s", nx * rows, t)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
