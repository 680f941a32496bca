#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--only C2,C5,...]

HEADLINE (the top-level keys of the JSON line; configs[1] of BASELINE.json, SURVEY 8(d) "C2"): 10^6
degree-10 lens polynomials of the triple lens a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197-0.95087i on
the trajectory w = linspace(-2, 2, 10^6) + 0.1i.  One step = one pass of the `ehrlich_aberth` primitive
over the batch; metric = roots/s (= polynomials/s x 10).  With N GPUs every rank solves its own
10^6-point slice of an N x 10^6-point trajectory (weak scaling, no collective on the data path).

  value  inputs resident in HBM, CUDA-event time of K launches (max over ranks)
  e2e    the same metric through the public host API (caustics_b200.poly_roots on pinned host
         arrays): H2D of the coefficients, kernel, D2H of the roots inside the timed region
  roofline  FP64-pipe roofline (the path is FP64 bound, not HBM or tensor bound): algorithmic flop
         per launch = root updates per polynomial (counted by the CPU port of the reference algorithm,
         DESIGN.md section 4) x F(10) = 28*10+21 flop / launch time, against the DFMA
         peak measured in the same run; the HBM figures are reported beside it
  cpu_baseline  the reference's own compiled solver (oracle/_ref) on the host, 1 core (it is
         single-threaded), on a bounded sample

`configs` carries the second half of BASELINE's metric and the other configurations, each with its own
value / e2e / roofline / cpu_baseline and, for N > 1, the multi-GPU behaviour BASELINE names:
  C5  the full 10^4 x 10^4 binary-lens magnification map, STRONG scaling: rows sharded over the ranks,
      per-pixel cold solves (the reference-identical kernel), timed without and WITH the final result
      gather -- the gather is the kernels' own stores into rank 0's buffer over NVLink
      (caustics_b200.sharding.PeerGather, no collective); `matches_single_gpu` is the bitwise comparison
      with the whole map computed on rank 0 alone
  C4  10^5 triple-lens uniform-disk extended-source magnifications, strong scaling over the sources
  C3  10^4-point binary-lens limb-darkened light curve through `mag` (hexadecapole gate); for N > 1 the
      points that fail the gate are dealt round-robin to the ranks after the gate
  C3x100  the same light curve with 10^6 points (workspace sized by the gate's survivors)
  C1  10^6 (and the reference-sized 10^4) degree-5 polynomials, plain and compensated; C2 compensated

--impl reference times the reference's CPU implementation with all host threads instead (headline and
the same `configs` keys, bounded samples).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POLY = 1_000_000
DEG = 10
LENS = dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)
HP2 = dict(s=0.9, q=0.2)                       # the binary lens of C1 / C3 / C5


def F_update(deg):
    """flop per plain root update (Horner value + derivative, Aberth sum, correction), SURVEY 8(d)"""
    return 28 * deg + 21


def F_confirm(deg):
    """flop of the evaluation that only confirms convergence: complex Horner of value and derivative"""
    return 16 * deg


F_UPDATE = F_update(DEG)
INIT_FLOP = 600                   # initial estimates + |coefficients| per polynomial
UPDATES_PER_POLY = 95.2           # plain root updates per polynomial on this workload
W10_CONTRACT = 2.63e4             # BASELINE.md section 4 fixed W10 (84.1 updates, another sample): reported beside
NCU_DRAM_BYTES_PER_LAUNCH = 2.98e8  # 176.1 MB read + 121.9 MB written, profiles/r01b_ncu_ea_kernel_deg10.csv
# Work models of the other configs: root updates / confirming evaluations counted on the device code
# compiled for the host with work counters (tests/hostsim, -DCB200_HOSTSIM_COUNT),
# plus a fixed part per solve (lens polynomial from the product form, image filter, Jacobian).
WORK = {
    # per polynomial: reference-compatible start, C1 trajectory
    "C1": {"updates": 25.1, "confirms": 5, "fixed": 300, "deg": 5},
    # per map pixel: Bini start, cold (18.4 updates + 5 confirming evaluations)
    "C5": {"updates": 18.4, "confirms": 5, "fixed": 300 + 500, "deg": 5},
    # per triple-lens source: 200 limb points x 10 roots
    "C4": {"updates": 2969.6, "confirms": 2000, "fixed": 200 * 1000, "deg": 10},
    # per binary-lens point of the gated light curve: gate solve everywhere ...
    "C3_gate": {"updates": 22.3, "confirms": 5, "fixed": 300 + 500 + 1500, "deg": 5},
    # ... and per fully integrated source (5.8 % of the points): solver part; the limb-darkening
    # quadrature is added from the measured vertex count (2 x npts_ld integrand evaluations per vertex)
    "C3_full": {"updates": 1986.8, "confirms": 1000, "fixed": 200 * 500, "deg": 5},
}
LD_INTEGRAND_FLOP = 70            # lens equation (2 reciprocals) + radius + square root + brightness
CONFIG = {"workload": "C2: ehrlich_aberth on 10^6 degree-10 triple-lens polynomials per GPU "
                      "(w=linspace(-2,2,N*10^6)+0.1i sliced per rank), plain mode, itmax=2500, "
                      "reference-compatible initial estimates",
          "polys_per_gpu": N_POLY, "deg": DEG,
          "l2": "inputs+outputs 336 MB per step > 126 MB L2 (no explicit flush needed)"}


def flop_model(key):
    m = WORK[key]
    return m["updates"] * F_update(m["deg"]) + m["confirms"] * F_confirm(m["deg"]) + m["fixed"]


def make_coeffs(rank, world, n=N_POLY, impl="ours"):
    """HIGH -> LOW coefficients (the poly_roots convention) of this rank's slice, complex128 NumPy.
    Our arm builds them with the product's own coefficient code on the GPU; the reference arm (no
    GPU needed) with the oracle's NumPy restatement -- the same polynomial to rounding."""
    w = np.linspace(-2, 2, world * n)[rank * n:(rank + 1) * n] + 0.1j
    out = np.empty((n, DEG + 1), dtype=np.complex128)
    step = 100000
    if impl == "ours":
        import torch
        from caustics_b200.point_source import _poly_coeffs_torch
        for i in range(0, n, step):
            out[i:i + step] = _poly_coeffs_torch(torch.from_numpy(w[i:i + step]).cuda(), 3, **LENS).cpu().numpy()
        return out
    from oracle import lens
    for i in range(0, n, step):
        out[i:i + step] = lens.poly_coeffs(w[i:i + step], 3, **LENS)
    return out


def c4_high_level():
    """the C2/C4 lens in the public (s, q, q3, r3, psi) parametrisation"""
    a, e1, e2, r3c = LENS["a"], LENS["e1"], LENS["e2"], LENS["r3"]
    q = e2 / e1
    return dict(s=2 * a, q=q, q3=q / e1 - 1 - q, r3=abs(r3c), psi=float(np.angle(r3c)))


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
        # wait for the first sample so the timed region is covered from its start
        t0 = time.time()
        while self.p is not None and time.time() - t0 < 3.0 and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any("Active" in r[4 + k] and "Not" not in r[4 + k] for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": reasons}


# =================================================================================================
# reference arm (CPU): the reference's compiled solver / the oracle restatement, all host threads
# =================================================================================================
def cpu_reference_time(coeffs_low_high, nthreads):
    from oracle import solver
    fn = solver.ref_solve if solver.ref_available() else solver.port_solve
    t0 = time.perf_counter()
    if nthreads == 1:
        fn(coeffs_low_high, itmax=2500)
    else:
        solver.threaded(fn, coeffs_low_high, nthreads, itmax=2500)
    return time.perf_counter() - t0


def _cpu_c5(w):
    from oracle import lens
    return lens.mag_point_source(w, 2, **HP2)


def _cpu_c4(w):
    from oracle import extended
    hp = c4_high_level()
    return np.array([extended.mag_extended_source(x, 1e-2, nlenses=3, npts_limb=200, **hp) for x in w])


def _cpu_c3(w):
    from oracle import extended
    return extended.mag(w, 1e-2, 2, 200, True, 0.7, 100, **HP2)


def _cpu_c1(c):
    from oracle import solver
    fn = solver.ref_solve if solver.ref_available() else solver.port_solve
    return fn(c, itmax=2500)


def cpu_config_samples():
    """bounded samples of every config for the CPU legs: name -> (function, input, units per element, what)"""
    from oracle import lens
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 10_000, 40_000)
    cols = rng.integers(0, 10_000, 40_000)
    w5 = (-1.5 + cols * (3.0 / 9999)) + 1j * (-1.5 + rows * (3.0 / 9999))
    hp = c4_high_level()
    _, x_cm3 = lens.lens_params(3, **hp)
    w4 = np.linspace(-2, 2, 100_000)[250::1667] + 0.1j - x_cm3
    w3 = np.linspace(-2, 2, 10_000)[::10] + 0.1j
    p2, x_cm2 = lens.lens_params(2, **HP2)
    c1 = np.ascontiguousarray(lens.poly_coeffs(np.linspace(-2, 2, 10_000) + 0.1j + x_cm2, 2, **p2)[:, ::-1])
    return {
        "C5": (_cpu_c5, w5, 1, "40 000 random pixels of the 10^4 x 10^4 map (oracle/lens.py + the reference's compiled solver)"),
        "C4": (_cpu_c4, w4, 1, "every 1667th source of the 10^5-point trajectory (oracle/extended.py + the reference's compiled solver)"),
        "C3": (_cpu_c3, w3, 1, "every 10th point of the 10^4-point light curve, gate on (oracle/extended.py)"),
        "C1": (_cpu_c1, c1, 5, "the full C1 case: 10^4 degree-5 polynomials (the reference's compiled solver)"),
    }


def cpu_configs(nproc):
    """time every config's CPU sample on `nproc` worker processes (the oracle layers are single-threaded
    NumPy/C; the sample is cut into nproc contiguous pieces).  Returns name -> cpu_baseline dict."""
    import multiprocessing as mp
    from oracle import solver
    kind = "reference" if solver.ref_available() else "port"
    units = {"C5": "evals/s", "C4": "evals/s", "C3": "evals/s", "C1": "roots/s"}
    out = {}
    pool = mp.get_context("fork").Pool(nproc) if nproc > 1 else None
    try:
        for name, (fn, x, per, what) in cpu_config_samples().items():
            if nproc > 1 and name in ("C5", "C1"):
                x = np.concatenate([x] * nproc)           # keep every worker busy for a measurable time
            parts = [x[i::nproc] for i in range(nproc) if len(x[i::nproc])]   # interleaved: even cost per piece
            if pool is not None:
                pool.map(fn, [p[:2] for p in parts])          # workers import the oracle outside the timing
            t0 = time.perf_counter()
            if pool is not None:
                pool.map(fn, parts)
            else:
                fn(x)
            dt = time.perf_counter() - t0
            out[name] = {"value": len(x) * per / dt, "unit": units[name], "cores": nproc, "kind": kind,
                         "sample": f"{what}; {len(x)} units in {dt:.2f} s"}
    finally:
        if pool is not None:
            pool.close()
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU solver, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import solver
    ncpu = os.cpu_count() or 1
    sample = 100_000
    c = np.ascontiguousarray(make_coeffs(0, 1, sample * 10, impl="reference")[::10][:, ::-1])
    for _ in range(args.warmup):
        cpu_reference_time(c[:20000], ncpu)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_time(c, ncpu)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample * DEG / dt
    kind = "reference" if solver.ref_available() else "port"
    cfgs = {k: {"value": b["value"], "unit": b["unit"], "cpu_baseline": b} for k, b in cpu_configs(ncpu).items()}
    print(json.dumps({
        "impl": "reference", "metric": "roots/s (deg 10, triple-lens trajectory)", "value": v,
        "unit": "roots/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": v, "unit": "roots/s", "cores": ncpu, "kind": kind,
                         "sample": f"{sample} of the 10^6 polynomials (every 10th) per step, "
                                   f"{ncpu} threads on disjoint slices of the reference's serial loop"},
        "e2e": {"value": v, "unit": "roots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "configs": cfgs}))


# =================================================================================================
# our arm
# =================================================================================================
class Ctx:
    """what every config bench needs: torch, the library, rank/world, barrier and max-over-ranks"""

    def __init__(self, args, torch, cb, _lib, dist, rank, world, local):
        self.args, self.torch, self.cb, self._lib, self.dist = args, torch, cb, _lib, dist
        self.rank, self.world, self.local = rank, world, local
        self.L = _lib.lib()
        self.stream = torch.cuda.current_stream().cuda_stream
        self.fp64_peak = None
        self.cpu = {}

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxreduce(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def time_device(self, fn, steps, warmup=3):
        """CUDA-event time per step of `fn` (enqueue only) on the current stream: barrier + synchronize on
        both sides, max over ranks; returns ms"""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        self.barrier()
        return self.maxreduce(a.elapsed_time(b) / steps)

    def time_wall(self, fn, steps, warmup=2):
        """wall-clock time per step of a synchronous host-API call (barrier on both sides, max over ranks)"""
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return self.maxreduce((time.perf_counter() - t0) / steps * 1e3)

    def roofline(self, flop_per_launch, ms, note, bytes_per_launch=None):
        ach = flop_per_launch / (ms * 1e-3) / 1e12
        r = {"bound": "fp64", "achieved": ach, "peak": self.fp64_peak, "unit": "TFLOP/s",
             "frac": ach / self.fp64_peak if self.fp64_peak else None, "traffic": None,
             "flop_per_launch": flop_per_launch, "work_model": note}
        if bytes_per_launch:
            r["hbm_gbs"] = bytes_per_launch / (ms * 1e-3) / 1e9
        return r


def bench_c5(cx, steps):
    """the full 10^4 x 10^4 binary map, rows sharded (strong scaling), cold per-pixel solves"""
    torch, L, _lib, cb = cx.torch, cx.L, cx._lib, cx.cb
    from caustics_b200 import sharding
    from caustics_b200.point_source import lens_params
    nx = ny = 10_000
    x0 = y0 = -1.5
    dx = dy = 3.0 / 9999
    p, x_cm = lens_params(2, **HP2)
    lens_c = cb.point_source._c_lens(2, x_cm, **p)
    lo, hi = sharding.row_block(ny, cx.world, cx.rank)
    local = torch.empty((hi - lo) * nx, dtype=torch.float64, device="cuda")
    peer = sharding.PeerGather(nx * ny * 8, dst=0)

    def launch(ptr, flags=0, r0=lo, r1=hi):
        _lib.check(L.caustics_mag_point_source_grid(x0, y0, dx, dy, nx, r0, r1, ptr, lens_c, 2500, 0, flags, cx.stream))

    ms_local = cx.time_device(lambda: launch(local.data_ptr()), steps)
    ms_gather = cx.time_device(lambda: launch(peer.ptr(lo * nx * 8)), steps)
    ms_walk = cx.time_device(lambda: launch(peer.ptr(lo * nx * 8), 4), steps)
    # bitwise check of the assembled map against the whole map computed on rank 0 alone (cold kernel)
    launch(peer.ptr(lo * nx * 8))
    peer.finish()
    same = None
    if cx.rank == 0:
        full = torch.empty(nx * ny, dtype=torch.float64, device="cuda")
        launch(full.data_ptr(), 0, 0, ny)
        torch.cuda.synchronize()
        got = peer.tensor(torch.float64, (nx * ny,))
        same = bool(torch.equal(got, full)) and bool(torch.isfinite(got).all().item())
        del full
    # end to end: every rank's row block lands in ONE host buffer (shared-memory segment, page-locked by
    # every rank), D2H inside the timed region, through the public map entry
    host = sharding.HostGather(nx * ny * 8, dst=0)
    hview = host.view(np.float64, (ny, nx))

    def e2e():
        cb.mag_point_source_map(x0, y0, dx, dy, nx, ny, rows=(lo, hi), walk=False, out=hview[lo:hi], **HP2)

    e2e_ms = cx.time_wall(e2e, max(2, min(5, steps)), warmup=1)
    same_host = None
    if cx.rank == 0:
        same_host = bool(np.array_equal(hview.reshape(-1), peer.tensor(torch.float64, (nx * ny,)).cpu().numpy()))
    del hview
    host.close()
    peer.close()
    n = nx * ny
    out = {"workload": "C5: binary-lens magnification map 10^4 x 10^4 (s=0.9, q=0.2, [-1.5,1.5]^2), per-pixel cold "
                       "solves; rows sharded over the ranks", "scaling": "strong", "unit": "evals/s", "n_gpus": cx.world,
           "value": n / (ms_gather * 1e-3), "ms_per_step": ms_gather,
           "value_no_gather": n / (ms_local * 1e-3), "ms_no_gather": ms_local, "gather_ms": ms_gather - ms_local,
           "gather": "kernel stores straight into rank 0's buffer over NVLink (CUDA IPC peer mapping), no collective",
           "matches_single_gpu": same,
           "walk_value": n / (ms_walk * 1e-3), "walk_ms": ms_walk,
           "walk_note": "opt-in warm-started column walks (CAUSTICS_FLAG_GRID_WALK), gather included; agrees with "
                        "the cold map to rounding x conditioning, not bit for bit",
           "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "evals/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": n * 8, "matches_device_path": same_host,
                   "api": "caustics_b200.mag_point_source_map(rows=..., out=host rows) -> "
                          "caustics_mag_point_source_grid_host; ranks write disjoint row blocks of one "
                          "page-locked shared-memory map"},
           "roofline": cx.roofline(n / cx.world * flop_model("C5"), ms_gather,
                                   "per rank: pixels x (18.4 updates x F(5)=161 + 5 confirming evaluations x 80 + 800 fixed)",
                                   n / cx.world * 8),
           "cpu_baseline": cx.cpu.get("C5"), "l2": "800 MB of output per step > 126 MB L2"}
    return out


def bench_c4(cx, steps):
    """10^5 triple-lens uniform-disk extended-source magnifications, sources sharded (strong scaling)"""
    torch, L, _lib, cb = cx.torch, cx.L, cx._lib, cx.cb
    from caustics_b200 import sharding
    n = 100_000
    hp = c4_high_level()
    w_all = np.linspace(-2, 2, n) + 0.1j            # centre-of-mass frame of the C2 lens (x_cm added by the library)
    lens3 = cb.point_source._c_lens(3, 0.0, **LENS)
    lo, hi = sharding.shard_bounds(n, cx.world, cx.rank)
    m = hi - lo
    w = torch.from_numpy(w_all[lo:hi]).cuda()
    nbytes = L.caustics_ext_workspace_bytes(m, 3, 200, 0, 100)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    local = torch.empty(m, dtype=torch.float64, device="cuda")
    peer = sharding.PeerGather(n * 8, dst=0)

    def launch(ptr, wp=None, mm=m, wsp=None, nb=nbytes):
        _lib.check(L.caustics_mag_extended_source((w if wp is None else wp).data_ptr(), ptr, mm, 1e-2, lens3, 200, 0, 0.0, 100, 2500, 0,
                                                  (ws if wsp is None else wsp).data_ptr(), nb, cx.stream))

    ms_local = cx.time_device(lambda: launch(local.data_ptr()), steps, warmup=2)
    ms_gather = cx.time_device(lambda: launch(peer.ptr(lo * 8)), steps, warmup=2)
    # forward value AND tangent w.r.t. (a, e1, e2, r3, w, rho) in one pass (C4 = "jax.grad through ...")
    grad = torch.empty((8, m), dtype=torch.float64, device="cuda")
    ms_grad = cx.time_device(lambda: _lib.check(L.caustics_mag_extended_source_grad(
        w.data_ptr(), local.data_ptr(), grad.data_ptr(), m, 1e-2, lens3, 200, 2500, 0, ws.data_ptr(), nbytes, cx.stream)),
        steps, warmup=2)
    grad_finite = bool(torch.isfinite(grad).all().item())
    del grad
    # weak scaling beside the strong-scaled figure: every rank its own 10^5 sources
    ms_weak = None
    if cx.world > 1:
        nbw = L.caustics_ext_workspace_bytes(n, 3, 200, 0, 100)
        wsw = torch.empty(nbw, dtype=torch.uint8, device="cuda")
        ww = torch.from_numpy(w_all).cuda()
        mw = torch.empty(n, dtype=torch.float64, device="cuda")
        ms_weak = cx.time_device(lambda: launch(mw.data_ptr(), ww, n, wsw, nbw), max(3, steps // 2), warmup=2)
        del wsw, ww, mw
    launch(peer.ptr(lo * 8))
    peer.finish()
    same, dev = None, None
    if cx.rank == 0:
        got = peer.tensor(torch.float64, (n,)).clone()
        if cx.world == 1:
            same, dev = True, 0.0
        else:
            del ws
            nb1 = L.caustics_ext_workspace_bytes(n, 3, 200, 0, 100)
            ws1 = torch.empty(nb1, dtype=torch.uint8, device="cuda")
            full = torch.empty(n, dtype=torch.float64, device="cuda")
            launch(full.data_ptr(), torch.from_numpy(w_all).cuda(), n, ws1, nb1)
            torch.cuda.synchronize()
            same = bool(torch.equal(got, full))
            dev = float(((got - full).abs() / full).max().item())
            del ws1, full
    # end to end through the public API on host arrays (high-level parameters; the library adds x_cm)
    from caustics_b200.point_source import lens_params
    _, x_cm = lens_params(3, **hp)
    w_host = w_all[lo:hi] - x_cm
    host = sharding.HostGather(n * 8, dst=0)
    hview = host.view(np.float64, (n,))

    def e2e():
        hview[lo:hi] = cb.mag_extended_source(w_host, 1e-2, nlenses=3, npts_limb=200, **hp)

    e2e_ms = cx.time_wall(e2e, max(2, min(5, steps)), warmup=1)
    e2e_dev = None
    if cx.rank == 0:
        e2e_dev = float(np.max(np.abs(hview / got.cpu().numpy() - 1)))
    del hview
    host.close()
    peer.close()
    return {"workload": "C4: triple-lens (C2 lens) uniform-disk mag_extended_source, 10^5 sources on w=linspace(-2,2)+0.1i, "
                        "rho=1e-2, npts_limb=200; sources sharded over the ranks", "scaling": "strong", "unit": "evals/s",
            "n_gpus": cx.world, "value": n / (ms_gather * 1e-3), "ms_per_step": ms_gather,
            "value_no_gather": n / (ms_local * 1e-3), "ms_no_gather": ms_local, "gather_ms": ms_gather - ms_local,
            "matches_single_gpu": same, "max_rel_dev_vs_single_gpu": dev,
            "gradient": {"value": n / (ms_grad * 1e-3), "unit": "evals/s", "ms_per_step": ms_grad, "finite": grad_finite,
                         "what": "caustics_mag_extended_source_grad: magnification + d mag/d(a, e1, e2, Re r3, Im r3, "
                                 "Re w, Im w, rho) per source in one pass (tangent accumulated while the contours are "
                                 "walked), device resident, no gather"},
            "weak_scaling": None if ms_weak is None else {"value": cx.world * n / (ms_weak * 1e-3), "ms_per_step": ms_weak,
                                                          "what": "every rank its own 10^5 sources, no gather"},
            "matches_note": "shards below 16 384 sources take the small-batch phase variants (other summation order "
                            "inside an Aberth sum): bitwise only when the shard stays on the same variants",
            "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "evals/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": n * 8, "max_rel_dev_vs_device_path": e2e_dev,
                    "api": "caustics_b200.mag_extended_source(numpy w, rho, nlenses=3, s=, q=, q3=, r3=, psi=)"},
            "roofline": cx.roofline(m * flop_model("C4"), ms_gather,
                                    "per rank: sources x (2969.6 updates x F(10)=301 + 2000 confirming evaluations x 160 "
                                    "+ 200 limb points x 1000 fixed), counts from the host-compiled device code"),
            "cpu_baseline": cx.cpu.get("C4"), "workspace_gb_per_rank": nbytes / 1e9,
            "l2": "7.2 GB workspace streamed per step > 126 MB L2"}


def bench_c3(cx, steps, n=10_000, key="C3"):
    """binary-lens limb-darkened light curve through `mag` with the hexadecapole gate; for N > 1 every rank
    runs the (cheap) gate on all points, the failing points are dealt round-robin and integrated into
    rank 0's buffer over NVLink"""
    torch, L, _lib, cb = cx.torch, cx.L, cx._lib, cx.cb
    from caustics_b200 import sharding
    from caustics_b200.point_source import lens_params
    p, x_cm = lens_params(2, **HP2)
    lens_c = cb.point_source._c_lens(2, x_cm, **p)
    w_np = np.linspace(-2, 2, n) + 0.1j
    w = torch.from_numpy(w_np).cuda()
    rho, u1, npts_ld = 1e-2, 0.7, 100
    res = {}
    torch.cuda.reset_peak_memory_stats()
    base_mem = torch.cuda.memory_allocated()

    if cx.world == 1:
        def step():
            res["m"], res["t"] = cb.mag(w, rho, nlenses=2, npts_limb=200, limb_darkening=True, u1=u1, npts_ld=npts_ld,
                                        return_test=True, **HP2)
        ms = cx.time_wall(step, steps, warmup=2)      # the public call reads the survivor count back: wall clock
        mag_dev = res["m"]
        nfull = int((~res["t"]).sum().item())
        same, dev = True, 0.0
    else:
        peer = sharding.PeerGather(n * 8, dst=0)
        mloc = torch.empty(n, dtype=torch.float64, device="cuda")
        used = torch.empty(n, dtype=torch.uint8, device="cuda")
        lst = torch.empty(n, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        state = {}

        def step():
            _lib.check(L.caustics_mag_gate(w.data_ptr(), mloc.data_ptr(), used.data_ptr(), lst.data_ptr(), cnt.data_ptr(),
                                           n, rho, lens_c, HP2["q"], 2500, 0, cx.stream))
            nf = int(cnt.item())
            # the gate's list is in atomic arrival order: sort it so that every rank deals the same cards
            mine = torch.sort(lst[:nf]).values[cx.rank::cx.world].contiguous()
            k = mine.numel()
            kc = torch.tensor([k], dtype=torch.int32, device="cuda")
            if cx.rank == 0:
                peer.tensor(torch.float64, (n,)).copy_(mloc)         # hexadecapole values everywhere ...
            cx.barrier()                                              # ... before anyone overwrites a failing point
            if k:
                nb = L.caustics_mag_workspace_bytes(k, k, 2, 200, 1, npts_ld)
                if state.get("nb", 0) < nb:
                    state["ws"], state["nb"] = torch.empty(nb, dtype=torch.uint8, device="cuda"), nb
                _lib.check(L.caustics_mag_extended_source_list(w.data_ptr(), peer.ptr(0), mine.data_ptr(), kc.data_ptr(), k,
                                                               rho, lens_c, 200, 1, u1, npts_ld, 2500, 0,
                                                               state["ws"].data_ptr(), state["nb"], cx.stream))
            peer.finish()
            state["nfull"] = nf

        ms = cx.time_wall(step, steps, warmup=2)
        nfull = state["nfull"]
        same, dev, mag_dev = None, None, None
        if cx.rank == 0:
            mag_dev = peer.tensor(torch.float64, (n,)).clone()
            ref = cb.mag(w, rho, nlenses=2, npts_limb=200, limb_darkening=True, u1=u1, npts_ld=npts_ld, **HP2)
            same = bool(torch.equal(mag_dev, ref))
            dev = float(((mag_dev - ref).abs() / ref).max().item())
    peak_gb = (torch.cuda.max_memory_allocated() - base_mem) / 1e9

    # end to end: host array in, host array out through the public API (every rank a contiguous slice of
    # the light curve at N > 1; the result lands in one shared host buffer)
    lo, hi = sharding.shard_bounds(n, cx.world, cx.rank)
    host = sharding.HostGather(n * 8, dst=0)
    hview = host.view(np.float64, (n,))

    def e2e():
        hview[lo:hi] = cb.mag(w_np[lo:hi], rho, nlenses=2, npts_limb=200, limb_darkening=True, u1=u1, npts_ld=npts_ld, **HP2)

    e2e_ms = cx.time_wall(e2e, max(2, min(5, steps)), warmup=1)
    e2e_dev = None
    if cx.rank == 0 and mag_dev is not None:
        e2e_dev = float(np.max(np.abs(hview / mag_dev.cpu().numpy() - 1)))
    del hview
    host.close()
    if cx.world > 1:
        peer.close()
    # work: the gate on every point + the solver part and the limb-darkening quadrature of the integrated ones
    verts = 5 * 200 * 0.6       # ~3 real image tracks of 200 limb points each per integrated binary source
    flop = (n * flop_model("C3_gate") + nfull * (flop_model("C3_full") + verts * 2 * npts_ld * LD_INTEGRAND_FLOP)) / cx.world
    return {"workload": f"{key}: binary-lens limb-darkened light curve through mag (hexadecapole gate), {n} points on "
                        "w=linspace(-2,2)+0.1i, rho=1e-2, u1=0.7, npts_limb=200, npts_ld=100", "scaling": "strong",
            "unit": "evals/s", "n_gpus": cx.world, "value": n / (ms * 1e-3), "ms_per_step": ms,
            "timing": "wall clock around the stream-ordered call incl. the 4-byte survivor-count read-back"
                      + ("" if cx.world == 1 else ", the barrier that orders the hexadecapole values before the "
                         "integrations and the final barrier; gather = kernel stores into rank 0's buffer"),
            "full_contour_integrations": nfull, "matches_single_gpu": same, "max_rel_dev_vs_single_gpu": dev,
            "peak_device_memory_gb": peak_gb,
            "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "evals/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * 16,
                    "d2h_bytes_per_step": n * 8, "max_rel_dev_vs_device_path": e2e_dev,
                    "api": "caustics_b200.mag(numpy w, rho, nlenses=2, limb_darkening=True, u1=, s=, q=)"},
            "roofline": cx.roofline(flop, ms, "per rank: points x gate (22.3 updates x 161 + 2700) + integrated sources x "
                                              "(1986.8 updates x 161 + 1000 x 80 + 1e5 + ~600 vertices x 200 integrand "
                                              "evaluations x 70)"),
            "cpu_baseline": cx.cpu.get("C3") if key == "C3" else None}


def bench_c1(cx, steps):
    """degree 5 (and compensated variants) on one GPU: roots/s with the deg-5 roofline"""
    torch, cb = cx.torch, cx.cb
    from caustics_b200.point_source import _poly_coeffs_torch, lens_params
    p, x_cm = lens_params(2, **HP2)
    out = {"workload": "C1: ehrlich_aberth on degree-5 binary-lens polynomials of w=linspace(-2,2,n)+0.1i "
                       "(n = 10^4 is the reference's CPU-runnable case, n = 10^6 fills the GPU)", "unit": "roots/s",
           "n_gpus": 1, "cpu_baseline": cx.cpu.get("C1")}
    for n, tag in ((1_000_000, "n1e6"), (10_000, "n1e4")):
        c5 = _poly_coeffs_torch(torch.from_numpy(np.linspace(-2, 2, n) + 0.1j + x_cm).cuda(), 2, **p)
        for comp in (False, True):
            ms = cx.time_device(lambda: cb.poly_roots(c5, itmax=2500, compensated=comp), steps)
            k = f"{tag}_{'compensated' if comp else 'plain'}"
            out[k] = {"value": 5 * n / (ms * 1e-3), "ms_per_step": ms}
            if not comp:
                out[k]["roofline"] = cx.roofline(n * flop_model("C1"), ms,
                                                 "polynomials x (25.1 updates x F(5)=161 + 5 x 80 + 300)", n * 176)
    out["value"] = out["n1e6_plain"]["value"]
    out["ms_per_step"] = out["n1e6_plain"]["ms_per_step"]
    out["roofline"] = out["n1e6_plain"]["roofline"]
    # kernel (1) of north_star on the headline batch: the compensated degree-10 solve
    c10 = torch.from_numpy(make_coeffs(0, 1)).cuda()
    ms = cx.time_device(lambda: cb.poly_roots(c10, itmax=2500, compensated=True), steps)
    out["C2_compensated"] = {"value": 10 * N_POLY / (ms * 1e-3), "ms_per_step": ms, "unit": "roots/s",
                             "roofline": cx.roofline(N_POLY * (UPDATES_PER_POLY * F_UPDATE + INIT_FLOP), ms,
                                                     "LOWER bound: the plain phase's work only (the polishing phase's "
                                                     "compensated Horner passes are not counted)")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only", default="", help="comma list of C1,C3,C3x100,C4,C5 (default: all); the headline always runs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import caustics_b200 as cb
    from caustics_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (caustics_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cx = Ctx(args, torch, cb, _lib, dist, rank, world, local)
    L, stream = cx.L, cx.stream
    barrier, maxreduce = cx.barrier, cx.maxreduce

    coeffs = make_coeffs(rank, world)                       # high -> low, host
    pin_in = torch.from_numpy(coeffs).pin_memory()
    pin_out = torch.empty((N_POLY, DEG), dtype=torch.complex128).pin_memory()
    d_in = pin_in.cuda()
    d_out = torch.empty((N_POLY, DEG), dtype=torch.complex128, device="cuda")
    flags = _lib.FLAG_COEFFS_HIGH_FIRST

    def step_device():
        _lib.check(L.caustics_ea_solve(d_in.data_ptr(), None, d_out.data_ptr(), None, N_POLY, DEG, 2500,
                                       0, 0, flags, stream))

    def step_e2e():
        cb.poly_roots(pin_in.numpy(), itmax=2500, out=pin_out.numpy())

    # ---- kernel-resident throughput ----------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms = maxreduce(ev0.elapsed_time(ev1) / args.steps)
    clocks = sampler.stop() if sampler else None
    value = world * N_POLY * DEG / (ms * 1e-3)

    # ---- end to end through the host API -----------------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(50, args.steps // 2))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = maxreduce((time.perf_counter() - t0) / e2e_steps * 1e3)
    e2e_value = world * N_POLY * DEG / (e2e_ms * 1e-3)
    # same result on both paths
    same = bool(torch.equal(d_out.cpu(), pin_out))

    # ---- FP64 peak measured in this run (DFMA microbenchmark), every rank (same device model) --------
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    blocks, iters = 148 * 8, 3 * (1 << 14)

    def dfma_peak(fn):
        for _ in range(2):
            fn(sink.data_ptr(), blocks, iters, stream)
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(sink.data_ptr(), blocks, iters, stream); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return 2.0 * 8 * 256 * blocks * iters / (best * 1e-3) / 1e12

    fp64_peak = dfma_peak(L.caustics_bench_fp64_peak)        # chains with two constant operands
    fp64_peak3 = dfma_peak(L.caustics_bench_fp64_peak3)      # three distinct register operands
    cx.fp64_peak = fp64_peak

    out = None
    cpu = None
    if rank == 0:
        # ---- algorithmic work per polynomial: the fixed constants of DESIGN.md section 4 ----------
        # (95.2 root updates per polynomial of this workload, counted by the CPU port of the
        # reference algorithm and pinned by tests/test_oracle.py::test_c2_update_count)
        upd_per_poly = UPDATES_PER_POLY
        flop_per_poly = upd_per_poly * F_UPDATE + INIT_FLOP
        achieved = N_POLY * flop_per_poly / (ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_bytes = N_POLY * ((DEG + 1) * 16 + DEG * 16)
        roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                    "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of "
                                      "ea_kernel<10,false> (profiles/); algorithmic bytes per launch = 3.36e8",
                    "peak_source": "DFMA microbenchmark in this run (MEASURED_PEAKS.json has no FP64 figure)",
                    "peak_3operand_dfma": fp64_peak3, "frac_of_3operand_peak": achieved / fp64_peak3,
                    "note": "a DFMA reading three distinct register pairs issues at ~69 % of the constant-operand "
                            "rate on B200 (register-file bandwidth); the solver's Horner/Aberth DFMAs are of that kind",
                    "flop_per_poly": flop_per_poly, "updates_per_poly": upd_per_poly,
                    "frac_with_contract_W10": N_POLY * W10_CONTRACT / (ms * 1e-3) / 1e12 / fp64_peak,
                    "contract_W10": W10_CONTRACT,
                    "hbm": {"achieved": hbm_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": hbm_bytes / (ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only (the other ranks would wait)
            from oracle import solver
            cs = np.ascontiguousarray(coeffs[:, ::-1])       # the whole step: ~10 s on one host core
            ncs = cs.shape[0]
            dt = cpu_reference_time(cs, 1)
            cpu = {"value": ncs * DEG / dt, "unit": "roots/s", "cores": 1,
                   "kind": "reference" if solver.ref_available() else "port",
                   "sample": f"one full step of the workload ({ncs} polynomials), one run, {dt:.1f} s; "
                             "the reference custom call is a serial loop (cpu_ops.cc:45-72)"}
            cx.cpu = cpu_configs(1)
        out = {"metric": "roots/s (deg 10, triple-lens trajectory)", "value": value, "unit": "roots/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": CONFIG, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "roots/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": N_POLY * (DEG + 1) * 16, "d2h_bytes_per_step": N_POLY * DEG * 16,
                       "api": "caustics_b200.poly_roots(numpy pinned) -> caustics_ea_solve_host",
                       "matches_device_path": same},
               "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu}
    del pin_in, pin_out, d_in, d_out, coeffs
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs (second half of the metric; multi-GPU where BASELINE shards them) ----
    want = [k for k in (args.only.split(",") if args.only else ["C5", "C4", "C3", "C3x100", "C1"]) if k]
    ksteps = max(3, min(10, args.steps))
    configs = {}
    for key in want:
        if key == "C5":
            c = bench_c5(cx, ksteps)
        elif key == "C4":
            c = bench_c4(cx, ksteps)
        elif key == "C3":
            c = bench_c3(cx, ksteps)
        elif key == "C3x100":
            c = bench_c3(cx, max(3, ksteps // 2), n=1_000_000, key="C3x100")
        elif key == "C1" and world == 1:
            c = bench_c1(cx, ksteps)
        else:
            continue
        configs[key] = c
        torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        out["configs"] = configs
        print(json.dumps(out))


if __name__ == "__main__":
    main()
