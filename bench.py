#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (configs[1] of BASELINE.json, SURVEY 8(d) "C2"): 10^6 degree-10 lens polynomials of the
triple lens a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197-0.95087i on the trajectory
w = linspace(-2, 2, 10^6) + 0.1i.  One step = one pass of the `ehrlich_aberth` primitive over the
batch.  metric = roots/s (= polynomials/s x 10).  With N GPUs every rank solves its own 10^6-point
slice of an N x 10^6-point trajectory (weak scaling, no collective on the data path).

  value  inputs resident in HBM, CUDA-event time of K launches (max over ranks)
  e2e    the same metric through the public host API (caustics_b200.poly_roots on pinned host
         arrays): H2D of the coefficients, kernel, D2H of the roots inside the timed region
  roofline  FP64-pipe roofline (the path is FP64 bound, not HBM or tensor bound): algorithmic flop
         per launch = (root updates per polynomial, counted by the CPU port of the reference
         algorithm on a sample) x F(10) = 28*10+21 flop / launch time, against the
         DFMA peak measured in the same run; the HBM figures are reported beside it
  cpu_baseline  the reference's own compiled solver (oracle/_ref) on the host, 1 core (it is
         single-threaded), on a bounded sample

--impl reference times the reference's CPU implementation with all host threads instead.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POLY = 1_000_000
DEG = 10
LENS = dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)
F_UPDATE = 28 * DEG + 21          # flop per plain root update, SURVEY 8(d)
INIT_FLOP = 600                   # initial estimates + |coefficients| per polynomial
UPDATES_PER_POLY = 95.2           # plain root updates per polynomial on this workload
NCU_DRAM_BYTES_PER_LAUNCH = 2.98e8  # 176.1 MB read + 121.9 MB written, profiles/r01b_ncu_ea_kernel_deg10.csv
CONFIG = {"workload": "C2: ehrlich_aberth on 10^6 degree-10 triple-lens polynomials per GPU "
                      "(w=linspace(-2,2,N*10^6)+0.1i sliced per rank), plain mode, itmax=2500, "
                      "reference-compatible initial estimates",
          "polys_per_gpu": N_POLY, "deg": DEG,
          "l2": "inputs+outputs 336 MB per step > 126 MB L2 (no explicit flush needed)"}


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


def other_configs(cb, L, _lib, torch):
    """The other BASELINE.json configs on this GPU (CUDA events, best of 3), reported beside the
    headline: C1 (degree-5 roots/s), C5 (binary point-source evals/s on a 2*10^7-point slice of the
    10^4 x 10^4 map), C4 (triple-lens uniform extended-source evals/s, 10^5 sources) and C3 (binary
    limb-darkened light curve through `mag`, 10^4 points with the hexadecapole gate)."""
    from caustics_b200.point_source import _poly_coeffs_torch, lens_params

    def best(fn, reps=3):
        fn(); torch.cuda.synchronize()
        t = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            t = min(t, a.elapsed_time(b) * 1e-3)
        return t

    out = {}
    hp2 = dict(s=0.9, q=0.2)
    p, x_cm = lens_params(2, **hp2)
    n = 1_000_000
    c5 = _poly_coeffs_torch(torch.from_numpy(np.linspace(-2, 2, n) + 0.1j + x_cm).cuda(), 2, **p)
    out["C1_ehrlich_aberth_deg5_roots_per_s"] = 5 * n / best(lambda: cb.poly_roots(c5, itmax=2500))
    out["C1_ehrlich_aberth_deg5_compensated_roots_per_s"] = 5 * n / best(lambda: cb.poly_roots(c5, itmax=2500, compensated=True))
    lens_c = cb.point_source._c_lens(2, x_cm, **p)
    nx, rows = 10_000, 2_000
    mag = torch.empty(nx * rows, dtype=torch.float64, device="cuda")
    out["C5_mag_point_source_binary_evals_per_s"] = nx * rows / best(lambda: _lib.check(
        L.caustics_mag_point_source_grid(-1.5, -1.5, 3.0 / 9999, 3.0 / 9999, nx, 4000, 4000 + rows, mag.data_ptr(),
                                         lens_c, 2500, 0, 0, None)))
    out["C5_mag_point_source_binary_walk_evals_per_s"] = nx * rows / best(lambda: _lib.check(
        L.caustics_mag_point_source_grid(-1.5, -1.5, 3.0 / 9999, 3.0 / 9999, nx, 4000, 4000 + rows, mag.data_ptr(),
                                         lens_c, 2500, 0, 4, None)))     # CAUSTICS_FLAG_GRID_WALK
    n4 = 100_000
    w4 = torch.from_numpy(np.linspace(-2, 2, n4) + 0.1j).cuda()
    lens3 = cb.point_source._c_lens(3, 0.0, **LENS)
    # C2 as the fused point-source magnification (coefficients + solve + filter + Jacobian in one kernel),
    # per-point cold solves and CAUSTICS_FLAG_PATH_WALK (the trajectory as warm-started runs)
    w2 = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
    m2 = torch.empty(n, dtype=torch.float64, device="cuda")
    for key, fl in (("C2_mag_point_source_triple_evals_per_s", 0), ("C2_mag_point_source_triple_path_walk_evals_per_s", 8)):
        out[key] = n / best(lambda: _lib.check(L.caustics_mag_point_source(w2.data_ptr(), m2.data_ptr(), None, n, lens3,
                                                                            2500, 0, fl, None)))
    m4 = torch.empty(n4, dtype=torch.float64, device="cuda")
    nbytes = L.caustics_ext_workspace_bytes(n4, 3, 200, 0, 100)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out["C4_mag_extended_source_triple_uniform_evals_per_s"] = n4 / best(lambda: _lib.check(
        L.caustics_mag_extended_source(w4.data_ptr(), m4.data_ptr(), n4, 1e-2, lens3, 200, 0, 0.0, 100, 2500, 0,
                                       ws.data_ptr(), nbytes, None)))
    del ws
    n3 = 10_000
    w3 = torch.from_numpy(np.linspace(-2, 2, n3) + 0.1j).cuda()
    res = {}

    def lc():
        res["m"], res["t"] = cb.mag(w3, 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100,
                                    return_test=True, **hp2)
    out["C3_mag_binary_ld_lightcurve_evals_per_s"] = n3 / best(lc)
    out["C3_full_contour_integrations"] = int((~res["t"]).sum().item())
    return out


def cpu_reference_time(coeffs_low_high, nthreads):
    from oracle import solver
    fn = solver.ref_solve if solver.ref_available() else solver.port_solve
    t0 = time.perf_counter()
    if nthreads == 1:
        fn(coeffs_low_high, itmax=2500)
    else:
        solver.threaded(fn, coeffs_low_high, nthreads, itmax=2500)
    return time.perf_counter() - t0


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
    print(json.dumps({
        "impl": "reference", "metric": "roots/s (deg 10, triple-lens trajectory)", "value": v,
        "unit": "roots/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": v, "unit": "roots/s", "cores": ncpu, "kind": kind,
                         "sample": f"{sample} of the 10^6 polynomials (every 10th) per step, "
                                   f"{ncpu} threads on disjoint slices of the reference's serial loop"},
        "e2e": {"value": v, "unit": "roots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    L = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream

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

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

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

    out = None
    if rank == 0:
        # ---- FP64 peak measured in this run (DFMA microbenchmark) ----------------------------
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
                    "hbm": {"achieved": hbm_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": hbm_bytes / (ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only (the other ranks would wait ~10 s)
            from oracle import solver
            cs = np.ascontiguousarray(coeffs[:, ::-1])       # the whole step: ~10 s on one host core
            ncs = cs.shape[0]
            dt = cpu_reference_time(cs, 1)
            cpu = {"value": ncs * DEG / dt, "unit": "roots/s", "cores": 1,
                   "kind": "reference" if solver.ref_available() else "port",
                   "sample": f"one full step of the workload ({ncs} polynomials), one run, {dt:.1f} s; "
                             "the reference custom call is a serial loop (cpu_ops.cc:45-72)"}
        extra = None
        if world == 1:
            extra = other_configs(cb, L, _lib, torch)
        out = {"metric": "roots/s (deg 10, triple-lens trajectory)", "value": value, "unit": "roots/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": CONFIG, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "roots/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": N_POLY * (DEG + 1) * 16, "d2h_bytes_per_step": N_POLY * DEG * 16,
                       "api": "caustics_b200.poly_roots(numpy pinned) -> caustics_ea_solve_host",
                       "matches_device_path": same},
               "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu, "extra": extra}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
