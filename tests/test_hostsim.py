"""CPU tier: the product's DEVICE headers (csrc/ea_core.cuh, lens_core.cuh) compiled for the host
with a one-lane "warp" (tests/hostsim/cuda_shim.h) and compared with the oracle.  This checks the
algorithmic logic of the kernels where no GPU exists; it is test scaffolding, not a CPU code path
of the product (the library has none), and says nothing about the real kernels' execution --
that is what the `-m gpu` tests are for."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, set_distance, C2_PARAMS
from oracle import lens, solver

HS = os.path.join(ROOT, "tests", "hostsim")
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def hs():
    so = os.path.join(HS, "libhostsim.so")
    srcs = [os.path.join(HS, "hostsim.cpp"), os.path.join(HS, "cuda_shim.h")] + [
        os.path.join(ROOT, "caustics_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "caustics_b200", "csrc"))
        if f.endswith(".cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                        os.path.join(HS, "hostsim.cpp"), "-lm"], check=True)
    return ctypes.CDLL(so)


def hs_solve(lib, c, comp=False, custom_init=False, ri=None, flags=0, itmax=2500):
    c = np.ascontiguousarray(c, dtype=complex)
    n, deg = c.shape[0], c.shape[1] - 1
    roots, sw = np.zeros((n, deg), complex), np.zeros(n, np.int32)
    rip = np.ascontiguousarray(ri, dtype=complex).ctypes.data_as(vp) if custom_init else None
    assert lib.hostsim_ea_solve(c.ctypes.data_as(vp), rip, roots.ctypes.data_as(vp), sw.ctypes.data_as(vp),
                                ctypes.c_int64(n), deg, itmax, int(comp), int(custom_init), flags) == 0
    return roots, sw


@pytest.mark.parametrize("name", ["fixture", "c1", "c2", "rand4", "rand6", "rand10"])
@pytest.mark.parametrize("comp", [False, True])
def test_device_solver_logic_vs_golden(hs, ea_golden, name, comp):
    c = ea_golden[name + "_coeffs"]
    c = c.reshape(-1, c.shape[-1])[:, ::-1]
    got, sw = hs_solve(hs, c, comp=comp)
    want = ea_golden[name + ("_roots_comp" if comp else "_roots_plain")]
    _, psw, _ = solver.port_solve(c, compensated=comp, return_stats=True)
    # reference-compatible init: the iteration path is the reference's -> same order, same sweeps
    if name != "fixture" and not comp:  # all-real polynomials: see the note in ea_core.cuh (Bini guesses there)
        assert (np.abs(sw) == psw).mean() > 0.97
    if comp and name != "fixture":  # two-stage schedule: plain sweeps, then polishing sweeps (never fewer than the reference's)
        assert (np.abs(sw) >= psw - 1).mean() > 0.97 and np.abs(sw).mean() < psw.mean() + 4
    assert (sw > 0).all()
    if comp:
        assert set_distance(got, ea_golden[name + "_roots_comp"]).max() < 1e-12
    if name == "fixture":
        assert set_distance(got, want).max() < 1e-12
    elif comp or name != "c2":
        assert np.abs(got - want).max() < 1e-12
    # Bini init: other order, same set
    got_b, sw_b = hs_solve(hs, c, comp=comp, flags=1)
    if comp:
        assert set_distance(got_b, ea_golden[name + "_roots_comp"]).max() < 1e-12
    assert np.abs(sw_b).mean() <= np.abs(sw).mean() + 0.2


def test_lazy_bound_is_bit_identical(hs, ea_golden):
    """CB200_LAZY_BOUND (ea_core.cuh) only skips work whose result cannot matter: the device solver compiled
    with and without it returns the same roots and the same sweep counts, bit for bit -- cold starts, warm
    starts, plain and compensated, reference-compatible and Bini estimates"""
    so = os.path.join(HS, "libhostsim_nolazy.so")
    srcs = [os.path.join(HS, "hostsim.cpp")] + [os.path.join(ROOT, "caustics_b200", "csrc", f)
                                                for f in os.listdir(os.path.join(ROOT, "caustics_b200", "csrc")) if f.endswith(".cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DCB200_LAZY_BOUND=0",
                        "-o", so, os.path.join(HS, "hostsim.cpp"), "-lm"], check=True)
    ref = ctypes.CDLL(so)
    for name in ("fixture", "c1", "c2", "rand6", "rand10"):
        c = ea_golden[name + "_coeffs"]
        c = c.reshape(-1, c.shape[-1])[:, ::-1]
        for comp in (False, True):
            for flags in (0, 1):
                a, sa = hs_solve(hs, c, comp=comp, flags=flags)
                b, sb = hs_solve(ref, c, comp=comp, flags=flags)
                assert np.array_equal(a.view(np.uint64), b.view(np.uint64)) and np.array_equal(sa, sb), (name, comp, flags)
            ri = np.roll(a, 1, axis=0) * (1 + 1e-4)          # warm start from a neighbour's roots (branchy step)
            a, sa = hs_solve(hs, c, comp=comp, custom_init=True, ri=ri)
            b, sb = hs_solve(ref, c, comp=comp, custom_init=True, ri=ri)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)) and np.array_equal(sa, sb), (name, comp, "warm")


def test_device_code_update_count_is_the_reference_algorithms(ea_golden):
    """the work the roofline is computed from: on a C2 sample the device solver (compiled for the host with work
    counters, -DCB200_HOSTSIM_COUNT) performs the reference algorithm's number of root updates -- 95.2 per polynomial,
    counted independently by the C oracle (test_oracle.py::test_c2_update_count) -- to a few parts per million, plus
    exactly one confirming evaluation per root"""
    so = os.path.join(HS, "libhostsim_count.so")
    srcs = [os.path.join(HS, "hostsim.cpp")] + [os.path.join(ROOT, "caustics_b200", "csrc", f)
                                                for f in os.listdir(os.path.join(ROOT, "caustics_b200", "csrc")) if f.endswith(".cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DCB200_HOSTSIM_COUNT",
                        "-o", so, os.path.join(HS, "hostsim.cpp"), "-lm"], check=True)
    lib = ctypes.CDLL(so)
    w = np.linspace(-2, 2, 1_000_000)[::500] + 0.1j
    c = np.ascontiguousarray(lens.poly_coeffs(w, 3, **C2_PARAMS)[:, ::-1])
    out = (ctypes.c_longlong * 2)()
    lib.hostsim_counters(out, 1)
    hs_solve(lib, c)
    lib.hostsim_counters(out, 1)
    evals, updates = out[0], out[1]
    _, _, st = solver.port_solve(c, return_stats=True)
    assert abs(updates - st[0]) <= 1e-4 * st[0]
    assert 95.0 < updates / len(c) < 95.4
    assert evals - updates == 10 * len(c)


def test_extreme_dynamic_range(hs):
    """(i) a polynomial whose largest coefficient is >= 2^1023 must not be scaled by +0.0;
    (ii) when EPS*b is so small that its square underflows (z^10 - 1e-140: roots of modulus 1e-14, stopping
    threshold ~1e-156) the stopping test compares unsquared moduli.  (Spans beyond ~1e+-150 between the
    moduli involved, e.g. z^10 - 1e-200, overflow the squared moduli inside the complex divisions as well;
    that range is documented as unsupported in ea_core.cuh -- lens polynomials are nowhere near it.)"""
    big = np.zeros((1, 6), complex)
    big[0] = [2.0**1023, -3 * 2.0**1020, 2.0**1000, 1.0, 0.5, 2.0**1023 * 0.75]     # low -> high
    got, sw = hs_solve(hs, big)
    want = np.roots((big[0] * 2.0**-1000)[::-1])[None, :]       # (the reference itself overflows on this input)
    assert np.isfinite(got).all() and sw[0] > 0
    assert set_distance(got, want).max() < 1e-10
    c = np.zeros((1, 11), complex)
    c[0, 0], c[0, 10] = -1e-140, 1.0
    got, sw = hs_solve(hs, c)
    exact = 1e-14 * np.exp(2j * np.pi * np.arange(10) / 10)
    assert 0 < sw[0] < 100                          # (squared compare: thr^2 = 0, spins until |h|^2 underflows or itmax)
    rel = np.abs(np.sort_complex(got[0]) - np.sort_complex(exact)) / 1e-14
    assert rel.max() < 1e-12
