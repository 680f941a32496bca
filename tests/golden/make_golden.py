"""Regenerates tests/golden/*.npz from the REFERENCE itself.  Runs only where /root/reference is
mounted (the build container): the reference's C++ solver is compiled unmodified into oracle/_ref
and its Python is imported unmodified through oracle/refshim.py (a NumPy stand-in for jax).

    python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import refshim, solver  # noqa: E402

solver.build()
C = refshim.install()
A = refshim.arr
HERE = os.path.dirname(os.path.abspath(__file__))
N = np.asarray


def ref_roots(coeffs_high_low, compensated):
    return solver.ref_solve(np.ascontiguousarray(coeffs_high_low[:, ::-1]), itmax=2500,
                            compensated=compensated)


def solver_golden():
    out = {}
    ps = C.point_source
    # the reference's own fixture, tests/test_ehrlich_aberth_primitive.py:19-27
    a, e1 = 0.45, 0.8
    w = np.linspace(0.3, 0.35, 10).astype(np.complex128)
    out["fixture_coeffs"] = N(ps._poly_coeffs_binary(A(w), a, e1)).reshape(5, 2, 6)
    # C1: binary trajectory, every 20th point
    t = np.linspace(-2, 2, 10000)[::20]
    out["c1_coeffs"] = N(ps._poly_coeffs_binary(A(t + 0.1j + 0.3), 0.45, 1 / 1.2))
    # C2: triple trajectory
    t = np.linspace(-2, 2, 1000000)[::2000]
    out["c2_coeffs"] = N(ps._poly_coeffs_triple(A(t + 0.1j), 0.698, -0.0197 - 0.95087j, 0.02809, 0.9687))
    rng = np.random.default_rng(0)
    for deg in (4, 5, 6, 10):
        out[f"rand{deg}_coeffs"] = rng.standard_normal((200, deg + 1)) + 1j * rng.standard_normal((200, deg + 1))
    for k in list(out):
        c = out[k].reshape(-1, out[k].shape[-1])
        name = k[:-7]
        out[name + "_roots_plain"] = ref_roots(c, False)
        out[name + "_roots_comp"] = ref_roots(c, True)
    np.savez_compressed(os.path.join(HERE, "ea_golden.npz"), **out)


def ps_golden():
    out = {}
    ps, mp, lc = C.point_source, C.multipole, C.lightcurve
    rng = np.random.default_rng(1)
    w = rng.uniform(-1.5, 1.5, 400) + 1j * rng.uniform(-1.5, 1.5, 400)
    out["w"] = w
    out["binary_coeffs"] = N(ps._poly_coeffs_binary(A(w), 0.45, 1 / 1.2))
    out["triple_coeffs"] = N(ps._poly_coeffs_triple(A(w), 0.698, -0.0197 - 0.95087j, 0.02809, 0.9687))
    # tests/test_point_source.py:32-42 grid (coarser: 30x30)
    x = np.linspace(-0.5, 0.5, 30)
    wg = (x[:, None] + 1j * x[None, :]).reshape(-1)
    out["grid_w"] = wg
    out["grid_mag_binary"] = N(C.mag_point_source(A(wg.copy()), nlenses=2, s=0.9, q=0.2))
    out["mag_binary"] = N(C.mag_point_source(A(w.copy()), nlenses=2, s=0.9, q=0.2))
    hp3 = dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0)
    out["mag_triple"] = N(C.mag_point_source(A(w.copy()), nlenses=3, **hp3))
    out["mag_triple_comp"] = N(C.mag_point_source(A(w.copy()), nlenses=3, roots_compensated=True, **hp3))
    # images, hexadecapole and gate for the binary lens (lightcurve.py:202-225)
    a, e1 = 0.45, 1 / 1.2
    ws = w + 0.3
    z, zm = ps._images_point_source(A(ws), nlenses=2, a=a, e1=e1)
    out["images_z"], out["images_mask"] = N(z), N(zm)
    for rho in (1e-2, 1e-1):
        mu, dmu = mp._mag_hexadecapole(z, zm, rho, nlenses=2, a=a, e1=e1)
        t1 = lc._caustics_proximity_test(A(ws), z, zm, rho, dmu, nlenses=2, a=a, e1=e1)
        t2 = lc._planetary_caustic_test(A(ws), rho, a=a, e1=e1)
        out[f"hex_mu_{rho}"], out[f"hex_dmu_{rho}"] = N(mu), N(dmu)
        out[f"gate_{rho}"], out[f"planet_{rho}"] = N(t1), N(t2)
    mu, dmu = mp._mag_hexadecapole(z, zm, 0.05, u1=0.4, nlenses=2, a=a, e1=e1)
    out["hex_mu_ld"], out["hex_dmu_ld"] = N(mu), N(dmu)
    # critical curves and caustics, point_source.py:1582-1649
    for nl, hp in ((2, dict(s=0.9, q=0.2)), (3, hp3)):
        zcr, zca = C.critical_and_caustic_curves(npts=50, nlenses=nl, **hp)
        out[f"crit{nl}_cr"], out[f"crit{nl}_ca"] = N(zcr), N(zca)
    np.savez_compressed(os.path.join(HERE, "ps_golden.npz"), **out)


if __name__ == "__main__":
    solver_golden()
    ps_golden()
    if "ext" in sys.argv or len(sys.argv) == 1:
        try:
            from make_golden_ext import ext_golden
            ext_golden(C, A, HERE)
        except ImportError:
            pass
    print("golden vectors written to", HERE)
