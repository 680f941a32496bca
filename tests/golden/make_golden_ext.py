"""Extended-source golden vectors from the REFERENCE's own Python (through oracle/refshim.py) and its
compiled solver.  Run via tests/golden/make_golden.py (build container only)."""
import os

import numpy as np


def ext_golden(C, A, here):
    N = np.asarray
    out = {}
    rng = np.random.default_rng(42)

    def caustic_points(nl, rho, n, **hp):
        # the reference's own test-point recipe (tests/test_extended_source.py:83-108): points of the
        # caustics displaced by r ~ U(0, 2 rho), phi ~ U(-pi, pi); NumPy RNG instead of jax's
        _, zca = C.critical_and_caustic_curves(npts=50, nlenses=nl, **hp)
        ca = N(zca).reshape(-1)
        ca = ca[rng.choice(len(ca), n, replace=False)]
        return ca + rng.uniform(0, 2 * rho, n) * np.exp(1j * rng.uniform(-np.pi, np.pi, n))

    hp2 = dict(s=0.9, q=0.2)
    hp3 = dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0)
    mag = lambda w, rho, nl, hp, **kw: np.array(
        [float(C.mag_extended_source(complex(x), rho, nlenses=nl, **kw, **hp)) for x in w])
    for rho in (1e-1, 1e-2, 1e-3):
        w = caustic_points(2, rho, 40, **hp2)
        out[f"b_w_{rho}"] = w
        out[f"b_unif_{rho}"] = mag(w, rho, 2, hp2, npts_limb=200)
    w = out["b_w_0.01"][:16]
    out["b_ld_0.01"] = mag(w, 1e-2, 2, hp2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100)
    out["b_unif400_0.01"] = mag(w, 1e-2, 2, hp2, npts_limb=400)
    for rho in (1e-1, 1e-2):
        w = caustic_points(3, rho, 24, **hp3)
        out[f"t_w_{rho}"] = w
        out[f"t_unif_{rho}"] = mag(w, rho, 3, hp3, npts_limb=200)
    out["t_ld_0.01"] = mag(out["t_w_0.01"][:6], 1e-2, 3, hp3, npts_limb=200, limb_darkening=True, u1=0.3, npts_ld=60)
    # single lens (tests/test_extended_source.py:134-183 settings)
    ws = np.array([0.0 + 0.0j, 0.3, 0.9 + 0.2j, 1.5j, 2.5, 4.0 - 1j])
    for rho in (1.0, 1e-1, 1e-2):
        out[f"s_w_{rho}"] = ws * rho
        out[f"s_unif_{rho}"] = mag(ws * rho + 1e-9, rho, 1, {}, npts_limb=150)
    out["s_ld_0.1"] = mag(ws * 0.1 + 1e-9, 0.1, 1, {}, npts_limb=300, limb_darkening=True, u1=0.7, npts_ld=100)
    # light curve through `mag` (lightcurve.py:99-254): C3-like trajectory, coarse
    wl = np.linspace(-2, 2, 161) + 0.1j
    out["lc_w"] = wl
    out["lc_unif"] = N(C.mag(A(wl.copy()), 1e-2, nlenses=2, npts_limb=200, **hp2))
    out["lc_ld"] = N(C.mag(A(wl[40:120].copy()), 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100, **hp2))
    # known-answer vector of the reference's own test (tests/test_extended_source.py:186-206)
    seg = np.zeros((3, 25), dtype=np.complex128)
    np.savez_compressed(os.path.join(here, "ext_golden.npz"), **out)
