"""Checks against truths that do NOT come from the reference: arbitrary-precision roots (mpmath) and
a brute-force area integral of the point-source magnification over the source disk.  They bound the
one thing the reference-vs-reference parity cannot: that the reference algorithm (and so the oracle
and the kernels) computes the right physical quantity.  The reference's own external truth
(MulensModel / VBBinaryLensing) is not installable here."""
import numpy as np
import pytest

from conftest import set_distance
from oracle import extended, lens, solver

HP2 = dict(s=0.9, q=0.2)


def _mp_roots(coeffs_high_low):
    import mpmath as mp
    mp.mp.dps = 60
    out = []
    for row in coeffs_high_low:
        r = mp.polyroots([mp.mpc(c.real, c.imag) for c in row], maxsteps=500, extraprec=400)
        out.append([complex(x) for x in r])
    return np.array(out)


def test_oracle_roots_vs_mpmath(ea_golden):
    c = np.concatenate([ea_golden["fixture_coeffs"].reshape(-1, 6)[:6], ea_golden["c1_coeffs"][::80]])
    truth = _mp_roots(c)
    got = solver.solve(np.ascontiguousarray(c[:, ::-1]), compensated=True)
    assert set_distance(got, truth).max() < 1e-13
    c10 = ea_golden["rand10_coeffs"][:8]
    assert set_distance(solver.solve(np.ascontiguousarray(c10[:, ::-1]), compensated=True), _mp_roots(c10)).max() < 1e-12


def _disk_average_ps(w0, rho, nr=24, nth=96):
    """(1 / pi rho^2) * integral of the point-source magnification over the disk, Gauss-Legendre in
    r^2 and the periodic trapezoid rule in theta (spectrally accurate for a smooth integrand)."""
    x, wt = np.polynomial.legendre.leggauss(nr)
    u = 0.5 * (x + 1.0)                      # u = (r / rho)^2 in (0, 1): dA = pi rho^2 du dtheta / (2 pi)
    th = 2 * np.pi * np.arange(nth) / nth
    pts = w0 + rho * np.sqrt(u)[:, None] * np.exp(1j * th)[None, :]
    mu = lens.mag_point_source(pts.reshape(-1), 2, roots_compensated=True, **HP2).reshape(nr, nth)
    return float((0.5 * wt[:, None] * mu).sum() / nth)


@pytest.mark.parametrize("w0,rho", [(0.6 + 0.45j, 0.05), (-0.9 + 0.3j, 0.1), (0.1 + 0.9j, 0.02)])
def test_contour_integration_vs_area_integral(w0, rho):
    """source disks that do not touch a caustic: Green's-theorem magnification == area average of the
    point-source magnification (uniform disk)"""
    truth = _disk_average_ps(w0, rho)
    assert abs(_disk_average_ps(w0, rho, 32, 128) / truth - 1) < 1e-9          # quadrature converged
    for npts, tol in ((200, 5e-4), (1000, 3e-5)):
        got = extended.mag_extended_source(w0, rho, 2, npts, **HP2)
        assert abs(got / truth - 1) < tol


def _disk_average_ps_ld(w0, rho, u1, nr=40, nth=128):
    """limb-darkened version: weight I(r) = 3/(3-u1) (1 - u1 + u1 sqrt(1 - r^2)) (integrate.py:29-44,
    unit mean over the disk); substitution (r/rho)^2 = 1 - t^2 removes the square-root end point"""
    x, wt = np.polynomial.legendre.leggauss(nr)
    t = 0.5 * (x + 1.0)
    u = 1.0 - t**2
    th = 2 * np.pi * np.arange(nth) / nth
    pts = w0 + rho * np.sqrt(u)[:, None] * np.exp(1j * th)[None, :]
    mu = lens.mag_point_source(pts.reshape(-1), 2, roots_compensated=True, **HP2).reshape(nr, nth)
    I = 3.0 / (3.0 - u1) * (1.0 - u1 + u1 * t)
    return float((0.5 * wt * 2 * t * I)[:, None].__mul__(mu).sum() / nth)


def test_limb_darkened_contour_integration_vs_area_integral():
    w0, rho, u1 = 0.6 + 0.45j, 0.05, 0.7
    truth = _disk_average_ps_ld(w0, rho, u1)
    assert abs(_disk_average_ps_ld(w0, rho, u1, 56, 160) / truth - 1) < 1e-8
    got = extended.mag_extended_source(w0, rho, 2, 400, True, u1, 100, **HP2)
    assert abs(got / truth - 1) < 3e-4
    # and the weight really has unit mean: u1 only redistributes light
    assert abs(_disk_average_ps_ld(w0 + 50.0, 1e-3, u1) - 1.0) < 1e-6


@pytest.mark.gpu
def test_kernels_vs_independent_truth(ea_golden):
    import torch
    import caustics_b200 as cb
    c = np.concatenate([ea_golden["fixture_coeffs"].reshape(-1, 6)[:6], ea_golden["c1_coeffs"][::80]])
    got = cb.poly_roots(torch.from_numpy(c).cuda(), itmax=2500, compensated=True).cpu().numpy()
    assert set_distance(got, _mp_roots(c)).max() < 1e-13
    for w0, rho in ((0.6 + 0.45j, 0.05), (-0.9 + 0.3j, 0.1)):
        truth = _disk_average_ps(w0, rho)
        assert abs(cb.mag_extended_source(w0, rho, nlenses=2, npts_limb=1000, **HP2) / truth - 1) < 3e-5
        # limb darkening with u1 = 0 is the uniform disk
        assert abs(cb.mag_extended_source(w0, rho, nlenses=2, npts_limb=1000, limb_darkening=True, u1=0.0, **HP2) / truth - 1) < 3e-5
    truth = _disk_average_ps_ld(0.6 + 0.45j, 0.05, 0.7)
    got = cb.mag_extended_source(0.6 + 0.45j, 0.05, nlenses=2, npts_limb=400, limb_darkening=True, u1=0.7, **HP2)
    assert abs(got / truth - 1) < 3e-4
