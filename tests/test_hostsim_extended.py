"""CPU tier: the DEVICE code of kernel family 3 (csrc/extended_core.cuh and friends) compiled for the
host (tests/hostsim, one-lane warps) against the oracle.  Checks the kernels' algorithmic logic --
limb walk, refinement, track matching, segment splitting, stitching, Green integrals, limb darkening,
hexadecapole gate -- where no GPU exists.  Test scaffolding only; the product has no CPU path."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import extended, lens
from test_hostsim import hs  # noqa: F401  (fixture: builds tests/hostsim/libhostsim.so)

vp, D_ = ctypes.c_void_p, ctypes.c_double
HP2 = dict(s=0.9, q=0.2)
HP3 = dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0)


def lens_const(nl, **p):
    r, eps = lens._lenses(nl, **p)
    r = np.concatenate([r, np.zeros(3 - len(r))]).astype(complex)
    eps = np.concatenate([eps, np.zeros(3 - len(eps))])
    one = lambda root: np.array([-root, 1.0 + 0j])
    H = np.array([1.0 + 0j])
    for ri in r[:nl]:
        H = lens._pmul(H, one(ri))
    G = np.zeros(1, complex)
    for j in range(nl):
        t = np.array([eps[j] + 0j])
        for i in range(nl):
            if i != j:
                t = lens._pmul(t, one(r[i]))
        G = lens._padd(G, t)
    Hp, Gp = np.zeros(4, complex), np.zeros(3, complex)
    Hp[:len(H)], Gp[:len(G)] = H, G
    return eps, r, Hp, Gp


def hs_ext(lib, w, rho, nl, hp, npts=200, ld=False, u1=0.0, npts_ld=100, gate=False, comp=False):
    p, xcm = lens.lens_params(nl, **hp)
    if nl == 1:
        eps, r, H, G = np.zeros(3), np.zeros(3, complex), np.zeros(4, complex), np.zeros(3, complex)
    else:
        eps, r, H, G = lens_const(nl, **p)
    w = np.ascontiguousarray(np.atleast_1d(w), dtype=complex)
    n = len(w)
    mag, test = np.zeros(n), np.zeros(n, np.uint8)
    rc = lib.hostsim_mag_extended(w.ctypes.data_as(vp), mag.ctypes.data_as(vp), test.ctypes.data_as(vp),
                                  ctypes.c_int64(n), D_(rho), nl, eps.ctypes.data_as(vp), r.ctypes.data_as(vp),
                                  H.ctypes.data_as(vp), G.ctypes.data_as(vp), D_(xcm), D_(hp.get("q", 1.0)),
                                  int(gate), npts, int(ld), D_(u1), npts_ld, 2500, int(comp), None, None)
    assert rc == 0
    return (mag, test.astype(bool)) if gate else mag


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "ext_golden.npz"))


@pytest.mark.parametrize("rho", [1e-1, 1e-2, 1e-3])
def test_binary_uniform(hs, g, rho):
    w = g[f"b_w_{rho}"][:16]
    want = np.array([extended.mag_extended_source(x, rho, 2, 200, **HP2) for x in w])
    got = hs_ext(hs, w, rho, 2, HP2)
    assert np.allclose(got, want, rtol=1e-8)
    assert np.abs(got / g[f"b_unif_{rho}"][:16] - 1).max() < 1e-9      # and the reference itself (same jitter stream)


def test_binary_limb_darkened(hs, g):
    w = g["b_w_0.01"][:6]
    want = np.array([extended.mag_extended_source(x, 1e-2, 2, 200, True, 0.7, 100, **HP2) for x in w])
    assert np.allclose(hs_ext(hs, w, 1e-2, 2, HP2, ld=True, u1=0.7), want, rtol=1e-8)


def test_triple(hs, g):
    w = g["t_w_0.01"][:8]
    want = np.array([extended.mag_extended_source(x, 1e-2, 3, 200, **HP3) for x in w])
    assert np.allclose(hs_ext(hs, w, 1e-2, 3, HP3), want, rtol=1e-8)
    want = np.array([extended.mag_extended_source(x, 1e-2, 3, 200, True, 0.3, 60, **HP3) for x in w[:3]])
    assert np.allclose(hs_ext(hs, w[:3], 1e-2, 3, HP3, ld=True, u1=0.3, npts_ld=60), want, rtol=1e-8)


def test_single_and_gate(hs, g):
    w = g["s_w_0.1"] + 1e-9
    want = np.array([extended.mag_extended_source(x, 0.1, 1, 150) for x in w])
    assert np.allclose(hs_ext(hs, w, 0.1, 1, {}, npts=150), want, rtol=1e-10)
    wl = g["lc_w"]
    want, t_want = extended.mag(wl, 1e-2, 2, 200, return_test=True, **HP2)
    got, t_got = hs_ext(hs, wl, 1e-2, 2, HP2, gate=True)
    assert (t_got == t_want).all()
    assert np.allclose(got[t_want], want[t_want], rtol=1e-10)          # hexadecapole: pure arithmetic
    assert np.allclose(got[~t_want], want[~t_want], rtol=1e-8)
    assert np.allclose(got, g["lc_unif"], rtol=1e-9)


@pytest.mark.parametrize("hp", [dict(s=1.5, q=0.5), dict(s=0.5, q=1.0), dict(s=1.2, q=1e-3)])
def test_other_geometries_and_radii(hs, hp):
    """wide / close / planetary binaries, source radii from 1e-4 to 1: device logic == oracle"""
    rng = np.random.default_rng(5)
    w = rng.uniform(-0.3, 0.3, 6) + 1j * rng.uniform(-0.3, 0.3, 6)
    for rho in (1.0, 5e-2, 1e-4):
        want = np.array([extended.mag_extended_source(x, rho, 2, 200, **hp) for x in w])
        assert np.allclose(hs_ext(hs, w, rho, 2, hp), want, rtol=1e-6)
    want, t_want = extended.mag(w, 5e-3, 2, 200, return_test=True, **hp)
    got, t_got = hs_ext(hs, w, 5e-3, 2, hp, gate=True)
    assert (t_got == t_want).all() and np.allclose(got, want, rtol=1e-7)


def hs_grid_walk(lib, x0, y0, dx, dy, nx, r0, r1, nl, hp, run=32, extrap=1, comp=False):
    p, xcm = lens.lens_params(nl, **hp)
    eps, r, H, G = lens_const(nl, **p)
    mag = np.zeros((r1 - r0, nx))
    rc = lib.hostsim_grid_walk(D_(x0), D_(y0), D_(dx), D_(dy), ctypes.c_int64(nx), ctypes.c_int64(r0), ctypes.c_int64(r1),
                               mag.ctypes.data_as(vp), nl, eps.ctypes.data_as(vp), r.ctypes.data_as(vp),
                               H.ctypes.data_as(vp), G.ctypes.data_as(vp), D_(xcm), 2500, int(comp), run, extrap)
    assert rc == 0
    return mag


@pytest.mark.parametrize("nl,hp,col0,r0", [(2, HP2, 4290, 5300), (3, HP3, 4350, 6180)])
def test_grid_walk(hs, nl, hp, col0, r0):
    """ps_walk.cuh (CAUSTICS_FLAG_GRID_WALK): warm-started column walks over a magnification map give the
    oracle's magnification at every pixel -- a patch of the C5 map that crosses the central caustic,
    with and without extrapolation, run lengths that do and do not divide the row count"""
    dx = 3.0 / 9999
    nx, r1 = 40, r0 + 41
    x0 = -1.5 + col0 * dx          # columns col0, col0 + 2, ...: a fold of the lens's caustic runs through the patch
    ix, iy = np.meshgrid(np.arange(nx), np.arange(r0, r1))
    w = (x0 + ix * (2 * dx)) + 1j * (-1.5 + iy * dx)
    want = lens.mag_point_source(w.reshape(-1), nl, **hp).reshape(w.shape)
    assert want.max() > 20          # the patch does contain near-caustic pixels
    for run, extrap in ((32, 1), (7, 1), (32, 0), (1, 1)):
        got = hs_grid_walk(hs, x0, -1.5, 2 * dx, dx, nx, r0, r1, nl, hp, run=run, extrap=extrap)
        rel = np.abs(got / want - 1)
        # rounding x conditioning: the same bound the cold kernel is held to on this map (test_c5_map_properties)
        assert rel.max() < 3e-9 and np.median(rel) < 1e-13, (run, extrap, rel.max())


def test_grid_walk_sweep_counts(hs):
    """what the walk buys: Ehrlich-Aberth sweeps per pixel from a cold start, from the previous row's roots
    and from the linear extrapolation of the previous two rows (the device solver compiled for the host)"""
    from test_hostsim import hs_solve
    p, x_cm = lens.lens_params(2, **HP2)
    dx = 3.0 / 9999
    rng = np.random.default_rng(1)
    xs = -1.5 + rng.integers(0, 10000, 500) * dx
    r0 = rng.integers(0, 10000 - 8, 500)
    sw = {"cold": [], "warm": [], "extrap": []}
    zw = ze = ze_prev = None
    for k in range(8):
        c = lens.poly_coeffs(xs + 1j * (-1.5 + (r0 + k) * dx) + x_cm, 2, **p)[:, ::-1]
        zc, s = hs_solve(hs, c, flags=1)
        sw["cold"].append(np.abs(s).mean())
        if k == 0:
            zw, ze, ze_prev = zc, zc, zc
            continue
        zw, s = hs_solve(hs, c, custom_init=True, ri=zw, flags=1)
        sw["warm"].append(np.abs(s).mean())
        z2, s = hs_solve(hs, c, custom_init=True, ri=2 * ze - ze_prev, flags=1)
        ze_prev, ze = ze, z2
        assert (s > 0).all()
        if k >= 2:
            sw["extrap"].append(np.abs(s).mean())
    assert np.mean(sw["cold"]) > 6.5 and np.mean(sw["warm"]) < 3.1 and np.mean(sw["extrap"]) < 2.15


def test_path_walk(hs):
    """CAUSTICS_FLAG_PATH_WALK device code: a trajectory solved as warm-started runs of consecutive
    elements == the oracle; an array that is NOT a path (random positions) is still solved correctly"""
    def walk(w, nl, hp, run, extrap=1):
        p, xcm = lens.lens_params(nl, **hp)
        eps, r, H, G = lens_const(nl, **p)
        w = np.ascontiguousarray(w, dtype=complex)
        mag = np.zeros(w.size)
        assert hs.hostsim_path_walk(w.ctypes.data_as(vp), mag.ctypes.data_as(vp), ctypes.c_int64(w.size), nl,
                                    eps.ctypes.data_as(vp), r.ctypes.data_as(vp), H.ctypes.data_as(vp),
                                    G.ctypes.data_as(vp), D_(xcm), 2500, 0, run, extrap) == 0
        return mag
    for nl, hp in ((2, HP2), (3, HP3)):
        w = np.linspace(-2, 2, 1_000_000)[400_000:404_000] + 0.1j      # C1/C2 spacing, inside the caustic region
        want = lens.mag_point_source(w, nl, **hp)
        for run in (8, 13):
            rel = np.abs(walk(w, nl, hp, run) / want - 1)
            assert rel.max() < 1e-9 and np.median(rel) < 1e-13, (nl, run, rel.max())
        rng = np.random.default_rng(3)
        wr = rng.uniform(-1.5, 1.5, 600) + 1j * rng.uniform(-1.5, 1.5, 600)
        rel = np.abs(walk(wr, nl, hp, 8) / lens.mag_point_source(wr, nl, **hp) - 1)
        assert rel.max() < 1e-8 and np.median(rel) < 1e-13, (nl, rel.max())


def test_walk_degenerate_inputs(hs):
    """walk device code on degenerate maps: zero step (every pixel the same), NaN / huge origins, itmax too small
    for any solve to converge -- terminates, and finite inputs still give the oracle's value"""
    dx = 3.0 / 9999
    same = hs_grid_walk(hs, 0.05, 0.1, 0.0, 0.0, 3, 0, 40, 2, HP2)
    want = lens.mag_point_source(np.array([0.05 + 0.1j]), 2, **HP2)[0]
    assert np.allclose(same, want, rtol=1e-10)
    for x0 in (np.nan, np.inf, 1e8, 1e-300):
        m = hs_grid_walk(hs, x0, 0.1, dx, dx, 4, 0, 9, 2, HP2, run=4)
        assert m.shape == (9, 4)
        if np.isfinite(x0):
            ix, iy = np.meshgrid(np.arange(4), np.arange(9))
            assert np.allclose(m, lens.mag_point_source(((x0 + ix * dx) + 1j * (0.1 + iy * dx)).reshape(-1), 2, **HP2).reshape(9, 4), rtol=1e-9)
    p, xcm = lens.lens_params(2, **HP2)
    eps, r, H, G = lens_const(2, **p)
    mag = np.zeros((20, 4))
    assert hs.hostsim_grid_walk(D_(-0.2), D_(0.05), D_(dx), D_(dx), ctypes.c_int64(4), ctypes.c_int64(0), ctypes.c_int64(20),
                                mag.ctypes.data_as(vp), 2, eps.ctypes.data_as(vp), r.ctypes.data_as(vp), H.ctypes.data_as(vp),
                                G.ctypes.data_as(vp), D_(xcm), 2, 0, 8, 1) == 0     # itmax = 2


def test_walk_row_blocks_aligned_to_the_run_are_the_unsharded_map(hs):
    """sharding.row_block(align=32): row blocks that start on walk boundaries reproduce the unsharded walked map
    bit for bit (what makes a walked map independent of the number of ranks)"""
    dx = 3.0 / 9999
    full = hs_grid_walk(hs, -0.25, 0.08, dx, dx, 6, 0, 80, 2, HP2, run=32)
    parts = [hs_grid_walk(hs, -0.25, 0.08, dx, dx, 6, lo, hi, 2, HP2, run=32) for lo, hi in ((0, 32), (32, 64), (64, 80))]
    assert np.array_equal(np.concatenate(parts), full)
    off = hs_grid_walk(hs, -0.25, 0.08, dx, dx, 6, 10, 50, 2, HP2, run=32)      # unaligned block: other walks,
    assert np.allclose(off, full[10:50], rtol=1e-9) and not np.array_equal(off, full[10:50])   # same map to rounding


@pytest.mark.parametrize("key,nl,hp", [("b", 2, HP2), ("t", 3, HP3)])
def test_grid_walk_vs_reference_map(hs, key, nl, hp):
    """the walk device code against the REFERENCE's own mag_point_source on two caustic-crossing patches of the
    C5 grid (tests/golden/map_golden.npz, generated by the reference's Python + compiled solver), and the
    oracle restatement against the same values"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "map_golden.npz"))
    x0, y0, dx, dy, nx, r0, r1 = g[key + "_spec"]
    nx, r0, r1 = int(nx), int(r0), int(r1)
    want = g[key + "_mag"]
    assert want.max() > 300          # pixels within a few 1e-6 of the fold
    # rtol 1e-10, widened next to the fold by the reference's own plain-vs-compensated spread and by the
    # conditioning of a fold image (mu ~ d^-1/2 in the distance d to the caustic: a relative 1e-16 in the
    # coefficients moves mu by ~1e-16 mu^2, i.e. 2e-10 at the mu = 1300 pixel of the binary patch)
    # (the degree-10 coefficients of the triple lens carry ~20x the rounding of the binary ones: 6e-10 at mu = 250)
    tol = 1e-10 + 10 * np.abs(want / g[key + "_mag_comp"] - 1) + (1e-15 if nl == 2 else 2e-14) * want**2
    ix, iy = np.meshgrid(np.arange(nx), np.arange(r0, r1))
    w = (x0 + ix * dx) + 1j * (y0 + iy * dy)
    assert (np.abs(lens.mag_point_source(w.reshape(-1), nl, **hp).reshape(w.shape) / want - 1) <= tol).all()
    for run, extrap in ((32, 1), (8, 1), (32, 0)):
        got = hs_grid_walk(hs, x0, y0, dx, dy, nx, r0, r1, nl, hp, run=run, extrap=extrap)
        rel = np.abs(got / want - 1)
        assert (rel <= tol).all() and np.median(rel) < 1e-13, (run, extrap, rel.max())


def test_device_jitters_are_the_references(hs):
    """csrc/jax_prng.cuh (compiled for the host) reproduces JAX's threefry draws of the reference's fixed keys bit
    for bit: the (deg, n) warm-start table and the (deg, npts) duplicate table equal oracle/jaxprng.py"""
    from oracle import jaxprng
    for deg, n, npts in ((5, 10, 200), (10, 10, 200), (5, 20, 400), (10, 64, 1280), (2, 7, 150)):
        limb, dup = np.zeros((deg, n), complex), np.zeros((deg, npts))
        hs.hostsim_jitters(deg, n, npts, limb.ctypes.data_as(vp), dup.ctypes.data_as(vp))
        assert np.array_equal(limb, jaxprng.limb_jitters(deg, n))
        assert np.array_equal(dup, jaxprng.duplicate_jitters(deg, npts))


def test_fused_tangent_device_logic(hs):
    """SURVEY 8 f2 on the CPU: the tangent accumulated inside contours_body (csrc/extended_core.cuh,
    GreenTangent) against what does not depend on it -- the closed form of a source centred on a single lens,
    d/d rho sqrt(1 + 4 / rho^2), and central differences of the forward pass for sources away from caustics
    (w, rho, and s, q through the chain a = s/2, e1 = 1/(1+q), x_cm = a (1-q)/(1+q))"""
    def with_grad(w, rho, nl, hp, **kw):
        w = np.atleast_1d(np.asarray(w, complex))
        gr = np.zeros((8, len(w)))
        hs.hostsim_set_grad(gr.ctypes.data_as(vp))
        try:
            m = hs_ext(hs, w, rho, nl, hp, **kw)
        finally:
            hs.hostsim_set_grad(None)
        return m, gr

    rho = 0.1
    m, gr = with_grad([1e-9 + 0j], rho, 1, {}, npts=300)
    assert abs(gr[7, 0] / (-4 / (rho**3 * np.sqrt(1 + 4 / rho**2))) - 1) < 1e-3          # 300-gon vs circle
    assert abs(gr[5, 0]) < 1e-3 and abs(gr[6, 0]) < 1e-3 and (gr[:5] == 0).all()          # symmetric; no lens parameters
    w0 = np.array([0.8 + 0.6j, -1.2 + 0.3j, 0.1 + 1.1j])
    s_, q_ = 0.9, 0.2
    m, gr = with_grad(w0, 1e-2, 2, HP2)
    assert np.array_equal(m, hs_ext(hs, w0, 1e-2, 2, HP2))                                # same forward value
    h = 1e-5
    fd = lambda f, x: (f(x + h) - f(x - h)) / (2 * h)
    assert np.allclose(gr[5], fd(lambda x: hs_ext(hs, w0 + x, 1e-2, 2, HP2), 0.0), rtol=1e-5, atol=1e-7)
    assert np.allclose(gr[6], fd(lambda x: hs_ext(hs, w0 + 1j * x, 1e-2, 2, HP2), 0.0), rtol=1e-5, atol=1e-7)
    assert np.allclose(gr[7], (hs_ext(hs, w0, 1e-2 + 1e-6, 2, HP2) - hs_ext(hs, w0, 1e-2 - 1e-6, 2, HP2)) / 2e-6, rtol=2e-3)
    ds = gr[0] * 0.5 + gr[5] * 0.5 * (1 - q_) / (1 + q_)
    dq = gr[1] * (-1 / (1 + q_)**2) + gr[5] * (s_ / 2) * (-2) / (1 + q_)**2
    assert np.allclose(ds, fd(lambda x: hs_ext(hs, w0, 1e-2, 2, dict(s=x, q=q_)), s_), rtol=1e-5, atol=1e-7)
    assert np.allclose(dq, fd(lambda x: hs_ext(hs, w0, 1e-2, 2, dict(s=s_, q=x)), q_), rtol=1e-5, atol=1e-7)
    assert (gr[2:5] == 0).all()                                                          # binary lens: no e2, r3


def test_adaptive_limb_darkening_quadrature(hs, g):
    """SURVEY 8 f4, opt-in (CAUSTICS_LD_ADAPTIVE): half-order Gauss-Legendre rule on short far panels of the P/Q
    integrals stays within 1e-4 of the reference's rule on the caustic-crossing golden sources, and the default
    (off) is untouched"""
    for w, rho, nl, hp, u1 in ((g["b_w_0.01"][:16], 1e-2, 2, HP2, 0.7), (g["b_w_0.1"][:8], 1e-1, 2, HP2, 0.7),
                               (g["t_w_0.01"][:4], 1e-2, 3, HP3, 0.3)):
        ref = hs_ext(hs, w, rho, nl, hp, ld=1, u1=u1)
        ada = hs_ext(hs, w, rho, nl, hp, ld=3, u1=u1)
        assert np.abs(ada / ref - 1).max() < 1e-4 and not np.array_equal(ada, ref)
    assert np.abs(hs_ext(hs, g["b_w_0.01"][:16], 1e-2, 2, HP2, ld=1, u1=0.7) / g["b_ld_0.01"] - 1).max() < 1e-9


def test_one_pass_path_equals_track_array_path(hs, g):
    """The plain uniform-disk call integrates closed tracks while it matches them (sweep_body) and stitches the
    open ones through the permutation (contours_body, resume); limb-darkened / tangent / export calls
    materialise the track arrays (tracks_body + contours_body).  Same sums in the same order: bit-identical."""
    hs.hostsim_force_tracks.argtypes = [ctypes.c_int]
    for nl, hp, w in ((2, HP2, g["b_w_0.01"][:24]), (3, HP3, g["t_w_0.01"][:10]), (1, {}, g["s_w_0.1"] + 1e-9)):
        rho = 0.1 if nl == 1 else 1e-2
        one = hs_ext(hs, w, rho, nl, hp)
        hs.hostsim_force_tracks(1)
        try:
            two = hs_ext(hs, w, rho, nl, hp)
        finally:
            hs.hostsim_force_tracks(0)
        assert np.array_equal(one, two), nl


@pytest.mark.parametrize("npts", [150, 400])
def test_other_limb_sampling(hs, g, npts):
    """other limb samplings: 150 gives an odd N0 and nadd = 7, 400 twenty new points per refinement round"""
    w = g["b_w_0.01"][:8]
    want = np.array([extended.mag_extended_source(x, 1e-2, 2, npts, **HP2) for x in w])
    assert np.allclose(hs_ext(hs, w, 1e-2, 2, HP2, npts=npts), want, rtol=1e-8)
