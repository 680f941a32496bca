"""GPU tier: kernel 2 (fused point-source images / magnification) vs the oracle and the golden
vectors generated from the reference.  Tolerance: rtol 1e-10."""
import numpy as np
import pytest
import torch

from conftest import set_distance, c1_w, C2_PARAMS, TRIPLE_HP
from oracle import lens

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb(built_lib):
    import caustics_b200
    assert torch.cuda.is_available()
    return caustics_b200


def test_mag_binary_reference_grid(cb, ps_golden):
    """tests/test_point_source.py:32-49 grid, against the reference's own values"""
    w = ps_golden["grid_w"]
    got = cb.mag_point_source(torch.from_numpy(w).cuda(), nlenses=2, s=0.9, q=0.2).cpu().numpy()
    assert np.allclose(got, ps_golden["grid_mag_binary"], rtol=1e-10, atol=0)
    got = cb.mag_point_source(ps_golden["w"], nlenses=2, s=0.9, q=0.2)   # host arrays
    assert isinstance(got, np.ndarray)
    assert np.allclose(got, ps_golden["mag_binary"], rtol=1e-10, atol=0)


def test_mag_triple_reference(cb, ps_golden):
    w = torch.from_numpy(ps_golden["w"]).cuda()
    got = cb.mag_point_source(w, nlenses=3, **TRIPLE_HP).cpu().numpy()
    assert np.allclose(got, ps_golden["mag_triple"], rtol=1e-9, atol=0)
    got = cb.mag_point_source(w, nlenses=3, roots_compensated=True, **TRIPLE_HP).cpu().numpy()
    assert np.allclose(got, ps_golden["mag_triple_comp"], rtol=1e-9, atol=0)
    assert (np.abs(got / ps_golden["mag_triple_comp"] - 1) < 1e-10).mean() > 0.99


def test_c1_trajectory(cb):
    """config 1 through the fused kernel: magnification rtol 1e-10, image counts in {3, 5}"""
    w = c1_w()
    want = lens.mag_point_source(w, 2, **dict(s=0.9, q=0.2))
    for flags in (0, 1):
        got = cb.mag_point_source(torch.from_numpy(w).cuda(), nlenses=2, flags=flags, s=0.9, q=0.2).cpu().numpy()
        assert np.allclose(got, want, rtol=1e-10, atol=0)
    p, x_cm = lens.lens_params(2, s=0.9, q=0.2)
    z, mask = cb.point_source._images_point_source(torch.from_numpy(w + x_cm).cuda(), nlenses=2, **p)
    assert z.shape == (5, w.size) and mask.dtype == torch.bool
    counts = np.bincount(mask.sum(0).cpu().numpy(), minlength=6)
    assert counts[3] + counts[5] == w.size and counts[5] == 632
    zo, mo = lens.images_point_source(w + x_cm, 2, roots_compensated=True, **p)
    assert set_distance(z.cpu().numpy().T, zo.T).max() < 1e-10
    assert np.array_equal(np.sort(mask.cpu().numpy(), axis=0), np.sort(mo, axis=0))


def test_c2_triple_trajectory_subset(cb):
    """config 2 lens (tests/test_extended_source.py:100).  This lens has a 0.3 % third mass whose
    images are badly conditioned: the reference's own magnification moves by up to ~1e-6 when its
    polynomial coefficients are perturbed in the last bit, and some roots sit within a decade of
    the 1e-6 image threshold (DESIGN.md "conditioning of C2").  Parity is therefore asserted to
    rtol 1e-9 PLUS the oracle's own measured sensitivity at each point."""
    from oracle import solver
    w = np.linspace(-2, 2, 1000000)[::50] + 0.1j
    c = lens.poly_coeffs(w, 3, **C2_PARAMS)

    def oracle_mag(coeffs):
        z = solver.solve(np.ascontiguousarray(coeffs[:, ::-1]), itmax=2500, compensated=True).T
        res = np.abs(lens.lens_eq(z, 3, **C2_PARAMS) - w)
        inv_det = 1.0 / np.abs(lens.lens_eq_det_jac(z, 3, **C2_PARAMS))
        amb = (res > 1e-8) & (res < 1e-4)           # may be classified either way
        return ((res < 1e-6) * inv_det).sum(0), (amb * inv_det).sum(0)

    want, slack = oracle_mag(c)
    rng = np.random.default_rng(0)
    spread = np.zeros_like(want)
    for _ in range(4):                               # 1-ulp-level relative perturbations
        pert = c * (1 + 2.2e-16 * (rng.standard_normal(c.shape) + 1j * rng.standard_normal(c.shape)))
        m2, s2 = oracle_mag(pert)
        spread = np.maximum(spread, np.abs(m2 - want))
        slack = np.maximum(slack, s2)
    L = cb._lib.lib()
    lens_c = cb.point_source._c_lens(3, 0.0, **C2_PARAMS)
    wd = torch.from_numpy(w).cuda()
    mag = torch.empty(w.size, dtype=torch.float64, device="cuda")
    for comp in (1, 0):
        cb._lib.check(L.caustics_mag_point_source(wd.data_ptr(), mag.data_ptr(), None, w.size, lens_c, 2500, comp, 0, None))
        got = mag.cpu().numpy()
        tol = 1e-9 * want + 50 * spread + 1.001 * slack
        assert (np.abs(got - want) <= tol).all()
        assert np.abs(got / want - 1).max() < 1e-4
        well = (spread < 1e-12) & (slack == 0)       # well-conditioned points: the plain 1e-9 bar
        if well.any():
            assert np.allclose(got[well], want[well], rtol=1e-9, atol=0)


def test_images_custom_init_and_layout(cb):
    p, x_cm = lens.lens_params(3, **TRIPLE_HP)
    rng = np.random.default_rng(3)
    w = rng.uniform(-1, 1, (7, 33)) + 1j * rng.uniform(-1, 1, (7, 33))
    wd = torch.from_numpy(w).cuda()
    z, mask = cb.point_source._images_point_source(wd, nlenses=3, roots_compensated=True, **p)
    assert z.shape == (10, 7, 33) and mask.shape == (10, 7, 33)
    zo, mo = lens.images_point_source(w, 3, roots_compensated=True, **p)
    assert set_distance(z.reshape(10, -1).T.cpu().numpy(), zo.reshape(10, -1).T).max() < 1e-9
    # warm start from the solution itself: unchanged order
    z2, _ = cb.point_source._images_point_source(wd, nlenses=3, custom_init=True,
                                                 z_init=torch.movedim(z, 0, -1), **p)
    assert torch.abs(z2 - z).max().item() < 1e-9
    assert set(np.unique(mask.sum(0).cpu().numpy())) <= {4, 6, 8, 10}


def test_mag_grid_equals_explicit(cb):
    """magnification-map entry (config 5 shape): w generated on device == explicit w"""
    L = cb._lib.lib()
    p, x_cm = lens.lens_params(2, s=0.9, q=0.2)
    lens_c = cb.point_source._c_lens(2, x_cm, **p)
    nx, ny = 257, 64
    x0, y0, dx, dy = -1.5, -1.5, 3.0 / 9999, 3.0 / 9999
    out = torch.empty(nx * 20, dtype=torch.float64, device="cuda")
    cb._lib.check(L.caustics_mag_point_source_grid(x0, y0, dx, dy, nx, 10, 30, out.data_ptr(), lens_c, 2500, 0, 0, None))
    ix, iy = np.meshgrid(np.arange(nx), np.arange(10, 30))
    w = (x0 + ix * dx) + 1j * (y0 + iy * dy)
    want = cb.mag_point_source(torch.from_numpy(w.reshape(-1)).cuda(), nlenses=2, s=0.9, q=0.2)
    assert torch.allclose(out, want, rtol=1e-12, atol=0)


def test_single_lens_and_grad(cb):
    w = torch.tensor([0.3 + 0.1j, 1.2 - 0.4j], dtype=torch.complex128, device="cuda")
    u = torch.abs(w)
    assert torch.allclose(cb.mag_point_source(w, nlenses=1), (u**2 + 2) / (u * torch.sqrt(u**2 + 4)), rtol=1e-12)
    # tests/test_point_source.py:52-57: gradient of the magnification w.r.t. s
    s = torch.tensor(0.9, dtype=torch.float64, device="cuda", requires_grad=True)
    w = torch.tensor([0.05 + 0.3j, -0.4 + 0.2j, 0.6 - 0.1j], dtype=torch.complex128, device="cuda")
    m = cb.mag_point_source(w, nlenses=2, s=s, q=0.2)
    m.sum().backward()
    h = 1e-6
    fd = (cb.mag_point_source(w, nlenses=2, s=0.9 + h, q=0.2).sum() -
          cb.mag_point_source(w, nlenses=2, s=0.9 - h, q=0.2).sum()) / (2 * h)
    assert abs(s.grad.item() - fd.item()) < 1e-4 * max(1.0, abs(fd.item()))


@pytest.mark.parametrize("k", ["b", "t"])
def test_sequential_images(cb, seq_golden, k):
    """_images_point_source_sequential: warm-started scan along a path; rows follow images like the
    reference's (same Gauss-Seidel order from the same warm starts)"""
    from conftest import SEQ_PARAMS
    from caustics_b200.point_source import _images_point_source_sequential
    nl, p = SEQ_PARAMS[k]
    w, zg, mg = seq_golden[f"{k}_w"], seq_golden[f"{k}_z"], seq_golden[f"{k}_mask"]
    z, m = _images_point_source_sequential(w, nlenses=nl, **p)
    assert z.shape == zg.shape and np.array_equal(m, mg)
    # ordered; 1e-9: coefficient rounding x conditioning of the triple-lens polynomial (tests/test_oracle.py)
    assert np.abs(z - zg)[mg].max() < (1e-12 if k == "b" else 1e-9)
    for i in range(z.shape[1]):                      # every root as an unordered set
        assert set_distance(z[:, i][None], zg[:, i][None]).max() < 1e-9
    # batched paths: each path is independent; compensated agrees
    wb = torch.from_numpy(np.stack([w, w[::-1].copy(), w + 1e-3])).cuda()
    zb, mb = _images_point_source_sequential(wb, nlenses=nl, **p)
    assert zb.shape == (3,) + zg.shape and torch.equal(zb[0].cpu(), torch.from_numpy(z))
    zc, mc = _images_point_source_sequential(wb, nlenses=nl, roots_compensated=True, **p)
    assert torch.equal(mc[0].cpu(), torch.from_numpy(mg))
    if k == "b":     # (triple: the reference's plain roots next to the 2.8 % mass are themselves only good to ~1e-9)
        assert np.abs(zc[0].cpu().numpy() - zg)[mg].max() < 1e-12
    zo, mo = lens.images_point_source_sequential(w, nl, roots_compensated=True, **p)
    assert np.array_equal(mo, mg) and np.abs(zc[0].cpu().numpy() - zo)[mo].max() < (1e-12 if k == "b" else 1e-9)


def test_c5_map_properties(cb):
    """config 5 (10^4 x 10^4 binary magnification map), 400 rows = 4e6 points through the grid entry:
    mu >= 1 (a point source is never demagnified), mirror symmetry about the lens axis, image counts
    3 or 5, and the reference on a sparse subset"""
    import ctypes
    from caustics_b200 import _lib
    from caustics_b200.point_source import _c_lens, lens_params
    L = _lib.lib()
    hp = dict(s=0.9, q=0.2)
    p, x_cm = lens_params(2, **hp)
    lens_c = _c_lens(2, x_cm, **p)
    nx, dx = 10_000, 3.0 / 9999
    rows = 200
    out = {}
    for name, r0 in (("lo", 4700), ("hi", 10_000 - 4700 - rows)):      # rows symmetric about y = 0
        mag = torch.empty(nx * rows, dtype=torch.float64, device="cuda")
        _lib.check(L.caustics_mag_point_source_grid(-1.5, -1.5, dx, dx, nx, r0, r0 + rows, mag.data_ptr(), lens_c,
                                                    2500, 0, 0, None))
        out[name] = mag.reshape(rows, nx)
    torch.cuda.synchronize()
    lo, hi = out["lo"], out["hi"].flip(0)
    assert torch.isfinite(lo).all() and (lo >= 1.0 - 1e-12).all()
    # y -> -y: the coefficients are complex conjugates, the roots too; agreement to rounding x conditioning
    rel = ((lo - hi).abs() / lo)
    assert rel.max() < 1e-6 and rel.median() < 1e-13
    # sparse comparison with the oracle
    iy = np.arange(0, rows, 40); ix = np.arange(0, nx, 500)
    ww = (-1.5 + ix[None, :] * dx) + 1j * (-1.5 + (4700 + iy[:, None]) * dx)
    want = lens.mag_point_source(ww.reshape(-1), 2, **hp).reshape(ww.shape)
    got = lo[iy][:, ix].cpu().numpy()
    assert np.allclose(got, want, rtol=1e-9)


@pytest.mark.parametrize("nl,hp,r0", [(2, dict(s=0.9, q=0.2), 5250), (3, dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0), 6150)])
def test_map_walk(cb, nl, hp, r0):
    """CAUSTICS_FLAG_GRID_WALK (csrc/ps_walk.cuh): the C5 map solved as warm-started column walks against
    the cold per-pixel kernel on 100 full-width rows that cross the caustic (10^6 pixels), and against the
    oracle on a sparse subset"""
    nx, dx, rows = 10_000, 3.0 / 9999, 100
    cold = cb.mag_point_source_map(-1.5, -1.5, dx, dx, nx, nx, nlenses=nl, rows=(r0, r0 + rows), walk=False, **hp)
    walk = cb.mag_point_source_map(-1.5, -1.5, dx, dx, nx, nx, nlenses=nl, rows=(r0, r0 + rows), walk=True, **hp)
    assert cold.shape == walk.shape == (rows, nx)
    assert cold.max().item() > 30                       # caustic crossings are in the block
    rel = ((walk - cold).abs() / cold).reshape(-1)
    assert torch.isfinite(walk).all() and (walk >= 1.0 - 1e-12).all()
    # rounding x conditioning: the bound the cold kernel meets against its own mirror image (test_c5_map_properties)
    assert rel.median().item() < 1e-13 and rel.max().item() < 1e-6
    assert (rel > 1e-10).float().mean().item() < 2e-3
    iy = np.arange(0, rows, 9); ix = np.arange(0, nx, 250)
    ww = (-1.5 + ix[None, :] * dx) + 1j * (-1.5 + (r0 + iy[:, None]) * dx)
    want = lens.mag_point_source(ww.reshape(-1), nl, **hp).reshape(ww.shape)
    assert np.allclose(walk[iy][:, ix].cpu().numpy(), want, rtol=1e-9)


def test_map_walk_shapes(cb):
    """walks shorter than a run, row counts the run length does not divide, widths below one CTA, a
    compensated solve, an empty block; every pixel against the cold kernel"""
    hp = dict(s=0.9, q=0.2)
    dx = 3.0 / 9999
    for nx, ny, rows in ((257, 64, (10, 30)), (5, 200, (0, 200)), (1000, 37, (0, 37)), (129, 3, (1, 2))):
        kw = dict(nlenses=2, rows=rows, **hp)
        a = cb.mag_point_source_map(-0.2, 0.05, dx, dx, nx, ny, walk=False, **kw)
        b = cb.mag_point_source_map(-0.2, 0.05, dx, dx, nx, ny, walk=True, **kw)
        assert a.shape == b.shape == (rows[1] - rows[0], nx)
        assert torch.allclose(a, b, rtol=1e-7, atol=0) and ((a - b).abs() / a).median().item() < 1e-13
    c = cb.mag_point_source_map(-0.2, 0.05, dx, dx, 257, 64, nlenses=2, roots_compensated=True, **hp)
    d = cb.mag_point_source_map(-0.2, 0.05, dx, dx, 257, 64, nlenses=2, roots_compensated=True, walk=False, **hp)
    assert torch.allclose(c, d, rtol=1e-7, atol=0)
    assert cb.mag_point_source_map(-0.2, 0.05, dx, dx, 257, 64, nlenses=2, rows=(5, 5), **hp).shape == (0, 257)
    with pytest.raises(ValueError):
        cb.mag_point_source_map(0, 0, dx, dx, 10, 10, nlenses=1)


def test_path_walk(cb):
    """CAUSTICS_FLAG_PATH_WALK: 10^6-point trajectories (C1 binary lens; a well-conditioned triple lens) as
    warm-started runs against the default per-point kernel and the oracle; a shuffled array (not a path)
    still gives the default kernel's answer; short arrays take the default kernel (bit-identical)"""
    n = 1_000_000
    w = np.linspace(-2, 2, n) + 0.1j
    wd = torch.from_numpy(w).cuda()
    for nl, hp in ((2, dict(s=0.9, q=0.2)), (3, TRIPLE_HP)):
        cold = cb.mag_point_source(wd, nlenses=nl, **hp)
        walk = cb.mag_point_source(wd, nlenses=nl, flags=8, **hp)
        rel = (walk - cold).abs() / cold
        assert torch.isfinite(walk).all()
        assert rel.median().item() < 1e-13 and rel.max().item() < 1e-6 and (rel > 1e-10).float().mean().item() < 2e-3
        sel = np.arange(0, n, 997)
        assert np.allclose(walk.cpu().numpy()[sel], lens.mag_point_source(w[sel], nl, **hp), rtol=1e-9)
        perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
        shuf = cb.mag_point_source(wd[perm], nlenses=nl, flags=8, **hp)
        rel = (shuf - cold[perm]).abs() / cold[perm]
        assert rel.median().item() < 1e-13 and rel.max().item() < 1e-6
        short = wd[:100_000]
        assert torch.equal(cb.mag_point_source(short, nlenses=nl, flags=8, **hp), cb.mag_point_source(short, nlenses=nl, **hp))


def test_path_walk_small_forced(cb):
    """the path-walk kernel on arrays whose length the run does not divide (caustics_set_tuning("path_run")
    forces the kernel for a batch that would otherwise take the default one), plus run lengths longer than
    the array"""
    from caustics_b200 import _lib
    tune = _lib.lib().caustics_set_tuning
    hp = dict(s=0.9, q=0.2)
    for n, run in ((1003, 8), (130, 32), (5, 16), (1, 4), (4097, 3)):
        w = torch.from_numpy(np.linspace(-0.5, 0.5, n) + 0.1j).cuda()
        for nl, h in ((2, hp), (3, TRIPLE_HP)):
            assert tune(b"path_run", run) == 0
            try:
                walk = cb.mag_point_source(w, nlenses=nl, flags=8, **h)
            finally:
                tune(b"path_run", -1)
            cold = cb.mag_point_source(w, nlenses=nl, **h)
            assert walk.shape == cold.shape and torch.allclose(walk, cold, rtol=1e-7, atol=0), (n, run, nl)
            assert ((walk - cold).abs() / cold).median().item() < 1e-13
