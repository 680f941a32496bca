"""GPU tier: kernel 1 (Ehrlich-Aberth) through the C ABI vs the oracle and the golden vectors.
Tolerances are BASELINE.md section 3: roots as unordered sets <= 1e-12 relative against the
COMPENSATED reference; residual |p(z)| <= 1e-10 on the reference fixture."""
import numpy as np
import pytest
import torch

from conftest import set_distance, c1_w, C2_PARAMS
from oracle import lens, solver

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb(built_lib):
    import caustics_b200
    assert torch.cuda.is_available()
    return caustics_b200


def _polyval_high_low(c, z):
    out = np.zeros_like(z)
    for k in range(c.shape[-1]):
        out = out * z + c[..., k:k + 1]
    return out


@pytest.mark.parametrize("comp", [False, True])
def test_poly_roots_fixture(cb, ea_golden, comp):
    """tests/test_ehrlich_aberth_primitive.py:30-35 on its own fixture, shape (5, 2, 6)"""
    coeffs = ea_golden["fixture_coeffs"]
    roots = cb.poly_roots(torch.from_numpy(coeffs).cuda(), compensated=comp).cpu().numpy()
    assert roots.shape == (5, 2, 5)
    assert np.abs(_polyval_high_low(coeffs, roots)).max() < 1e-10
    want = ea_golden["fixture_roots_comp"]
    assert set_distance(roots.reshape(-1, 5), want).max() < 1e-12
    # host (numpy) entry point gives the same roots
    roots_h = cb.poly_roots(coeffs, compensated=comp)
    assert isinstance(roots_h, np.ndarray) and np.array_equal(roots_h, roots)


@pytest.mark.parametrize("name", ["c1", "c2", "rand4", "rand5", "rand6", "rand10"])
@pytest.mark.parametrize("comp", [False, True])
def test_roots_vs_reference_golden(cb, ea_golden, name, comp):
    c = ea_golden[name + "_coeffs"]
    c = c.reshape(-1, c.shape[-1])
    got, sweeps = cb.primitive._solve_flat(torch.from_numpy(c).cuda(), None, 2500, comp, False,
                                           cb._lib.FLAG_COEFFS_HIGH_FIRST, return_sweeps=True)
    got, sweeps = got.cpu().numpy(), sweeps.cpu().numpy()
    assert (sweeps > 0).all()                      # every polynomial converged
    want = ea_golden[name + ("_roots_comp" if comp else "_roots_plain")]
    if comp:
        assert set_distance(got, ea_golden[name + "_roots_comp"]).max() < 1e-12
    if comp or name != "c2":                        # (plain c2: ill-conditioned)
        assert np.abs(got - want).max() < 1e-11    # same root ORDER as the reference, too
    # sweep counts follow the reference's iteration path
    _, psw, _ = solver.port_solve(np.ascontiguousarray(c[:, ::-1]), compensated=comp, return_stats=True)
    if not comp:
        assert (sweeps == psw).mean() > 0.95
    else:   # two-stage schedule (plain sweeps, then polishing sweeps): a few sweeps more than interleaved
        assert sweeps.mean() < psw.mean() + 4


@pytest.mark.parametrize("deg", [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 16])
def test_all_degrees_random(cb, deg):
    rng = np.random.default_rng(deg)
    n = 4099  # ragged tail: not a multiple of the CTA size
    c = rng.standard_normal((n, deg + 1)) + 1j * rng.standard_normal((n, deg + 1))
    want = solver.solve(c, compensated=True)
    for flags in (0, 1):
        got = cb.ehrlich_aberth(torch.from_numpy(c).cuda(), torch.empty(0), itmax=2500, compensated=True,
                                flags=flags).reshape(n, deg).cpu().numpy()
        assert set_distance(got, want).max() < 1e-12
    got = cb.ehrlich_aberth(torch.from_numpy(c).cuda(), torch.empty(0), itmax=2500, compensated=False)
    assert got.shape == (n * deg,)                  # flat, like the primitive
    assert set_distance(got.reshape(n, deg).cpu().numpy(), want).max() < 1e-9


def test_c1_full_size(cb):
    """config 1: 10^4 degree-5 polynomials of the binary trajectory, plain + compensated"""
    p, x_cm = lens.lens_params(2, s=0.9, q=0.2)
    c = lens.poly_coeffs(c1_w() + x_cm, 2, **p)
    want = solver.solve(np.ascontiguousarray(c[:, ::-1]), compensated=True)
    for comp in (False, True):
        got = cb.poly_roots(torch.from_numpy(c).cuda(), itmax=2500, compensated=comp).cpu().numpy()
        assert set_distance(got, want).max() < 1e-12


def test_custom_init_continuity(cb):
    """custom_init: root j stays the continuation of roots_init[j] and the
    result equals the reference's on the same warm start."""
    p, x_cm = lens.lens_params(2, s=0.9, q=0.2)
    w = c1_w(4000) + x_cm
    c = lens.poly_coeffs(w, 2, **p)
    z_prev = solver.solve(np.ascontiguousarray(c[:-1, ::-1]), compensated=True)
    want = solver.solve(np.ascontiguousarray(c[1:, ::-1]), custom_init=True, roots_init=z_prev)
    got = cb.poly_roots(torch.from_numpy(c[1:]).cuda(), itmax=2500, custom_init=True,
                        roots_init=torch.from_numpy(z_prev).cuda()).cpu().numpy()
    assert np.abs(got - want).max() < 1e-10         # ordered: same track per column
    got_h = cb.poly_roots(c[1:], itmax=2500, custom_init=True, roots_init=z_prev)
    assert np.array_equal(got_h, got)


def test_itmax_and_empty_and_dtype(cb):
    c = np.random.default_rng(0).standard_normal((64, 6)) + 0j
    c[:, 1] += 1j
    got, sw = cb.primitive._solve_flat(torch.from_numpy(c).cuda(), None, 2, False, False, 0, return_sweeps=True)
    assert (sw.cpu().numpy() == -2).all()            # itmax reached: reported, not an error
    assert np.isfinite(got.cpu().numpy().view(float)).all()
    out = cb.poly_roots(torch.empty((0, 6), dtype=torch.complex128, device="cuda"))
    assert out.shape == (0, 5)
    with pytest.raises(NotImplementedError):
        cb.poly_roots(torch.ones((4, 6), dtype=torch.complex64, device="cuda"))
    from caustics_b200._lib import CausticsError
    with pytest.raises(CausticsError):
        cb.poly_roots(torch.ones((4, 19), dtype=torch.complex128, device="cuda"))


def test_real_coefficients_and_scaling(cb):
    """all-real polynomials (sources on the lens axis) and badly scaled coefficients"""
    rng = np.random.default_rng(7)
    c = rng.standard_normal((512, 6)) + 0j
    want = solver.solve(c, compensated=True)
    got, sw = cb.primitive._solve_flat(torch.from_numpy(c).cuda(), None, 2500, True, False, 0, return_sweeps=True)
    assert (sw.cpu().numpy() > 0).all() and sw.max().item() < 40
    assert set_distance(got.cpu().numpy(), want).max() < 1e-12
    for scale in (1e-200, 1e200):
        got = cb.primitive._solve_flat(torch.from_numpy(c * scale).cuda(), None, 2500, True, False, 0)
        assert set_distance(got.cpu().numpy(), want).max() < 1e-12


def test_xla_custom_call_entry(cb, ea_golden):
    """caustics_ea_xla: the symbol an XLA custom call binds (buffers + opaque descriptor)"""
    import ctypes
    L = cb._lib.lib()
    c = np.ascontiguousarray(ea_golden["rand5_coeffs"][:, ::-1])
    dc = torch.from_numpy(c).cuda()
    dri = torch.zeros((c.shape[0], 5), dtype=torch.complex128, device="cuda")
    out = torch.empty((c.shape[0] * 5,), dtype=torch.complex128, device="cuda")
    d = cb._lib.EADescriptor()
    n = L.caustics_ea_make_descriptor(ctypes.byref(d), c.shape[0], 5, 2500, 1, 0, 0)
    bufs = (ctypes.c_void_p * 3)(dc.data_ptr(), dri.data_ptr(), out.data_ptr())
    L.caustics_ea_xla(torch.cuda.current_stream().cuda_stream, bufs, bytes(d), n)
    assert L.caustics_last_xla_error() == 0
    torch.cuda.synchronize()
    assert set_distance(out.cpu().numpy().reshape(-1, 5), ea_golden["rand5_roots_comp"]).max() < 1e-12


def test_autograd_matches_finite_differences(cb):
    """tests/test_ehrlich_aberth_primitive.py:59-64: gradients atol=rtol=1e-4"""
    rng = np.random.default_rng(11)
    c = torch.from_numpy(rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6))).cuda().requires_grad_()
    wgt = torch.from_numpy(rng.standard_normal((6, 5)) + 1j * rng.standard_normal((6, 5))).cuda()
    def f(cc):
        z = cb.poly_roots(cc, compensated=True)
        return (z * wgt).real.sum() + (z * z).imag.sum()
    f(c).backward()
    g = c.grad.clone()
    h = 1e-6
    base = c.detach()
    z0 = cb.poly_roots(base, compensated=True)
    for idx in [(0, 0), (2, 3), (5, 5)]:
        for d in (1.0, 1j):
            e = torch.zeros_like(base); e[idx] = d * h
            def fz(cc):
                z = cb.poly_roots(cc, compensated=True, custom_init=True, roots_init=z0)
                return ((z * wgt).real.sum() + (z * z).imag.sum()).item()
            fd = (fz(base + e) - fz(base - e)) / (2 * h)
            an = g[idx].real.item() if d == 1.0 else g[idx].imag.item()
            assert abs(fd - an) <= 1e-4 + 1e-4 * abs(fd)


def test_batch_layout_invariance(cb, ea_golden):
    """A polynomial's roots do not depend on which other polynomials share its warp or on chunking:
    same bits for a shuffled batch, for the host-pipeline chunks and for a one-polynomial call."""
    rng = np.random.default_rng(99)
    c = np.concatenate([ea_golden["c2_coeffs"], ea_golden["rand10_coeffs"],
                        rng.standard_normal((3000, 11)) * np.exp(rng.uniform(-3, 3, (3000, 1))) + 1j * rng.standard_normal((3000, 11))])
    base = cb.poly_roots(torch.from_numpy(c).cuda(), itmax=2500).cpu().numpy()
    perm = rng.permutation(len(c))
    shuf = cb.poly_roots(torch.from_numpy(np.ascontiguousarray(c[perm])).cuda(), itmax=2500).cpu().numpy()
    assert np.array_equal(shuf, base[perm])
    assert np.array_equal(cb.poly_roots(np.ascontiguousarray(c), itmax=2500), base)       # host chunks
    for i in (0, 517, len(c) - 1):
        one = cb.poly_roots(torch.from_numpy(c[i:i + 1].copy()).cuda(), itmax=2500).cpu().numpy()
        assert np.array_equal(one[0], base[i])


def test_jvp_vjp_kernels(cb):
    """on-device tangent / cotangent vs the Python rule of the reference"""
    rng = np.random.default_rng(21)
    for deg in (5, 10):
        c = rng.standard_normal((300, deg + 1)) + 1j * rng.standard_normal((300, deg + 1))
        dc = rng.standard_normal((300, deg + 1)) + 1j * rng.standard_normal((300, deg + 1))
        z = solver.solve(c, compensated=True)
        want = lens.jvp_roots(c, z, dc)
        got = cb.roots_jvp(torch.from_numpy(c).cuda(), torch.from_numpy(z).cuda(), torch.from_numpy(dc).cuda())
        assert np.allclose(got.cpu().numpy(), want, rtol=1e-10, atol=1e-12)
        # VJP is the conjugate transpose of the JVP: Re<g, J dp> == Re<J^H g, dp>
        g = rng.standard_normal((300, deg)) + 1j * rng.standard_normal((300, deg))
        ct = torch.from_numpy(c).cuda().requires_grad_()
        zz = cb.primitive._PolyRoots.apply(ct, None, 2500, True, False, 0)
        (zz * torch.from_numpy(np.conj(g)).cuda()).real.sum().backward()
        lhs = np.real(np.sum(np.conj(g) * want))
        rhs = np.real(np.sum(np.conj(ct.grad.cpu().numpy()) * dc))
        assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))


def test_lazy_conj_inputs(cb):
    """torch marks conj()/neg as lazy bits; the ABI must see the materialised values"""
    rng = np.random.default_rng(5)
    c = torch.from_numpy(rng.standard_normal((64, 6)) + 1j * rng.standard_normal((64, 6))).cuda()
    a = cb.poly_roots(torch.conj(c), compensated=True)
    b = cb.poly_roots(torch.conj(c).clone(), compensated=True)
    assert torch.equal(a, b)
    w = torch.from_numpy(rng.uniform(-1, 1, 50) + 1j * rng.uniform(-1, 1, 50)).cuda()
    assert torch.equal(cb.mag_point_source(torch.conj(w), nlenses=2, s=0.9, q=0.2),
                       cb.mag_point_source(torch.conj(w).clone(), nlenses=2, s=0.9, q=0.2))


def test_pathological_inputs_terminate(cb):
    """NaN / Inf / all-zero / zero-leading-coefficient polynomials: bounded by itmax, no hang, the
    healthy polynomials in the same warps are unaffected"""
    rng = np.random.default_rng(3)
    c = rng.standard_normal((256, 6)) + 1j * rng.standard_normal((256, 6))
    good = c.copy()
    c[3] = np.nan
    c[40] = 0.0
    c[77, 0] = 0.0            # zero leading coefficient (high -> low order)
    c[100, -1] = 0.0          # zero constant term: a root at the origin
    c[130, 2] = np.inf
    out, sw = cb.primitive._solve_flat(torch.from_numpy(c).cuda(), None, 60, False, False,
                                       cb._lib.FLAG_COEFFS_HIGH_FIRST, return_sweeps=True)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    ref = cb.poly_roots(torch.from_numpy(good).cuda(), itmax=60).cpu().numpy()
    keep = np.setdiff1d(np.arange(256), [3, 40, 77, 100, 130])
    assert np.array_equal(out[keep], ref[keep])
    assert np.abs(out[100]).min() < 1e-300      # the root at the origin


def test_poly_roots_under_cuda_graph(cb, ea_golden):
    """the launcher only enqueues: capture once, replay on new coefficients"""
    c = torch.from_numpy(np.ascontiguousarray(ea_golden["rand5_coeffs"])).cuda()
    cb.poly_roots(c, itmax=2500)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        z = cb.poly_roots(c, itmax=2500)
        zc = cb.poly_roots(c, itmax=2500, compensated=True)
    c.mul_(1.5 - 0.25j).add_(0.01)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(z, cb.poly_roots(c, itmax=2500)) and torch.equal(zc, cb.poly_roots(c, itmax=2500, compensated=True))


def test_c2_full_size_properties(cb):
    """config 2 at full size -- 10^6 degree-10 triple-lens polynomials through the C ABI -- checked by
    size-independent properties (the oracle would need minutes): backward error of every root,
    Vieta's sum and product, plain vs compensated, and the reference on every 1000th polynomial."""
    from caustics_b200.point_source import _poly_coeffs_torch
    n = 1_000_000
    w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
    c = _poly_coeffs_torch(w, 3, **C2_PARAMS)                      # (n, 11) high -> low
    z = cb.poly_roots(c, itmax=2500)                                # (n, 10)
    zc = cb.poly_roots(c, itmax=2500, compensated=True)
    assert torch.isfinite(torch.view_as_real(z)).all()
    # backward error: |p(z)| <= 1e-13 * sum |p_k| |z|^k  (the solver's stopping test is 2^-53 * ~40 * that sum)
    val = torch.zeros_like(z)
    bound = torch.zeros(z.shape, dtype=torch.float64, device="cuda")
    az = z.abs()
    for k in range(11):
        val = val * z + c[:, k:k + 1]
        bound = bound * az + c[:, k:k + 1].abs()
    assert (val.abs() <= 1e-13 * bound).all()
    # Vieta: sum of roots = -c1/c0, product = c10/c0 (degree even)
    s_want, p_want = -c[:, 1] / c[:, 0], c[:, 10] / c[:, 0]
    scale = z.abs().sum(1)
    assert ((zc.sum(1) - s_want).abs() <= 1e-12 * scale).all()
    assert ((zc.prod(1) - p_want).abs() <= 1e-10 * p_want.abs() + 1e-300).all()
    # plain vs compensated as unordered sets, through sorted symmetric functions of the differences
    sub = slice(0, None, 1000)
    want = solver.solve(np.ascontiguousarray(c[sub].cpu().numpy()[:, ::-1]), compensated=True)
    assert set_distance(zc[sub].cpu().numpy(), want).max() < 1e-12
    # plain mode: the C2 polynomials are ill-conditioned (the roots next to the 2.8 % mass move by
    # ~1e-8 for a last-bit change of the coefficients); the reference's plain solver is no better
    assert set_distance(z[sub].cpu().numpy(), want).max() < 1e-6


def test_c2_plain_is_no_worse_than_the_reference(cb):
    """The benched mode (plain, C2) against the reference's OWN plain solve, polynomial by polynomial
    (VERDICT r1, weak point 1): the C2 polynomials are ill-conditioned, the reference's plain roots are
    only ~1e-8 from its compensated ones, so 1e-12 agreement is not defined in this mode -- what is defined
    is that the GPU's plain roots are no further from the truth (the compensated reference) than the
    reference's plain roots are (same error distribution quantile by quantile; a single polynomial's
    forward error is conditioning x a rounding-dependent factor, so pointwise only an order of magnitude is
    asserted), and that their backward error is no larger than the reference's."""
    from caustics_b200.point_source import _poly_coeffs_torch
    n = 1_000_000
    w = torch.from_numpy(np.linspace(-2, 2, n)[123::500] + 0.1j).cuda()        # 2000 polynomials of the C2 batch
    c = _poly_coeffs_torch(w, 3, **C2_PARAMS)
    z = cb.poly_roots(c, itmax=2500).cpu().numpy()
    cl = np.ascontiguousarray(c.cpu().numpy()[:, ::-1])
    ref_plain = solver.solve(cl, compensated=False)
    ref_comp = solver.solve(cl, compensated=True)
    d_gpu, d_ref = set_distance(z, ref_comp), set_distance(ref_plain, ref_comp)
    qs = (50, 90, 99, 100)
    q_gpu, q_ref = np.percentile(d_gpu, qs), np.percentile(d_ref, qs)
    assert (q_gpu <= 1.25 * q_ref + 1e-15).all(), (q_gpu, q_ref)
    # pointwise: an order of magnitude, with room for the odd polynomial on which the reference's plain solve is
    # accidentally good (1 of these 2000 on the B200), and never beyond the reference's own worst case
    bad = d_gpu > 10 * d_ref + 1e-10
    assert bad.sum() <= 2, (np.flatnonzero(bad), d_gpu[bad], d_ref[bad])
    assert d_gpu.max() <= 1.25 * d_ref.max(), (d_gpu.max(), d_ref.max())
    # (two independent draws of 'conditioning x rounding factor' are within 2x of each other for 94 % of these
    # polynomials on the B200; the quantile check above is the sharp statement)
    assert (d_gpu <= 2 * d_ref + 1e-12).mean() > 0.9

    def backward_error(roots):
        ch = cl[:, ::-1]
        val, bound = np.zeros_like(roots), np.zeros(roots.shape)
        for k in range(ch.shape[1]):
            val = val * roots + ch[:, k:k + 1]
            bound = bound * np.abs(roots) + np.abs(ch[:, k:k + 1])
        return (np.abs(val) / bound).max(axis=1)

    be_gpu, be_ref = backward_error(z), backward_error(ref_plain)
    assert (be_gpu <= np.maximum(be_ref, 40 * 2.0**-53) * (1 + 1e-9)).mean() > 0.99     # the stopping test's bound
    assert be_gpu.max() <= 2 * be_ref.max()
    # and the iteration path is the reference's: same root ORDER to the conditioning of the polynomial
    assert (np.abs(z - ref_plain).max(axis=1) <= 4 * d_ref + 1e-12).mean() > 0.95

