"""CPU tier: pins the oracle (oracle/) against the reference's own outputs -- the committed golden
vectors (generated from the reference by tests/golden/make_golden.py) and, where oracle/_ref is
built, the compiled reference solver itself."""
import numpy as np
import pytest

from conftest import set_distance
from oracle import lens, solver


def _polyval_low_high(c, z):
    out = np.zeros_like(z)
    for k in range(c.shape[-1] - 1, -1, -1):
        out = out * z + c[..., k:k + 1]
    return out


@pytest.mark.parametrize("name", ["fixture", "c1", "c2", "rand4", "rand5", "rand6", "rand10"])
@pytest.mark.parametrize("comp", [False, True])
def test_port_matches_reference_golden(ea_golden, name, comp):
    c = ea_golden[name + "_coeffs"]
    c = c.reshape(-1, c.shape[-1])[:, ::-1]
    want = ea_golden[name + ("_roots_comp" if comp else "_roots_plain")]
    got = solver.port_solve(c, itmax=2500, compensated=comp)
    # same algorithm, same order of roots; libm/thrust differences only
    if name == "c2" and not comp:
        # ill-conditioned roots of this triple lens: plain mode is only ~1e-8 accurate
        assert np.abs(got - want).max() < 1e-6
    else:
        assert np.abs(got - want).max() < 5e-13
    # against the compensated reference as an unordered set: the 1e-12 bar of BASELINE.md
    if comp:
        assert set_distance(got, ea_golden[name + "_roots_comp"]).max() < 1e-12


@pytest.mark.parametrize("comp", [False, True])
def test_reference_fixture_residual(ea_golden, comp):
    """tests/test_ehrlich_aberth_primitive.py:30-35 -- |p(root)| <= 1e-10 on the fixture"""
    c = ea_golden["fixture_coeffs"].reshape(-1, 6)[:, ::-1]
    z = solver.port_solve(c, itmax=2000, compensated=comp)
    assert np.abs(_polyval_low_high(c, z)).max() < 1e-10


@pytest.mark.skipif(not solver.ref_available(), reason="oracle/_ref not built")
def test_port_matches_compiled_reference():
    rng = np.random.default_rng(5)
    for deg in (3, 5, 8, 10):
        c = rng.standard_normal((2000, deg + 1)) + 1j * rng.standard_normal((2000, deg + 1))
        for comp in (False, True):
            assert np.abs(solver.ref_solve(c, compensated=comp) -
                          solver.port_solve(c, compensated=comp)).max() < 5e-13
    # custom_init: warm start from perturbed roots keeps root j the continuation of init j
    c = rng.standard_normal((500, 6)) + 1j * rng.standard_normal((500, 6))
    z0 = solver.ref_solve(c)
    init = z0 + 1e-4 * (rng.standard_normal(z0.shape) + 1j * rng.standard_normal(z0.shape))
    z1 = solver.ref_solve(c, custom_init=True, roots_init=init)
    z2 = solver.port_solve(c, custom_init=True, roots_init=init)
    assert np.abs(z1 - z0).max() < 1e-9 and np.abs(z2 - z1).max() < 5e-13


def test_coefficients_match_reference(ps_golden):
    w = ps_golden["w"]
    cb = lens.poly_coeffs(w, 2, a=0.45, e1=1 / 1.2)
    ct = lens.poly_coeffs(w, 3, a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)
    for got, want in ((cb, ps_golden["binary_coeffs"]), (ct, ps_golden["triple_coeffs"])):
        scale = np.abs(want).max(axis=1, keepdims=True)
        assert (np.abs(got - want) / scale).max() < 1e-13


def test_mag_point_source_matches_reference(ps_golden):
    w = ps_golden["w"]
    assert np.allclose(lens.mag_point_source(w, 2, s=0.9, q=0.2), ps_golden["mag_binary"], rtol=1e-10, atol=0)
    assert np.allclose(lens.mag_point_source(ps_golden["grid_w"], 2, s=0.9, q=0.2),
                       ps_golden["grid_mag_binary"], rtol=1e-10, atol=0)
    hp3 = dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0)
    assert np.allclose(lens.mag_point_source(w, 3, **hp3), ps_golden["mag_triple"], rtol=1e-9, atol=0)
    assert np.allclose(lens.mag_point_source(w, 3, roots_compensated=True, **hp3),
                       ps_golden["mag_triple_comp"], rtol=1e-9, atol=0)


def test_hexadecapole_and_gate_match_reference(ps_golden):
    a, e1 = 0.45, 1 / 1.2
    z, m, ws = ps_golden["images_z"], ps_golden["images_mask"], ps_golden["w"] + 0.3
    for rho in (1e-2, 1e-1):
        mu, dmu = lens.mag_hexadecapole(z, m, rho, nlenses=2, a=a, e1=e1)
        assert np.allclose(mu, ps_golden[f"hex_mu_{rho}"], rtol=1e-10)
        assert np.allclose(dmu, ps_golden[f"hex_dmu_{rho}"], rtol=1e-9, atol=1e-14)
        t = lens.caustics_proximity_test(ws, z, m, rho, ps_golden[f"hex_dmu_{rho}"], a=a, e1=e1)
        assert (t == ps_golden[f"gate_{rho}"]).all()
        assert (lens.planetary_caustic_test(ws, rho, a=a, e1=e1) == ps_golden[f"planet_{rho}"]).all()
    mu, dmu = lens.mag_hexadecapole(z, m, 0.05, u1=0.4, nlenses=2, a=a, e1=e1)
    assert np.allclose(mu, ps_golden["hex_mu_ld"], rtol=1e-10)
    assert np.allclose(dmu, ps_golden["hex_dmu_ld"], rtol=1e-9, atol=1e-14)


def test_jvp_rule_vs_finite_differences():
    """ehrlich_aberth_primitive.py:290-324 restated (oracle.lens.jvp_roots) vs central differences,
    tolerance of tests/test_ehrlich_aberth_primitive.py:59-64 (atol=rtol=1e-4)."""
    rng = np.random.default_rng(2)
    c = rng.standard_normal((50, 6)) + 1j * rng.standard_normal((50, 6))
    dc = rng.standard_normal((50, 6)) + 1j * rng.standard_normal((50, 6))
    z = solver.port_solve(c, compensated=True)
    h = 1e-6
    zp = solver.port_solve(c + h * dc, compensated=True, custom_init=True, roots_init=z)
    zm = solver.port_solve(c - h * dc, compensated=True, custom_init=True, roots_init=z)
    fd = (zp - zm) / (2 * h)
    assert np.allclose(lens.jvp_roots(c, z, dc), fd, rtol=1e-4, atol=1e-4)


def test_c2_update_count():
    """the algorithmic-work constant bench.py uses for the roofline: plain root
    updates per polynomial of the C2 workload, counted by the restatement of the reference algorithm"""
    w = np.linspace(-2, 2, 1000000)[::500] + 0.1j
    c = lens.poly_coeffs(w, 3, a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)
    _, sweeps, stats = solver.port_solve(np.ascontiguousarray(c[:, ::-1]), itmax=2500, return_stats=True)
    assert abs(stats[0] / len(w) - 95.2) < 1.5
    assert stats[2] == 0


@pytest.mark.parametrize("k", ["b", "t"])
def test_sequential_images_match_reference(seq_golden, k):
    """_images_point_source_sequential (point_source.py:1711-1759) vs the reference's own Python:
    same warm-start chain, hence the same ROW ORDER of images along the path"""
    from conftest import SEQ_PARAMS
    nl, p = SEQ_PARAMS[k]
    z, m = lens.images_point_source_sequential(seq_golden[f"{k}_w"], nl, **p)
    assert z.shape == seq_golden[f"{k}_z"].shape
    assert np.array_equal(m, seq_golden[f"{k}_mask"])
    # ordered comparison.  The reference builds the coefficients from its expanded monomials, the
    # oracle from the product form: roots differ by (condition number) x 1e-16, which reaches 4e-10 for
    # the images next to the 2.8 % mass of the triple lens (plain solver on both sides)
    assert np.abs(z - seq_golden[f"{k}_z"]).max() < 1e-9
    if k == "b":
        assert np.abs(z - seq_golden[f"{k}_z"])[seq_golden[f"{k}_mask"]].max() < 1e-12


def test_jax_prng_known_answers():
    """oracle/jaxprng.py (the reference's jitter stream, extended_source.py:76-85,146) against the published
    known answers: the Random123 test vectors of threefry2x32 (Salmon et al., SC'11 -- also JAX's own
    tests/random_test.py::testThreefry2x32), JAX's documented `random.split(PRNGKey(0))` and
    `random.uniform(PRNGKey(0))`"""
    from oracle import jaxprng as J
    u = lambda *v: np.array(v, dtype=np.uint32)
    for key, ctr, want in (((0, 0), (0, 0), (0x6b200159, 0x99ba4efe)),
                           ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
                           ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))):
        y0, y1 = J.threefry2x32(key[0], key[1], u(ctr[0]), u(ctr[1]))
        assert (int(y0[0]), int(y1[0])) == want
    assert J.prng_key(0).tolist() == [0, 0] and J.prng_key(2**32 + 5).tolist() == [1, 5]
    assert J.split(J.prng_key(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert abs(float(J.uniform(J.prng_key(0), (), np.float32)) - 0.41845703) < 1e-8
    # float64 draws: in range, reproducible, a different stream per key, C-order element i is a pure function
    # of (key, i, N)
    t = J.limb_jitters(5, 10)
    assert t.shape == (5, 10) and (np.abs(t.real) <= 1e-6).all() and (np.abs(t.imag) <= 1e-6).all()
    assert np.array_equal(t, J.limb_jitters(5, 10)) and not np.array_equal(t.real, t.imag)
    assert len(np.unique(t)) == 50 and 2e-7 < np.abs(t.real).mean() < 8e-7
    d = J.duplicate_jitters(10, 200)
    assert d.shape == (10, 200) and (np.abs(d) <= 1e-9).all() and abs(d.mean()) < 1e-10
