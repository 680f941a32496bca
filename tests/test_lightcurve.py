"""Trajectory and flux-marginalised likelihood: oracle vs the reference's own code
(tests/golden/lc_golden.npz, made by tests/golden/make_golden_lc.py) on the CPU tier, kernels vs both
on the GPU tier."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import lightcurve as olc


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "lc_golden.npz"))


def _tables(g):
    return tuple(g[k] for k in ("traj_t_jpl", "traj_s_e", "traj_s_n", "traj_s_e_dot", "traj_s_n_dot"))


def _tp(g):
    return dict(zip(("t0", "tE", "u0", "piEE", "piEN"), g["traj_params"]))


def test_oracle_trajectory_matches_reference(g):
    tp = _tp(g)
    w = olc.trajectory(g["traj_t"], _tables(g), "cartesian", **tp)
    assert np.abs(w - g["traj_w_cartesian"]).max() < 1e-14
    w = olc.trajectory(g["traj_t"], _tables(g), "polar", t0=tp["t0"], tE=tp["tE"], u0=tp["u0"], psi=0.7, piE=0.3)
    assert np.abs(w - g["traj_w_polar"]).max() < 1e-14


def test_oracle_likelihood_matches_reference(g):
    betas, ll = olc.marginalized_log_likelihood([g["ll_A0"], g["ll_A1"]], [g["ll_f0"], g["ll_f1"]],
                                                [g["ll_c0"], g["ll_c1"]])
    assert abs(ll - g["ll_total"]) < 1e-9 * abs(g["ll_total"])
    for i in range(2):
        assert np.allclose(betas[i], g[f"ll_beta{i}"], rtol=1e-11)


@pytest.mark.gpu
def test_trajectory_kernel(g):
    import torch
    import caustics_b200 as cb
    tp = _tp(g)
    tab = dict(zip(("t_jpl", "s_e", "s_n", "s_e_dot", "s_n_dot"), _tables(g)))
    tr = cb.AnnualParallaxTrajectory(g["traj_t"], **tab)
    w = tr.compute(g["traj_t"], **tp)
    # |w| ~ 10 (|t - t0| / tE): 1e-13 absolute is a few ulp (sincos and fma contraction differ from NumPy)
    assert np.abs(w - g["traj_w_cartesian"]).max() < 1e-13
    w = tr.compute(torch.from_numpy(g["traj_t"]).cuda(), "polar", t0=tp["t0"], tE=tp["tE"], u0=tp["u0"], psi=0.7, piE=0.3)
    assert w.is_cuda and np.abs(w.cpu().numpy() - g["traj_w_polar"]).max() < 1e-13
    # rectilinear (no tables) and times outside the table (interp clamps like numpy)
    t2 = np.concatenate([[g["traj_t_jpl"][0] - 30.0], g["traj_t"][:5], [g["traj_t_jpl"][-1] + 12.5]])
    assert np.abs(tr.compute(t2, **tp) - olc.trajectory(t2, _tables(g), **tp)).max() < 1e-13
    w0 = cb.AnnualParallaxTrajectory().compute(g["traj_t"], **tp)
    assert np.abs(w0 - olc.trajectory(g["traj_t"], None, **tp)).max() < 1e-13
    with pytest.raises(ValueError):
        tr.compute(g["traj_t"], "spherical", **tp)


@pytest.mark.gpu
def test_likelihood_kernel(g):
    import torch
    import caustics_b200 as cb
    As, fs, cs = ([g[f"ll_{k}{i}"] for i in range(2)] for k in "Afc")
    betas, ll = cb.marginalized_log_likelihood(As, fs, cs)
    assert abs(ll - g["ll_total"]) < 1e-10 * abs(g["ll_total"])
    for i in range(2):
        assert np.allclose(betas[i], g[f"ll_beta{i}"], rtol=1e-11)
    # deterministic: bit-identical on repeat; torch inputs stay torch
    b2, ll2 = cb.marginalized_log_likelihood([torch.from_numpy(a).cuda() for a in As], fs, cs)
    assert ll2 == ll and isinstance(b2[0], torch.Tensor)
    # ragged / tiny / large
    rng = np.random.default_rng(3)
    def direct(A, f, c):   # the same normal equations without the oracle's n x n diag(C_inv)
        M = np.stack([A, np.ones_like(A)]).T
        S = np.linalg.inv(M.T @ (c[:, None] * M))
        b = S @ (M.T @ (c * f))
        return b, -0.5 * np.sum((f - M @ b) ** 2 * c) + 0.5 * np.log(np.linalg.det(2 * np.pi * S))
    for n in (2, 3, 1023, 1025, 200001):
        A = 1 + rng.uniform(0, 5, n); c = rng.uniform(1, 4, n); f = 2 * A + 1 + 0.1 * rng.standard_normal(n)
        (b,), l = cb.marginalized_log_likelihood([A], [f], [c])
        bo, lo = direct(A, f, c)
        if n <= 1025:      # the oracle follows the reference and materialises diag(C_inv): small n only
            (bo2,), lo2 = olc.marginalized_log_likelihood([A], [f], [c])
            assert np.allclose(bo, bo2, rtol=1e-9) and abs(lo - lo2) < 1e-9 * max(1.0, abs(lo2))
        assert np.allclose(b, bo, rtol=1e-9) and abs(l - lo) < 1e-9 * max(1.0, abs(lo))
    with pytest.raises(NotImplementedError):
        cb.marginalized_log_likelihood(As, fs, cs, dense_covariance=True)
    with pytest.raises(Exception):
        cb.marginalized_log_likelihood([As[0][:1]], [fs[0][:1]], [cs[0][:1]])


@pytest.mark.gpu
def test_light_curve_log_likelihood_closure(g):
    """trajectory -> mag (gated, limb-darkened) -> likelihood in one enqueue == the same composition on
    the oracle side"""
    import caustics_b200 as cb
    from oracle import extended
    rng = np.random.default_rng(11)
    t = np.linspace(-25.0, 25.0, 400)
    tp = dict(t0=0.7, tE=20.0, u0=0.1, piEE=0.0, piEN=0.0)
    hp = dict(s=0.9, q=0.2)
    w = olc.trajectory(t, None, **tp)
    A = np.asarray(extended.mag(w[::8], 1e-2, 2, 100, True, 0.3, 50, **hp))
    Afull = cb.mag(w, 1e-2, nlenses=2, npts_limb=100, limb_darkening=True, u1=0.3, npts_ld=50, **hp)
    assert np.allclose(Afull[::8], A, rtol=1e-4)
    sig = 0.01 * np.ones_like(t)
    f = 2.5 * Afull + 0.4 + sig * rng.standard_normal(len(t))
    beta, ll = cb.light_curve_log_likelihood(t, f, 1 / sig**2, cb.AnnualParallaxTrajectory(), 1e-2, tp, hp,
                                             npts_limb=100, limb_darkening=True, u1=0.3, npts_ld=50)
    (bo,), lo = olc.marginalized_log_likelihood([Afull], [f], [1 / sig**2])
    assert np.allclose(beta, bo, rtol=1e-10) and abs(ll - lo) < 1e-9 * abs(lo)
    assert abs(beta[0] - 2.5) < 0.01 and abs(beta[1] - 0.4) < 0.01


@pytest.mark.gpu
def test_closure_under_cuda_graph(g):
    """trajectory -> mag -> likelihood captured once, replayed with new times/fluxes: the per-step cost of
    an HMC likelihood is one graph launch and one 24-byte read"""
    import torch
    import caustics_b200 as cb
    from caustics_b200.lightcurve import _loglike_device
    hp = dict(s=0.9, q=0.2)
    tp = dict(t0=0.7, tE=20.0, u0=0.1, piEE=0.0, piEN=0.0)
    traj = cb.AnnualParallaxTrajectory()
    t = torch.linspace(-25.0, 25.0, 300, dtype=torch.float64, device="cuda")
    f = torch.empty_like(t)
    cinv = torch.full_like(t, 1e4)
    kw = dict(nlenses=2, npts_limb=100, limb_darkening=True, u1=0.3, npts_ld=50, **hp)

    def eager(tt, ff):
        A = cb.mag(traj.compute(tt, **tp), 1e-2, **kw)
        return A, _loglike_device(A, ff, cinv)

    A0, _ = eager(t, f)
    f.copy_(2.5 * A0 + 0.4)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        A, out = eager(t, f)
    t.add_(0.37)                       # new times -> new trajectory, new magnifications
    f.mul_(1.01)
    graph.replay()
    torch.cuda.synchronize()
    A_want, out_want = eager(t, f)
    assert torch.equal(A, A_want) and torch.equal(out, out_want)
