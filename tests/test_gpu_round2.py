"""GPU tier, round-2 additions: the gather without a collective (PeerGather / HostGather, the map entry with a
host result buffer), gated calls whose workspace is sized by the gate's survivors, and -- on a box with two
GPUs -- the one-process-per-GPU path itself (NCCL for the rendezvous, kernel stores over NVLink for the data)."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
HP2 = dict(s=0.9, q=0.2)


@pytest.fixture(scope="module")
def cb():
    import caustics_b200
    return caustics_b200


def test_map_into_peer_buffer_and_host_buffer(cb):
    """the map written straight into a PeerGather buffer (single process: a plain local allocation) and into
    a host array through caustics_mag_point_source_grid_host equals the ordinary map bit for bit, cold and
    walked (host chunks are cut on 32-row walk boundaries)"""
    from caustics_b200 import _lib
    from caustics_b200.sharding import PeerGather
    nx, ny = 3000, 1100
    x0, y0, dx = -1.2, -0.5, 1e-3
    # maps this small would get shorter walks the smaller the row block: pin the walk length, as a map large
    # enough to fill the GPU has it (sharding.row_block(align=32) documents the condition)
    _lib.lib().caustics_set_tuning(b"grid_run", 32)
    for walk in (False, True):
        want = cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, walk=walk, **HP2)
        pg = PeerGather(nx * ny * 8)
        # two row blocks, like two ranks would write them (32-row aligned so the walked map is the same walks)
        for lo, hi in ((0, 544), (544, ny)):
            assert cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, rows=(lo, hi), walk=walk, out=pg.ptr(lo * nx * 8), **HP2) is None
        pg.finish()
        assert torch.equal(pg.tensor(torch.float64, (ny, nx)), want)
        pg.close()
        host = np.empty((ny, nx))
        assert cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, walk=walk, out=host, **HP2) is host
        assert np.array_equal(host, want.cpu().numpy())
        part = np.full((100, nx), -1.0)
        cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, rows=(64, 164), walk=walk, out=part, **HP2)
        assert np.array_equal(part, want[64:164].cpu().numpy())
    _lib.lib().caustics_set_tuning(b"grid_run", -1)
    with pytest.raises(ValueError):
        cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, out=np.empty((ny, nx), dtype=np.float32), **HP2)


def test_gated_workspace_sized_by_survivors(cb):
    """caustics_mag with a workspace for a fraction of the points integrates the gate's survivors in windows
    (no host read-back) and gives the same bits as the full-size workspace; the two-call form
    caustics_mag_gate -> caustics_mag_extended_source_list likewise"""
    from caustics_b200 import _lib
    L = _lib.lib()
    n = 20_000
    w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
    p, x_cm = cb.lens_params(2, **HP2)
    lens = cb.point_source._c_lens(2, x_cm, **p)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for ld in (0, 1):
        for cap in (n, 700, 97):
            nb = L.caustics_mag_workspace_bytes(n, cap, 2, 200, ld, 100)
            assert nb > 0
            ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
            mag = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
            used = torch.empty(n, dtype=torch.uint8, device="cuda")
            L.caustics_set_tuning(b"ext_variants", 31)        # same phase variants whatever the window size
            try:
                _lib.check(L.caustics_mag(w.data_ptr(), mag.data_ptr(), used.data_ptr(), n, 1e-2, lens, 0.2, 200, ld, 0.7,
                                          100, 2500, 0, ws.data_ptr(), nb, st))
            finally:
                L.caustics_set_tuning(b"ext_variants", -1)
            torch.cuda.synchronize()
            res[(ld, cap)] = (mag, used)
        nfull = int((res[(ld, n)][1] == 0).sum().item())
        assert 1000 < nfull < 1400                  # several windows of 700, many of 97
        for cap in (700, 97):
            assert torch.equal(res[(ld, cap)][0], res[(ld, n)][0]) and torch.equal(res[(ld, cap)][1], res[(ld, n)][1])
        # the public call (gate -> count -> exact workspace -> list integration)
        m, t = cb.mag(w, 1e-2, nlenses=2, npts_limb=200, limb_darkening=bool(ld), u1=0.7, npts_ld=100, return_test=True, **HP2)
        assert torch.equal(t, res[(ld, n)][1].bool())
        assert torch.allclose(m, res[(ld, n)][0], rtol=1e-9, atol=0)
    # too small for one source + the list: a clean argument error, nothing launched
    ws = torch.empty(1000, dtype=torch.uint8, device="cuda")
    mag = torch.empty(n, dtype=torch.float64, device="cuda")
    assert L.caustics_mag(w.data_ptr(), mag.data_ptr(), None, n, 1e-2, lens, 0.2, 200, 0, 0.0, 100, 2500, 0, ws.data_ptr(), 1000, st) == 1


def test_million_point_gated_light_curve_fits_4gb(cb):
    """a 10^6-point gated binary light curve (limb-darkened) needs < 4 GB of device memory: the workspace is
    sized by the points that fail the gate, not by the length of the light curve"""
    n = 1_000_000
    w = torch.from_numpy(np.linspace(-2, 2, n) + 0.1j).cuda()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    m, t = cb.mag(w, 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100, return_test=True, **HP2)
    torch.cuda.synchronize()
    assert torch.cuda.max_memory_allocated() - base < 4 << 30
    assert bool(torch.isfinite(m).all()) and 0.03 < float((~t).double().mean()) < 0.09
    # every 100th point is a point of the 10^4-point light curve C3 (same linspace ends): same values
    sub = cb.mag(w[::100].contiguous(), 1e-2, nlenses=2, npts_limb=200, limb_darkening=True, u1=0.7, npts_ld=100, **HP2)
    assert torch.allclose(m[::100], sub, rtol=1e-9, atol=0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _two_gpu_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import caustics_b200 as cb
    from caustics_b200 import _lib, sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = True
    # C5-like map: every rank's kernel stores its row block into rank 0's buffer over NVLink
    nx, ny = 2000, 1024
    x0, y0, dx = -1.2, -0.5, 1.5e-3
    _lib.lib().caustics_set_tuning(b"grid_run", 32)       # see test_map_into_peer_buffer_and_host_buffer
    pg = sharding.PeerGather(nx * ny * 8, dst=0)
    lo, hi = sharding.row_block(ny, world, rank, align=32)
    for walk in (False, True):
        cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, rows=(lo, hi), walk=walk, out=pg.ptr(lo * nx * 8), **HP2)
        pg.finish()
        if rank == 0:
            want = cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, walk=walk, **HP2)
            ok = ok and bool(torch.equal(pg.tensor(torch.float64, (ny, nx)), want))
        dist.barrier()
    pg.close()
    # the host-side assembly: both ranks' D2H copies into one shared, page-locked map
    hg = sharding.HostGather(nx * ny * 8, dst=0)
    v = hg.view(np.float64, (ny, nx))
    cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, rows=(lo, hi), walk=False, out=v[lo:hi], **HP2)
    hg.finish()
    want = cb.mag_point_source_map(x0, y0, dx, dx, nx, ny, walk=False, **HP2).cpu().numpy()
    ok = ok and bool(np.array_equal(v, want))
    del v
    hg.close()
    # extended sources sharded, results into rank 0's buffer
    n = 600
    w_all = np.linspace(-0.4, 0.4, n) + 0.05j
    L = _lib.lib()
    p, x_cm = cb.lens_params(2, **HP2)
    lens = cb.point_source._c_lens(2, x_cm, **p)
    slo, shi = sharding.shard_bounds(n, world, rank)
    w = torch.from_numpy(w_all[slo:shi]).cuda()
    nb = L.caustics_ext_workspace_bytes(shi - slo, 2, 200, 0, 100)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    pg = sharding.PeerGather(n * 8, dst=0)
    _lib.check(L.caustics_mag_extended_source(w.data_ptr(), pg.ptr(slo * 8), shi - slo, 1e-2, lens, 200, 0, 0.0, 100, 2500, 0,
                                              ws.data_ptr(), nb, torch.cuda.current_stream().cuda_stream))
    pg.finish()
    if rank == 0:
        want = cb.mag_extended_source(torch.from_numpy(w_all).cuda(), 1e-2, nlenses=2, npts_limb=200, **HP2)
        ok = ok and bool(torch.allclose(pg.tensor(torch.float64, (n,)), want, rtol=1e-9, atol=0))
    pg.close()
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_two_process_gather_over_nvlink():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True), (1, True)]


def test_fused_tangent_kernel_vs_python_rule(cb):
    """the tangent the kernels emit alongside the forward pass (caustics_mag_extended_source_grad,
    implicit-function step at every contour vertex + tangent of the trapezoid sum, on the device) equals the
    Python rule (`_mag_from_contours`: the same step written in torch on the exported contours) for every
    parameter -- binary and triple lens, caustic-crossing sources, sources inside and outside caustics"""
    from caustics_b200 import extended_source as es
    g = np.load(os.path.join(ROOT, "tests", "golden", "ext_golden.npz"))
    cases = [(2, dict(s=0.9, q=0.2), g["b_w_0.01"][:24]),
             (3, dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0), g["t_w_0.01"][:12]),
             (1, {}, np.array([0.003 + 0.001j, 0.02 - 0.01j, 0.5 + 0.2j]))]
    # The two tangents are the same formula evaluated at vertices that differ by the solver's residual (the
    # specification re-polishes each vertex); near a fold dz/dt grows like 1 / det J, so the agreement is the
    # roots' accuracy times that factor.  With compensated roots (1e-16) the two agree to 1e-7 for every lens;
    # with plain roots the triple lens -- whose plain-mode roots are only ~1e-9 accurate (DESIGN section 2,
    # conditioning) -- agrees to 4e-5 at one fold-crossing source and 2e-8 elsewhere (scripts/tangent_probe.py).
    GRAD_TOL = {(1, False): 1e-8, (2, False): 1e-8, (3, False): 2e-4, (1, True): 1e-7, (2, True): 1e-7, (3, True): 1e-7}
    for nl, hp, w_np in cases:
      for comp in (False, True):
        def run(fn):
            w = torch.from_numpy(w_np).cuda().requires_grad_(True)
            rho = torch.tensor(1e-2, dtype=torch.float64, device="cuda", requires_grad=True)
            t = {k: torch.tensor(v, dtype=torch.float64, device="cuda", requires_grad=True) for k, v in hp.items()}
            wt = torch.from_numpy(np.random.default_rng(1).uniform(0.5, 1.5, len(w_np))).cuda()
            m = fn(w, rho, t)
            (m * wt).sum().backward()
            return m.detach(), w.grad, rho.grad, {k: v.grad for k, v in t.items()}

        kern = run(lambda w, rho, t: es._mag_uniform_kernel_grad(w, rho, nl, 200, 2500, comp, t))
        spec = run(lambda w, rho, t: es._mag_from_contours(es._get_contours(w, rho, nl, 200, 2500, comp, t),
                                                           w.reshape(-1), rho, nl, t))
        # the specification re-polishes every vertex by one Newton step (z0 - J^-1 F(z0)): F(z0) is the solver's
        # residual, not 0, and near a caustic J^-1 is large, so the two magnifications differ at the level of the
        # roots' own accuracy -- the forward value of the tangent kernel IS the plain kernel's, asserted below
        dev = ((kern[0] - spec[0]).abs() / spec[0].abs()).max().item()
        assert dev < 1e-8, (nl, comp, dev)
        plain = es._run(torch.from_numpy(w_np).cuda(), 1e-2, nl, 200, False, 0.0, 100, 2500, comp, False, 0.0, hp)
        assert torch.allclose(kern[0], plain.reshape(-1), rtol=1e-13, atol=0), (nl, comp)
        tol = GRAD_TOL[(nl, comp)]
        gdev = ((kern[1] - spec[1]).abs() / (spec[1].abs() + 1e-9 * spec[1].abs().max())).max().item()
        assert gdev < tol, (nl, comp, "d/dw", gdev)
        rdev = abs(kern[2].item() - spec[2].item()) / abs(spec[2].item())
        assert rdev < tol, (nl, comp, "d/drho", rdev)
        for k in hp:
            pdev = abs(kern[3][k].item() - spec[3][k].item()) / max(abs(spec[3][k].item()), 1e-3)
            assert pdev < tol, (nl, comp, k, pdev)
    # the public entry takes the kernel path and chunked calls agree with one call
    w = torch.from_numpy(g["b_w_0.01"][:24]).cuda().requires_grad_(True)
    m = cb.mag_extended_source(w, 1e-2, nlenses=2, npts_limb=200, s=0.9, q=0.2)
    m.sum().backward()
    old = es._MAX_WS_BYTES
    try:
        es._MAX_WS_BYTES = 400_000          # ~10 sources per chunk
        w2 = torch.from_numpy(g["b_w_0.01"][:24]).cuda().requires_grad_(True)
        m2 = cb.mag_extended_source(w2, 1e-2, nlenses=2, npts_limb=200, s=0.9, q=0.2)
        m2.sum().backward()
    finally:
        es._MAX_WS_BYTES = old
    assert torch.allclose(m, m2, rtol=1e-9) and torch.allclose(w.grad, w2.grad, rtol=1e-6, atol=1e-9)
