"""Host-side mirror of the reference's point-source layer
(/root/reference/src/caustics/point_source.py): same function names, arguments and return layouts.

  mag_point_source(w, nlenses, roots_itmax, roots_compensated, s=, q=[, q3=, r3=, psi=])   :1762-1830
  _images_point_source(w, nlenses, ..., custom_init, z_init, a=, e1=[, e2=, r3=])           :1655-1709
  lens_eq / lens_eq_det_jac                                                                  :1536-1580

The work is done by kernel 2 (coefficients + Ehrlich-Aberth + lens-equation filter + Jacobian fused,
csrc/kernels.cu).  When an argument requires grad (torch), the same quantities are composed from
differentiable pieces instead -- coefficients in torch, `poly_roots` with its implicit-function
backward, det J in torch -- which is exactly the structure the reference differentiates through.
"""
import math

import numpy as np
import torch

from . import _lib
from .primitive import poly_roots

__all__ = ["mag_point_source", "mag_point_source_map", "lens_eq", "lens_eq_det_jac", "lens_params",
           "critical_and_caustic_curves"]


def lens_params(nlenses, **params):
    """High-level (s, q, q3, r3, psi) -> low-level `_params` and the centre-of-mass shift,
    point_source.py:1796-1819 (triple-lens e1/e2 exactly as the reference defines them)."""
    if nlenses == 1:
        return {}, 0.0
    s, q = params["s"], params["q"]
    a = 0.5 * s
    x_cm = a * (1 - q) / (1 + q)
    if nlenses == 2:
        return {"a": a, "e1": 1 / (1 + q)}, x_cm
    if nlenses == 3:
        q3, r3, psi = params["q3"], params["r3"], params["psi"]
        e1 = q / (1 + q + q3)
        if isinstance(psi, torch.Tensor) or isinstance(r3, torch.Tensor):
            r3c = r3 * torch.exp(1j * torch.as_tensor(psi))
        else:
            r3c = r3 * complex(math.cos(psi), math.sin(psi))
        return {"a": a, "e1": e1, "e2": q * e1, "r3": r3c}, x_cm
    raise ValueError("`nlenses` has to be set to be <= 3.")


def _c_lens(nlenses, x_cm=0.0, **p):
    f = lambda v: float(v.detach().cpu()) if isinstance(v, torch.Tensor) else float(v)
    L = _lib.Lens()
    L.nlenses = nlenses
    L.x_cm = f(x_cm)
    if nlenses >= 2:
        L.a, L.e1 = f(p["a"]), f(p["e1"])
    if nlenses == 3:
        r3 = p["r3"]
        r3 = complex(r3.detach().cpu()) if isinstance(r3, torch.Tensor) else complex(r3)
        L.e2, L.r3_re, L.r3_im = f(p["e2"]), r3.real, r3.imag
    return L


def _lenses(nlenses, **p):
    if nlenses == 1:
        return [0.0], [1.0]
    a, e1 = p["a"], p["e1"]
    if nlenses == 2:
        return [a, -a], [e1, 1.0 - e1]
    return [a, -a, p["r3"]], [e1, p["e2"], 1.0 - e1 - p["e2"]]


def _xp(z):
    return torch if isinstance(z, torch.Tensor) else np


def lens_eq(z, nlenses=2, **params):
    xp = _xp(z)
    zbar = xp.conj(z)
    r, eps = _lenses(nlenses, **params)
    out = z
    for rj, ej in zip(r, eps):
        rjb = xp.conj(rj) if isinstance(rj, torch.Tensor) else np.conj(rj)
        out = out - ej / (zbar - rjb)
    return out


def lens_eq_det_jac(z, nlenses=2, **params):
    xp = _xp(z)
    zbar = xp.conj(z)
    r, eps = _lenses(nlenses, **params)
    acc = 0.0
    for rj, ej in zip(r, eps):
        rjb = xp.conj(rj) if isinstance(rj, torch.Tensor) else np.conj(rj)
        acc = acc + ej / (zbar - rjb) ** 2
    return 1.0 - xp.abs(acc) ** 2


# ---- differentiable coefficient builder (torch), product form of SURVEY App. A.4 ---------------
def _tmul(a, b):
    na, nb = a.shape[-1], b.shape[-1]
    out = [0] * (na + nb - 1)
    for i in range(na):
        for j in range(nb):
            out[i + j] = out[i + j] + a[..., i] * b[..., j]
    return torch.stack(torch.broadcast_tensors(*out), dim=-1)


def _tadd(a, b):
    n = max(a.shape[-1], b.shape[-1])
    pad = lambda x: torch.nn.functional.pad(x, (0, n - x.shape[-1]))
    return pad(a) + pad(b)


def _poly_coeffs_torch(w, nlenses, **p):
    """HIGH -> LOW coefficients like the reference's `_poly_coeffs_binary/_triple`, differentiable."""
    dev = w.device
    C = lambda v: torch.as_tensor(v, dtype=torch.complex128, device=dev)
    r, eps = _lenses(nlenses, **p)
    r, eps = [C(x) for x in r], [C(x) for x in eps]
    n = len(r)
    one = lambda root: torch.stack([-root, C(1.0)])
    H = C(1.0).reshape(1)
    for ri in r:
        H = _tmul(H, one(ri))
    G = C(0.0).reshape(1)
    for j in range(n):
        t = eps[j].reshape(1)
        for i in range(n):
            if i != j:
                t = _tmul(t, one(r[i]))
        G = _tadd(G, t)
    wbar = torch.conj(w)[..., None]
    A = [_tadd(G.expand(w.shape + G.shape), -(torch.conj(r[j]) - wbar) * H) for j in range(n)]
    first = torch.stack([-w, torch.ones_like(w)], dim=-1)
    for j in range(n):
        first = _tmul(first, A[j])
    second = None
    for j in range(n):
        t = eps[j] * torch.ones(w.shape + (1,), dtype=torch.complex128, device=dev)
        for i in range(n):
            if i != j:
                t = _tmul(t, A[i])
        second = t if second is None else _tadd(second, t)
    second = _tmul(second, H.expand(w.shape + H.shape))
    return torch.flip(_tadd(first, -second), dims=[-1])


def _needs_grad(w, p):
    vals = [w] + list(p.values())
    return any(isinstance(v, torch.Tensor) and v.requires_grad for v in vals)


def _images_point_source(w, nlenses=2, roots_itmax=2500, roots_compensated=False, custom_init=False,
                         z_init=None, flags=0, **params):
    """Images z (root axis FIRST) and the real-image mask for source positions `w` (any shape).
    `params` are the low-level a, e1[, e2, r3]; `z_init` has the root axis LAST (reference layout)."""
    xp = _xp(w)
    if nlenses == 1:
        w_abs_sq = w.real**2 + w.imag**2
        sq = xp.sqrt(1 + 4 / w_abs_sq)
        z = xp.stack([0.5 * w * (1.0 + sq), 0.5 * w * (1.0 - sq)])
        return z, (torch.ones_like(z, dtype=torch.bool) if xp is torch else np.ones(z.shape, dtype=bool))
    if nlenses not in (2, 3):
        raise ValueError("`nlenses` has to be set to be <= 3.")
    deg = nlenses**2 + 1
    if isinstance(w, torch.Tensor) and _needs_grad(w, params):
        coeffs = _poly_coeffs_torch(w, nlenses, **params)
        z = poly_roots(coeffs, itmax=roots_itmax, compensated=roots_compensated,
                       custom_init=custom_init, roots_init=z_init, flags=flags)
        z = torch.movedim(z, -1, 0)
        mask = torch.abs(lens_eq(z, nlenses, **params) - w) < 1e-6
        return z, mask
    L = _lib.lib()
    lens = _c_lens(nlenses, 0.0, **params)
    shape = tuple(w.shape)
    if isinstance(w, torch.Tensor) and w.is_cuda:
        wf = w.to(torch.complex128).resolve_conj().resolve_neg().contiguous().reshape(-1)
        n = wf.numel()
        z = torch.empty((deg, n), dtype=torch.complex128, device=w.device)
        mask = torch.empty((deg, n), dtype=torch.uint8, device=w.device)
        zi = None
        if custom_init:
            zi = z_init.to(device=w.device, dtype=torch.complex128).resolve_conj().resolve_neg().contiguous().reshape(n, deg)
        with torch.cuda.device(w.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.caustics_images_point_source(
                wf.data_ptr(), zi.data_ptr() if custom_init else None, z.data_ptr(), mask.data_ptr(),
                n, lens, int(roots_itmax), int(bool(roots_compensated)), int(bool(custom_init)),
                int(flags), st))
        return z.reshape((deg,) + shape), mask.bool().reshape((deg,) + shape)
    # host arrays: stage through torch on the current device
    _lib.require_cuda()
    is_t = isinstance(w, torch.Tensor)
    wd = torch.as_tensor(np.asarray(w, dtype=np.complex128) if not is_t else w).to("cuda")
    zi = None
    if custom_init:
        zi = torch.as_tensor(np.asarray(z_init) if not isinstance(z_init, torch.Tensor) else z_init).to("cuda")
    z, mask = _images_point_source(wd, nlenses, roots_itmax, roots_compensated, custom_init, zi,
                                   flags, **params)
    z, mask = z.cpu(), mask.cpu()
    return (z, mask) if is_t else (z.numpy(), mask.numpy())


def _images_point_source_sequential(w, nlenses=2, roots_itmax=2500, roots_compensated=False, **params):
    """point_source.py:1711-1759: images along a 1-D path `w` (n,), every position warm-started from the
    previous one's images, so row j follows one image.  Returns z (deg, n), mask (deg, n).  Leading
    batch axes are allowed: w (..., n) -> z (..., deg, n); each path is one thread of one launch."""
    if nlenses not in (2, 3):
        raise ValueError("`nlenses` has to be 2 or 3.")
    _lib.require_cuda()
    is_t = isinstance(w, torch.Tensor)
    wd = w if is_t else torch.as_tensor(np.asarray(w, dtype=np.complex128))
    on_dev = wd.is_cuda
    wd = wd.to(device="cuda" if not on_dev else wd.device, dtype=torch.complex128).resolve_conj().resolve_neg().contiguous()
    if wd.dim() < 1:
        raise ValueError("`w` has to be at least one-dimensional (a path of source positions)")
    n = wd.shape[-1]
    npaths = wd.numel() // max(n, 1)
    deg = nlenses**2 + 1
    z = torch.empty(tuple(wd.shape[:-1]) + (deg, n), dtype=torch.complex128, device=wd.device)
    mask = torch.empty(z.shape, dtype=torch.uint8, device=wd.device)
    lens = _c_lens(nlenses, 0.0, **params)
    with torch.cuda.device(wd.device):
        _lib.check(_lib.lib().caustics_images_point_source_sequential(
            wd.data_ptr(), z.data_ptr(), mask.data_ptr(), npaths, n, lens, int(roots_itmax),
            int(bool(roots_compensated)), torch.cuda.current_stream().cuda_stream))
    mask = mask.bool()
    if not on_dev:
        z, mask = z.cpu(), mask.cpu()
    return (z, mask) if is_t else (z.numpy(), mask.numpy())


def mag_point_source(w, nlenses=2, roots_itmax=2500, roots_compensated=False, flags=0, **params):
    """Point-source magnification at source positions `w` (complex128, any shape); high-level
    parameters s, q[, q3, r3, psi] as in the reference (point_source.py:1762-1830)."""
    if not isinstance(w, (torch.Tensor, np.ndarray)):
        w = np.asarray(w, dtype=np.complex128)      # Python scalars / lists
    xp = _xp(w)
    if nlenses == 1:
        z, mask = _images_point_source(w, nlenses=1)
        det = lens_eq_det_jac(z, nlenses=1)
        return ((1.0 / xp.abs(det)) * mask).sum(0).reshape(w.shape)
    p, x_cm = lens_params(nlenses, **params)
    if isinstance(w, torch.Tensor) and _needs_grad(w, {**p, "x_cm": x_cm}):
        ws = w + x_cm
        z, mask = _images_point_source(ws, nlenses, roots_itmax, roots_compensated, flags=flags, **p)
        det = lens_eq_det_jac(z, nlenses, **p)
        return ((1.0 / torch.abs(det)) * mask).sum(0).reshape(w.shape)
    L = _lib.lib()
    lens = _c_lens(nlenses, x_cm, **p)
    shape = tuple(w.shape)
    if isinstance(w, torch.Tensor) and w.is_cuda:
        wf = w.to(torch.complex128).resolve_conj().resolve_neg().contiguous().reshape(-1)
        mag = torch.empty(wf.numel(), dtype=torch.float64, device=w.device)
        with torch.cuda.device(w.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.caustics_mag_point_source(wf.data_ptr(), mag.data_ptr(), None, wf.numel(),
                                                   lens, int(roots_itmax),
                                                   int(bool(roots_compensated)), int(flags), st))
        return mag.reshape(shape)
    _lib.require_cuda()
    is_t = isinstance(w, torch.Tensor)
    wn = np.ascontiguousarray(w.numpy() if is_t else np.asarray(w), dtype=np.complex128).reshape(-1)
    mag = np.empty(wn.size, dtype=np.float64)
    _lib.check(L.caustics_mag_point_source_host(wn.ctypes.data, mag.ctypes.data, wn.size, lens,
                                                int(roots_itmax), int(bool(roots_compensated)),
                                                int(flags)))
    mag = mag.reshape(shape)
    return torch.from_numpy(mag) if is_t else mag


GRID_WALK = 4   # CAUSTICS_FLAG_GRID_WALK (include/caustics_b200.h)


def mag_point_source_map(x0, y0, dx, dy, nx, ny, nlenses=2, rows=None, walk=True, roots_itmax=2500,
                         roots_compensated=False, device=None, out=None, **params):
    """Point-source magnification map (BASELINE config C5): `mag_point_source` on the regular grid
    w[iy, ix] = (x0 + ix dx) + i (y0 + iy dy) without materialising w (the reference's user would build
    the grid with jnp.meshgrid and call mag_point_source, point_source.py:1762-1830).  `rows = (begin,
    end)` computes a row block (how a multi-GPU driver shards the map).  `walk=True` solves each column
    segment as a warm-started walk (csrc/ps_walk.cuh: ~3x fewer root updates; agrees with the cold
    solves to rounding x conditioning), `walk=False` solves every pixel from cold, bit-identical to
    `mag_point_source` on the explicit grid.  Returns a (rows, nx) float64 CUDA tensor.
    `out`: a C-contiguous float64 HOST array of shape (rows, nx) (NumPy, e.g. a slice of a shared-memory map
    several ranks fill) -- the row block is then computed in chunks whose device-to-host copies overlap the
    kernels (caustics_mag_point_source_grid_host) and `out` is returned; or an int, the DEVICE address the
    block is written to (e.g. `PeerGather.ptr(...)`: another GPU's buffer), in which case nothing is returned."""
    if nlenses not in (2, 3):
        raise ValueError("mag_point_source_map supports nlenses = 2 or 3")
    _lib.require_cuda()
    r0, r1 = (0, int(ny)) if rows is None else (int(rows[0]), int(rows[1]))
    if not (0 <= r0 <= r1 <= int(ny)) or int(nx) <= 0:
        raise ValueError("bad map extent")
    p, x_cm = lens_params(nlenses, **params)
    lens = _c_lens(nlenses, x_cm, **p)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    fl = GRID_WALK if walk else 0
    if isinstance(out, np.ndarray):
        if out.dtype != np.float64 or out.shape != (r1 - r0, int(nx)) or not out.flags.c_contiguous:
            raise ValueError("`out` has to be a C-contiguous float64 array of shape (rows, nx)")
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().caustics_mag_point_source_grid_host(
                float(x0), float(y0), float(dx), float(dy), int(nx), r0, r1, out.ctypes.data, lens, int(roots_itmax),
                int(bool(roots_compensated)), fl))
        return out
    if isinstance(out, int):
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().caustics_mag_point_source_grid(
                float(x0), float(y0), float(dx), float(dy), int(nx), r0, r1, out, lens, int(roots_itmax),
                int(bool(roots_compensated)), fl, torch.cuda.current_stream().cuda_stream))
        return None
    mag = torch.empty((r1 - r0, int(nx)), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().caustics_mag_point_source_grid(
            float(x0), float(y0), float(dx), float(dy), int(nx), r0, r1, mag.data_ptr(), lens, int(roots_itmax),
            int(bool(roots_compensated)), GRID_WALK if walk else 0, torch.cuda.current_stream().cuda_stream))
    return mag


def critical_and_caustic_curves(npts=200, nlenses=2, **params):
    """Critical curves and caustics (point_source.py:1582-1649): the 2N roots of the critical-curve
    polynomial at `npts` phases phi in [-pi, pi] (kernel 1, degree 4 / 6), ordered into continuous
    curves by greedy track matching on the device, and mapped through the lens equation.  Returns
    (z_cr, z_ca), CUDA tensors of shape (2 * nlenses, npts), shifted by the centre of mass like the
    reference."""
    _lib.require_cuda()
    dev = torch.device("cuda")
    phi = torch.linspace(-math.pi, math.pi, npts, dtype=torch.float64, device=dev)
    if nlenses == 1:
        return torch.exp(-1j * phi), torch.zeros(npts, dtype=torch.complex128, device=dev)
    p, x_cm = lens_params(nlenses, **params)
    x = torch.exp(-1j * phi)
    one, zero = torch.ones_like(x), torch.zeros_like(x)
    a, e1 = p["a"], p["e1"]
    if nlenses == 2:                       # point_source.py:1482-1495
        coeffs = [x, zero, -2 * a**2 * x - 1.0, (-4 * a * e1 + 2 * a) * one, a**4 * x - a**2]
    elif nlenses == 3:                     # point_source.py:1498-1534
        e2, r3 = p["e2"], complex(p["r3"])
        coeffs = [x, -2 * x * r3, -2 * a**2 * x - 1 + x * r3**2,
                  4 * a**2 * x * r3 - 2 * a * e1 + 2 * a * e2 + 2 * e1 * r3 + 2 * e2 * r3,
                  a**4 * x - 3 * a**2 * e1 - 3 * a**2 * e2 + 2 * a**2 - 2 * a**2 * x * r3**2
                  + 4 * a * e1 * r3 - 4 * a * e2 * r3 - e1 * r3**2 - e2 * r3**2,
                  -2 * a**4 * x * r3 + 2 * a**2 * e1 * r3 + 2 * a**2 * e2 * r3 - 2 * a * e1 * r3**2 + 2 * a * e2 * r3**2,
                  a**4 * e1 + a**4 * e2 - a**4 + a**4 * x * r3**2 - a**2 * e1 * r3**2 - a**2 * e2 * r3**2]
    else:
        raise ValueError("`nlenses` has to be set to be <= 3.")
    c = torch.stack([torch.as_tensor(v, dtype=torch.complex128, device=dev) * one for v in coeffs], dim=-1)
    z = poly_roots(c)                                            # (npts, 2N), default itmax like the reference
    out = torch.empty_like(z)
    _lib.check(_lib.lib().caustics_match_tracks(z.data_ptr(), out.data_ptr(), 1, npts, z.shape[1],
                                                torch.cuda.current_stream().cuda_stream))
    z_cr = out.T.contiguous()
    z_ca = lens_eq(z_cr, nlenses, **p)
    return z_cr - x_cm, z_ca - x_cm
