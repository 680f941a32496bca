"""Host-side mirror of the reference's extended-source API
(/root/reference/src/caustics/extended_source.py:741-904 `mag_extended_source`, and
/root/reference/src/caustics/lightcurve.py:99-254 `mag`) on top of kernel family 3.

Same names, keyword arguments and defaults as the reference.  `w0` / `w_points` may be a Python
complex, a NumPy array or a torch tensor (CPU or CUDA) of any shape: every source position is an
independent unit of work and the whole batch is evaluated by one sequence of kernel launches
(the reference evaluates them one at a time under lax.map).  Results come back in the same kind of
container.  All compute runs on the GPU; there is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib
from .point_source import _c_lens, lens_params

__all__ = ["mag_extended_source", "mag", "mag_gate"]

_MAX_WS_BYTES = 24 << 30      # per-call workspace budget; larger batches are processed in chunks
_MAX_GATED_WS_BYTES = 3_950_000_000  # gated light curves: the survivors are integrated in windows of this much state
# (a 10^6-point binary light curve, 5.6e4 limb-darkened integrations, is ONE window of 3.92 GB: a second, short
# window costs 5 ms of latency-bound thread-per-source kernels)


def _to_device(w):
    """-> (flat complex128 CUDA tensor, restore function)"""
    if isinstance(w, torch.Tensor):
        shape, dev = tuple(w.shape), w.device
        flat = w.detach().to(device="cuda" if not w.is_cuda else dev, dtype=torch.complex128).reshape(-1).resolve_conj().resolve_neg().contiguous()
        if w.is_cuda:
            return flat, lambda m: m.reshape(shape)
        return flat, lambda m: m.cpu().reshape(shape)
    arr = np.asarray(w, dtype=np.complex128)
    shape = arr.shape
    flat = torch.from_numpy(np.ascontiguousarray(arr.reshape(-1))).cuda()
    if shape == ():
        return flat, lambda m: np.float64(m.cpu().numpy()[0])
    return flat, lambda m: m.cpu().numpy().reshape(shape)


_WS_MAG_ONLY = 4   # CAUSTICS_WS_MAG_ONLY: a uniform-disk workspace for magnifications alone (no track arrays)


def _chunk_len(L, n, nlenses, npts_limb, ld, npts_ld, budget=None, mag_only=False):
    ld = int(bool(ld)) or (_WS_MAG_ONLY if mag_only else 0)
    per1 = L.caustics_ext_workspace_bytes(1, nlenses, npts_limb, ld, npts_ld)
    if per1 == 0:
        raise ValueError("unsupported extended-source configuration "
                         "(need 8 <= npts_limb <= 1280, nlenses in 1..3, npts_ld <= 2048)")
    per = (L.caustics_ext_workspace_bytes(1024, nlenses, npts_limb, int(ld), npts_ld) + 1023) // 1024
    deg = 2 if nlenses == 1 else nlenses**2 + 1
    cap_idx = (2**31 - 1) // (deg * (npts_limb + 4) + 16)
    return max(1, min(n, (budget or _MAX_WS_BYTES) // per, cap_idx))


def _run(w, rho, nlenses, npts_limb, limb_darkening, u1, npts_ld, roots_itmax, roots_compensated,
         gate, q, params, return_test=False):
    _lib.require_cuda()
    L = _lib.lib()
    p, x_cm = lens_params(nlenses, **params) if nlenses > 1 else ({}, 0.0)
    lens = _c_lens(nlenses, x_cm, **p)
    flat, restore = _to_device(w)
    n = flat.numel()
    mag = torch.empty(n, dtype=torch.float64, device=flat.device)
    test = torch.empty(n, dtype=torch.uint8, device=flat.device) if gate else None
    if n == 0:
        return (restore(mag), restore(test.bool())) if return_test else restore(mag)
    rho = float(rho)
    # limb_darkening = "adaptive": CAUSTICS_LD_ADAPTIVE (fewer quadrature nodes on short far panels, opt-in)
    ld = (3 if limb_darkening == "adaptive" else 1) if limb_darkening else 0
    comp = int(bool(roots_compensated))
    cfg = (int(npts_limb), ld, float(u1), int(npts_ld), int(roots_itmax), comp)
    with torch.cuda.device(flat.device):
        st = torch.cuda.current_stream().cuda_stream
        if gate and nlenses == 2 and n < 2**31 and not torch.cuda.is_current_stream_capturing():
            # Two-call form: the gate alone, then the workspace is sized by the points that FAILED it
            # (a few per cent of a light curve; one 4-byte read-back) instead of by n.
            lst = torch.empty(n, dtype=torch.int32, device=flat.device)
            cnt = torch.empty(1, dtype=torch.int32, device=flat.device)
            _lib.check(L.caustics_mag_gate(flat.data_ptr(), mag.data_ptr(), test.data_ptr(), lst.data_ptr(),
                                           cnt.data_ptr(), n, rho, lens, float(q), cfg[4], comp, st))
            nfull = int(cnt.item())
            if nfull:
                chunk = _chunk_len(L, nfull, nlenses, npts_limb, limb_darkening, npts_ld, _MAX_GATED_WS_BYTES, mag_only=True)
                chunk = -(-nfull // -(-nfull // chunk))      # windows of equal length
                nbytes = L.caustics_mag_workspace_bytes(nfull, chunk, nlenses, cfg[0], ld, cfg[3])
                ws = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
                _lib.check(L.caustics_mag_extended_source_list(flat.data_ptr(), mag.data_ptr(), lst.data_ptr(),
                                                               cnt.data_ptr(), nfull, rho, lens, *cfg,
                                                               ws.data_ptr(), nbytes, st))
        elif gate:
            # one stream-ordered call (CUDA-graph capturable; triple lens: every point is integrated):
            # the survivor count stays on the device, the workspace holds `chunk` sources at a time
            chunk = _chunk_len(L, n, nlenses, npts_limb, limb_darkening, npts_ld, mag_only=True)
            nbytes = L.caustics_mag_workspace_bytes(n, chunk, nlenses, cfg[0], ld, cfg[3])
            ws = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
            _lib.check(L.caustics_mag(flat.data_ptr(), mag.data_ptr(), test.data_ptr(), n, rho, lens, float(q),
                                      *cfg, ws.data_ptr(), nbytes, st))
        else:
            chunk = _chunk_len(L, n, nlenses, npts_limb, limb_darkening, npts_ld, mag_only=True)
            nbytes = L.caustics_ext_workspace_bytes(chunk, nlenses, cfg[0], ld or _WS_MAG_ONLY, cfg[3])
            ws = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
            for off in range(0, n, chunk):
                m = min(chunk, n - off)
                _lib.check(L.caustics_mag_extended_source(flat.data_ptr() + 16 * off, mag.data_ptr() + 8 * off, m,
                                                          rho, lens, *cfg, ws.data_ptr(), nbytes, st))
    if return_test:
        return restore(mag), restore(test.bool())
    return restore(mag)


def _requires_grad(*vals):
    return any(isinstance(v, torch.Tensor) and v.requires_grad for v in vals)


def _detached(v):
    return v.detach() if isinstance(v, torch.Tensor) else v


def _get_contours(w0, rho, nlenses, npts_limb, roots_itmax, roots_compensated, params):
    """Closed image contours of every source (caustics_ext_contours): dict of device tensors
    vz, vth, vcid (nv, n), valid (nv, n), cpar (CMAX, n), plus the flattened centres `wf`."""
    import ctypes
    _lib.require_cuda()
    L = _lib.lib()
    dev = w0.device if isinstance(w0, torch.Tensor) and w0.is_cuda else torch.device("cuda")
    w0t = (w0.to(dev, torch.complex128) if isinstance(w0, torch.Tensor)
           else torch.as_tensor(w0, dtype=torch.complex128, device=dev))
    wf = w0t.reshape(-1)
    n = wf.numel()
    p, x_cm = lens_params(nlenses, **params) if nlenses > 1 else ({}, 0.0)
    p0 = {k: _detached(v) for k, v in p.items()}
    lens = _c_lens(nlenses, _detached(x_cm), **p0)
    vmax, cmax = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(L.caustics_ext_contour_capacity(nlenses, int(npts_limb), vmax, cmax))
    VM, CM = vmax.value, cmax.value
    vz = torch.empty((VM, n), dtype=torch.complex128, device=dev)
    vth = torch.empty((VM, n), dtype=torch.float64, device=dev)
    vcid = torch.empty((VM, n), dtype=torch.uint8, device=dev)
    vcount = torch.empty(n, dtype=torch.int32, device=dev)
    cpar = torch.zeros((CM, n), dtype=torch.float64, device=dev)
    cstart = torch.empty((CM + 1, n), dtype=torch.int32, device=dev)
    ncont = torch.empty(n, dtype=torch.int32, device=dev)
    nbytes = L.caustics_ext_workspace_bytes(n, nlenses, int(npts_limb), 0, 100)
    with torch.cuda.device(dev):
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        wd = wf.detach().resolve_conj().resolve_neg().contiguous()
        _lib.check(L.caustics_ext_contours(wd.data_ptr(), None, n, float(_detached(rho)), lens, int(npts_limb),
                                           int(roots_itmax), int(bool(roots_compensated)), ws.data_ptr(), nbytes,
                                           vz.data_ptr(), vth.data_ptr(), vcid.data_ptr(), vcount.data_ptr(),
                                           cpar.data_ptr(), cstart.data_ptr(), ncont.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
    nv = int(vcount.max().item())
    valid = torch.arange(nv, device=dev)[:, None] < vcount[None, :]
    return {"vz": vz[:nv], "vth": vth[:nv], "vcid": vcid[:nv].long(), "valid": valid, "cpar": cpar,
            "wf": wf, "shape": tuple(w0t.shape)}


def _mag_from_contours(cont, wf, rho, nlenses, params, newton_steps=1, ld=None):
    """Uniform-disk magnification from a fixed set of contour vertices (fixed limb angles, fixed
    topology) as a differentiable function of the source centres `wf`, `rho` and the lens parameters.
    Every vertex is an image of w = wf + x_cm + rho e^{i theta}, a zero of F(z) = lens_eq(z) - w.
    One Newton step z0 - J^{-1} F(z0; params), with J = dF/d(z, zbar) frozen at the kernel's z0,
    leaves the value unchanged (F(z0) = 0) and carries the exact first-order dependence -- the
    implicit-function rule of ehrlich_aberth_primitive.py:254-324 written on the lens equation.
    `newton_steps` > 1 re-polishes the vertices (used to evaluate nearby parameters)."""
    from .point_source import lens_eq, _lenses
    vz, vth, vcid, valid, cpar = cont["vz"], cont["vth"], cont["vcid"], cont["valid"], cont["cpar"]
    dev = vz.device
    p, x_cm = lens_params(nlenses, **params) if nlenses > 1 else ({}, 0.0)
    rho_t = torch.as_tensor(rho, dtype=torch.float64, device=dev)
    w = (wf + x_cm)[None, :] + rho_t * torch.exp(1j * vth)
    z = torch.where(valid, vz, torch.ones_like(vz))              # keep padding away from the lenses
    for _ in range(newton_steps):
        zd = z.detach()
        if nlenses == 1:
            F = z - 1.0 / torch.conj(z) - w
            g = 1.0 / torch.conj(zd) ** 2
        else:
            F = lens_eq(z, nlenses, **p) - w
            r0, e0 = _lenses(nlenses, **{k: _detached(v) for k, v in p.items()})
            g = sum(ej / (torch.conj(zd) - (torch.conj(rj) if isinstance(rj, torch.Tensor) else np.conj(rj))) ** 2
                    for rj, ej in zip(r0, e0))
        g = g.detach()
        z = z + (-F + g * torch.conj(F)) / (1.0 - torch.abs(g) ** 2)
    x, y = z.real, z.imag
    edge = (vcid[:-1] == vcid[1:]) & valid[1:]
    par = torch.gather(cpar, 0, vcid.clamp(max=cpar.shape[0] - 1))
    if ld is None:
        cross = x[:-1] * y[1:] - x[1:] * y[:-1]
        total = 0.5 * (torch.where(edge, cross, torch.zeros_like(cross)) * par[:-1]).sum(0)
        return (torch.abs(total) / (np.pi * rho_t**2)).reshape(cont["shape"])
    # limb darkening: Dominik (1998) P/Q integrals, integrate.py:47-121, in differentiable form
    u1, npts_ld = ld
    u1 = torch.as_tensor(u1, dtype=torch.float64, device=dev)
    CM, n = cpar.shape[0], z.shape[1]
    cid = vcid.clamp(max=CM - 1)
    zsum = torch.zeros((CM, n), dtype=torch.complex128, device=dev).scatter_add(0, cid, torch.where(valid, z, torch.zeros_like(z)))
    cnt = torch.zeros((CM, n), dtype=torch.float64, device=dev).scatter_add(0, cid, valid.double())
    z0 = torch.gather(zsum / cnt.clamp(min=1.0), 0, cid)       # contour centroid incl. the closing vertex
    w0 = (wf + x_cm)[None, :, None]
    n1 = int(npts_ld / 2)
    n2 = npts_ld - n1

    def brightness(zz):
        if nlenses == 1:
            ww = zz - 1.0 / torch.conj(zz)
        else:
            ww = lens_eq(zz, nlenses, **p)
        r2 = ((ww - w0).real ** 2 + (ww - w0).imag ** 2) / rho_t**2
        inside = r2 <= 1.0
        ain = torch.where(inside, 1.0 - r2, torch.ones_like(r2)).clamp(min=0.0)
        aout = torch.where(inside, torch.ones_like(r2), 1.0 - 1.0 / r2).clamp(min=0.0)
        safe = lambda t: torch.sqrt(torch.where(t > 0, t, torch.ones_like(t))) * (t > 0)
        B = torch.where(inside, 1.0 + safe(ain), 1.0 - safe(aout))
        return 3.0 / (3.0 - u1) * (u1 * B + 1.0 - 2.0 * u1)

    def two_panel(a, b, point):
        ad = torch.abs(b - a)
        split = torch.where(b > a, b - 2 * rho_t, b + 2 * rho_t)
        split = torch.where(0.5 * ad <= 2 * rho_t, a + 0.5 * ad, split)
        tot = 0.0
        for lo, hi, nn in ((a, split, n1), (split, b, n2)):
            xg, wg = np.polynomial.legendre.leggauss(nn)
            xg = torch.as_tensor(xg, device=dev)[None, None, :]
            wg = torch.as_tensor(wg, device=dev)[None, None, :]
            hw, mid = 0.5 * (hi - lo)[..., None], 0.5 * (hi + lo)[..., None]
            tot = tot + (hw * brightness(point(hw * xg + mid)) * wg).sum(-1)
        return tot

    P = -0.5 * two_panel(z0.imag, y, lambda t: torch.complex(x[..., None].expand_as(t), t))
    Q = 0.5 * two_panel(z0.real, x, lambda t: torch.complex(t, y[..., None].expand_as(t)))
    seg = 0.5 * (P[:-1] + P[1:]) * (x[1:] - x[:-1]) + 0.5 * (Q[:-1] + Q[1:]) * (y[1:] - y[:-1])
    total = (torch.where(edge, seg, torch.zeros_like(seg)) * par[:-1]).sum(0)
    return (torch.abs(total) / (np.pi * rho_t**2)).reshape(cont["shape"])


class _UniformMagGrad(torch.autograd.Function):
    """Uniform-disk magnification whose backward pass is the tangent the kernels emit alongside the forward
    pass (caustics_mag_extended_source_grad: d mag / d(a, e1, e2, r3, w + x_cm, rho) per source, the
    implicit-function rule on every contour vertex pushed through the trapezoid sum on the device).  The
    chain from the low-level to the user's (s, q, q3, r3, psi) parameters is ordinary torch autograd around
    this Function (`lens_params`), i.e. the JVP rule itself stays in Python."""

    @staticmethod
    def forward(ctx, wf, rho, a, e1, e2, r3, x_cm, nlenses, npts_limb, itmax, comp):
        L = _lib.lib()
        lens = _lib.Lens()
        lens.nlenses, lens.x_cm = nlenses, float(x_cm)
        if nlenses >= 2:
            lens.a, lens.e1 = float(a), float(e1)
        if nlenses == 3:
            lens.e2, lens.r3_re, lens.r3_im = float(e2), float(r3.real), float(r3.imag)
        wd = wf.detach().to(torch.complex128).resolve_conj().resolve_neg().contiguous()
        n = wd.numel()
        mag = torch.empty(n, dtype=torch.float64, device=wd.device)
        chunk = _chunk_len(L, n, nlenses, npts_limb, False, 100)
        parts = []
        with torch.cuda.device(wd.device):
            st = torch.cuda.current_stream().cuda_stream
            nbytes = L.caustics_ext_workspace_bytes(chunk, nlenses, int(npts_limb), 0, 100)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=wd.device)
            for off in range(0, n, chunk):
                m = min(chunk, n - off)
                g = torch.empty((8, m), dtype=torch.float64, device=wd.device)
                _lib.check(L.caustics_mag_extended_source_grad(wd.data_ptr() + 16 * off, mag.data_ptr() + 8 * off,
                                                               g.data_ptr(), m, float(rho), lens, int(npts_limb),
                                                               int(itmax), int(bool(comp)), ws.data_ptr(), nbytes, st))
                parts.append(g)
        ctx.save_for_backward(parts[0] if len(parts) == 1 else torch.cat(parts, dim=1))
        return mag

    @staticmethod
    def backward(ctx, gout):
        G, = ctx.saved_tensors
        t = G * gout[None, :]
        s = t.sum(dim=1)
        # complex inputs: torch's convention for a real loss is dL/dRe + i dL/dIm
        return (torch.complex(t[5], t[6]), s[7], s[0], s[1], s[2], torch.complex(s[3], s[4]), s[5],
                None, None, None, None)


def _mag_uniform_kernel_grad(w0, rho, nlenses, npts_limb, roots_itmax, roots_compensated, params):
    """differentiable uniform-disk magnification through the fused forward + tangent kernel"""
    _lib.require_cuda()
    dev = w0.device if isinstance(w0, torch.Tensor) and w0.is_cuda else torch.device("cuda")
    w0t = (w0.to(dev, torch.complex128) if isinstance(w0, torch.Tensor)
           else torch.as_tensor(w0, dtype=torch.complex128, device=dev))
    p, x_cm = lens_params(nlenses, **params) if nlenses > 1 else ({}, 0.0)
    R = lambda v: v.to(dev, torch.float64) if isinstance(v, torch.Tensor) else torch.tensor(float(v), dtype=torch.float64, device=dev)
    C = lambda v: v.to(dev, torch.complex128) if isinstance(v, torch.Tensor) else torch.tensor(complex(v), dtype=torch.complex128, device=dev)
    out = _UniformMagGrad.apply(w0t.reshape(-1), R(rho), R(p.get("a", 0.0)), R(p.get("e1", 0.0)), R(p.get("e2", 0.0)),
                                C(p.get("r3", 0.0)), R(x_cm), int(nlenses), int(npts_limb), int(roots_itmax),
                                bool(roots_compensated))
    return out.reshape(tuple(w0t.shape))


def _mag_uniform_differentiable(w0, rho, nlenses, npts_limb, roots_itmax, roots_compensated, params, ld=None):
    """Magnification with gradients w.r.t. w0, rho, u1 and the lens parameters (sampling, masks and
    contour topology are constants, exactly as in the reference's jax.grad).
    Uniform disk: forward value and tangent come out of ONE kernel pass (`_UniformMagGrad`).
    `ld` = (u1, npts_ld), limb darkening: contours from the kernels, the Dominik P/Q quadrature and its
    gradient in torch on those vertices (`_mag_from_contours`, which is also the executable specification the
    kernel tangent is tested against); sources are processed in slices so the (vertex, source, node)
    quadrature tensors stay small."""
    if ld is None:
        return _mag_uniform_kernel_grad(w0, rho, nlenses, npts_limb, roots_itmax, roots_compensated, params)
    cont = _get_contours(w0, rho, nlenses, npts_limb, roots_itmax, roots_compensated, params)
    n = cont["wf"].numel()
    outs = []
    for lo in range(0, n, 16):
        sl = slice(lo, min(n, lo + 16))
        sub = {k: (v[:, sl] if k in ("vz", "vth", "vcid", "valid", "cpar") else v) for k, v in cont.items()}
        sub["shape"] = (sl.stop - sl.start,)
        outs.append(_mag_from_contours(sub, cont["wf"][sl], rho, nlenses, params, ld=ld))
    return torch.cat(outs).reshape(cont["shape"])


def mag_extended_source(w0, rho, nlenses=2, npts_limb=150, limb_darkening=False, u1=0.0, npts_ld=100,
                        roots_itmax=2500, roots_compensated=False, **params):
    """Magnification of a (limb-darkened) disk of radius `rho` centred on `w0` by contour
    integration in the image plane; arguments as in the reference (extended_source.py:741-805).
    `limb_darkening="adaptive"` (no reference counterpart) integrates short far panels of the
    Dominik P/Q integrals with a lower-order Gauss-Legendre rule; True is the reference's quadrature."""
    if nlenses not in (1, 2, 3):
        raise ValueError("`nlenses` has to be set to be <= 3.")
    if _requires_grad(w0, rho, u1, *params.values()):
        return _mag_uniform_differentiable(w0, rho, nlenses, npts_limb, roots_itmax, roots_compensated, params,
                                           ld=(u1, npts_ld) if limb_darkening else None)
    return _run(w0, rho, nlenses, npts_limb, limb_darkening, u1, npts_ld, roots_itmax, roots_compensated,
                False, 0.0, params)


def mag_gate(w_points, rho, roots_itmax=2500, roots_compensated=False, **params):
    """Binary-lens gate of `mag` on its own (lightcurve.py:202-225): returns (mu_hexadecapole, valid)
    for every point -- the cheap pass a multi-GPU driver runs first to balance the expensive full
    integrations (caustics_b200.sharding.balanced_order)."""
    _lib.require_cuda()
    L = _lib.lib()
    p, x_cm = lens_params(2, **params)
    lens = _c_lens(2, x_cm, **p)
    flat, restore = _to_device(w_points)
    n = flat.numel()
    mag_ = torch.empty(n, dtype=torch.float64, device=flat.device)
    used = torch.empty(n, dtype=torch.uint8, device=flat.device)
    lst = torch.empty(max(n, 1), dtype=torch.int32, device=flat.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=flat.device)
    with torch.cuda.device(flat.device):
        _lib.check(L.caustics_mag_gate(flat.data_ptr(), mag_.data_ptr(), used.data_ptr(), lst.data_ptr(),
                                       cnt.data_ptr(), n, float(rho), lens, float(params.get("q", 1.0)),
                                       int(roots_itmax), int(bool(roots_compensated)),
                                       torch.cuda.current_stream().cuda_stream))
    return restore(mag_), restore(used.bool())


def mag(w_points, rho, nlenses=2, npts_limb=200, limb_darkening=False, u1=0.0, npts_ld=100,
        roots_itmax=2500, roots_compensated=False, return_test=False, **params):
    """Light-curve magnification (lightcurve.py:99-254): the hexadecapole approximation where the
    reference's validity tests pass (binary lens), full contour integration elsewhere.  For the
    triple lens the reference forces full integration everywhere but leaves `mu_multi` unassigned and
    cannot run; here it runs, with full integration at every point.
    `return_test=True` also returns the boolean "hexadecapole was used" per point."""
    if nlenses not in (1, 2, 3):
        raise ValueError("nlenses must be <= 3")
    q = params.get("q", 1.0) if nlenses == 2 else 1.0
    if _requires_grad(w_points, rho, u1, *params.values()):
        return _mag_differentiable(w_points, rho, nlenses, npts_limb, roots_itmax, roots_compensated,
                                   params, return_test, ld=(u1, npts_ld) if limb_darkening else None)
    return _run(w_points, rho, nlenses, npts_limb, limb_darkening, u1, npts_ld, roots_itmax,
                roots_compensated, True, _detached(q), params, return_test=return_test)


def _implicit_images(z0, w, nlenses, p):
    """Images as differentiable functions of (w, lens parameters): one Newton step on the lens equation
    from the kernel's images z0 with the Jacobian frozen there (see _mag_from_contours)."""
    from .point_source import lens_eq, _lenses
    F = lens_eq(z0, nlenses, **p) - w
    r0, e0 = _lenses(nlenses, **{k: _detached(v) for k, v in p.items()})
    g = sum(ej / (torch.conj(z0) - (torch.conj(rj) if isinstance(rj, torch.Tensor) else np.conj(rj))) ** 2
            for rj, ej in zip(r0, e0)).detach()
    return z0 + (-F + g * torch.conj(F)) / (1.0 - torch.abs(g) ** 2)


def _mag_differentiable(w_points, rho, nlenses, npts_limb, roots_itmax, roots_compensated, params, return_test,
                        ld=None):
    """`mag` with gradients (uniform disk): the gate decision comes from the kernels and is a constant,
    as in the reference where lax.cond predicates carry no gradient; hexadecapole points go through
    the torch form of the Cassan expansion evaluated at implicitly-differentiated images, the others
    through the differentiable contour integration."""
    from .multipole import hexadecapole_terms
    from .point_source import _images_point_source, _lenses
    dev = w_points.device if isinstance(w_points, torch.Tensor) and w_points.is_cuda else torch.device("cuda")
    w = (w_points.to(dev, torch.complex128) if isinstance(w_points, torch.Tensor)
         else torch.as_tensor(w_points, dtype=torch.complex128, device=dev))
    shape = tuple(w.shape)
    wf = w.reshape(-1)
    detp = {k: _detached(v) for k, v in params.items()}
    q = detp.get("q", 1.0) if nlenses == 2 else 1.0
    _, used = _run(wf.detach(), float(_detached(rho)), nlenses, npts_limb, False, 0.0, 100, roots_itmax,
                   roots_compensated, True, q, detp, return_test=True)
    out = torch.zeros(wf.numel(), dtype=torch.float64, device=dev)
    hx = torch.nonzero(used).reshape(-1)
    fl = torch.nonzero(~used).reshape(-1)
    if hx.numel():
        p, x_cm = lens_params(nlenses, **params)
        ws = wf[hx] + x_cm
        p0 = {k: _detached(v) for k, v in p.items()}
        z0, mask = _images_point_source(ws.detach(), nlenses, roots_itmax, roots_compensated, **p0)
        z = _implicit_images(torch.where(mask, z0, torch.ones_like(z0)), ws[None, :], nlenses, p)
        r, eps = _lenses(nlenses, **p)
        mu0, dq, dh = hexadecapole_terms(z, torch.as_tensor(rho, dtype=torch.float64, device=dev), 0.0, r, eps)
        val = torch.where(mask, torch.abs(mu0 + dq + dh), torch.zeros_like(mu0)).sum(0)
        out = out.index_put((hx,), val)
    if fl.numel():
        val = _mag_uniform_differentiable(wf[fl], rho, nlenses, npts_limb, roots_itmax, roots_compensated, params,
                                          ld=ld)
        out = out.index_put((fl,), val)
    out = out.reshape(shape)
    return (out, used.reshape(shape)) if return_test else out
