"""Host-side mirror of the reference's `ehrlich_aberth` primitive and its `poly_roots` wrapper
(/root/reference/src/caustics/ehrlich_aberth_primitive.py:34-125) on top of the C ABI.

Same names, argument meaning and shapes as the reference.  Array kinds:
  * torch CUDA tensor  -> device launcher on torch's current stream, returns a CUDA tensor (no sync)
  * numpy array / torch CPU tensor -> the HOST entry point (chunked H2D/kernel/D2H pipeline),
    returns the same kind.
Only complex128 is accepted, like the reference (ehrlich_aberth_primitive.py:187-190).

Differentiation: `poly_roots` on torch tensors that require grad goes through
`_PolyRoots` (torch.autograd.Function) whose backward is the reference's implicit-function rule
(ehrlich_aberth_primitive.py:254-324): dz_j = -(sum_k dp_k z_j^k) / p'(z_j).  The same rule in
JAX form lives in caustics_b200/jax_glue.py.
"""
import numpy as np
import torch

from . import _lib

__all__ = ["poly_roots", "ehrlich_aberth", "roots_jvp"]


def _solve_flat(coeffs, roots_init, itmax, compensated, custom_init, flags, return_sweeps=False,
                out=None):
    """coeffs (size, deg+1), roots_init (size, deg) or None -> roots (size, deg) [, sweeps]"""
    L = _lib.lib()
    if isinstance(coeffs, torch.Tensor) and coeffs.is_cuda:
        if coeffs.dtype != torch.complex128:
            raise NotImplementedError(f"Unsupported dtype {coeffs.dtype}")
        c = coeffs.resolve_conj().resolve_neg().contiguous()   # data_ptr ignores lazy conj/neg bits
        size, deg = c.shape[0], c.shape[1] - 1
        ri_ptr = None
        if custom_init:
            ri = roots_init.to(device=c.device, dtype=torch.complex128).resolve_conj().resolve_neg().contiguous()
            if ri.shape != (size, deg):
                raise ValueError("roots_init must have shape (size, deg)")
            ri_ptr = ri.data_ptr()
        roots = torch.empty((size, deg), dtype=torch.complex128, device=c.device)
        sweeps = torch.empty(size, dtype=torch.int32, device=c.device) if return_sweeps else None
        with torch.cuda.device(c.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.caustics_ea_solve(c.data_ptr(), ri_ptr, roots.data_ptr(),
                                           sweeps.data_ptr() if return_sweeps else None, size, deg,
                                           int(itmax), int(bool(compensated)), int(bool(custom_init)),
                                           int(flags), st))
        return (roots, sweeps) if return_sweeps else roots
    # host path
    is_t = isinstance(coeffs, torch.Tensor)
    c = coeffs.numpy() if is_t else np.asarray(coeffs)
    if c.dtype != np.complex128:
        raise NotImplementedError(f"Unsupported dtype {c.dtype}")
    _lib.require_cuda()
    c = np.ascontiguousarray(c)
    size, deg = c.shape[0], c.shape[1] - 1
    ri_ptr = None
    if custom_init:
        ri = roots_init.numpy() if isinstance(roots_init, torch.Tensor) else np.asarray(roots_init)
        ri = np.ascontiguousarray(ri, dtype=np.complex128).reshape(size, deg)
        ri_ptr = ri.ctypes.data
    if out is not None:
        roots = out.reshape(size, deg)  # caller-provided (e.g. pinned) result buffer
        assert roots.dtype == np.complex128 and roots.flags.c_contiguous
    else:
        roots = np.empty((size, deg), dtype=np.complex128)
    sweeps = np.empty(size, dtype=np.int32) if return_sweeps else None
    _lib.check(L.caustics_ea_solve_host(c.ctypes.data, ri_ptr, roots.ctypes.data,
                                        sweeps.ctypes.data if return_sweeps else None, size, deg,
                                        int(itmax), int(bool(compensated)), int(bool(custom_init)),
                                        int(flags)))
    if is_t:
        roots = torch.from_numpy(roots)
        sweeps = torch.from_numpy(sweeps) if return_sweeps else None
    return (roots, sweeps) if return_sweeps else roots


def ehrlich_aberth(coeffs, roots_init, itmax=None, compensated=None, custom_init=False, flags=0):
    """The primitive: coeffs (size, deg+1) complex128 LOW->HIGH, roots_init (size, deg);
    returns the FLAT (size*deg,) roots like the reference (ehrlich_aberth_primitive.py:98-125)."""
    if coeffs.ndim != 2:
        raise ValueError("coeffs must have shape (size, deg + 1)")
    itmax = 2000 if itmax is None else itmax
    roots = _solve_flat(coeffs, roots_init, itmax, bool(compensated), custom_init, flags)
    return roots.reshape(-1)


def roots_jvp(coeffs_low_high, roots, dcoeffs):
    """Tangent of the roots for a coefficient tangent (ehrlich_aberth_primitive.py:304-319):
    dz = -(sum_k dp_k z^k) / p'(z).  Works on numpy arrays or torch tensors; coeffs/dcoeffs
    (size, deg+1) low->high, roots (size, deg)."""
    if isinstance(roots, torch.Tensor) and roots.is_cuda and roots.dim() == 2:
        p = coeffs_low_high.to(torch.complex128).resolve_conj().resolve_neg().contiguous()
        z = roots.resolve_conj().resolve_neg().contiguous()
        dp = dcoeffs.to(torch.complex128).resolve_conj().resolve_neg().contiguous()
        out = torch.empty_like(z)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().caustics_ea_jvp(p.data_ptr(), z.data_ptr(), dp.data_ptr(), out.data_ptr(),
                                                  z.shape[0], z.shape[1], torch.cuda.current_stream().cuda_stream))
        return out
    xp = torch if isinstance(roots, torch.Tensor) else np
    deg = coeffs_low_high.shape[-1] - 1
    # Horner for sum_k dp_k z^k and for p'(z) (no (size, deg, deg+1) temporary)
    num = xp.zeros_like(roots)
    der = xp.zeros_like(roots)
    for k in range(deg, -1, -1):
        num = num * roots + dcoeffs[..., k:k + 1]
        if k >= 1:
            der = der * roots + k * coeffs_low_high[..., k:k + 1]
    return -num / der


class _PolyRoots(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeffs_low_high, roots_init, itmax, compensated, custom_init, flags):
        roots = _solve_flat(coeffs_low_high.detach(), roots_init, itmax, compensated, custom_init, flags)
        ctx.save_for_backward(coeffs_low_high.detach(), roots)
        return roots

    @staticmethod
    def backward(ctx, grad_roots):
        p, z = ctx.saved_tensors
        # z_j = h_j(p) is holomorphic in p: dz_j/dp_k = -z_j^k / p'(z_j); torch's convention for a
        # holomorphic map is grad_p = conj(dz/dp) * grad_z (the tangent w.r.t. roots_init is zero,
        # ehrlich_aberth_primitive.py:299-302).  On the device this is one memory-bound kernel.
        if p.is_cuda:
            g = grad_roots.to(torch.complex128).resolve_conj().resolve_neg().contiguous()
            out = torch.empty_like(p)
            with torch.cuda.device(p.device):
                _lib.check(_lib.lib().caustics_ea_vjp(p.data_ptr(), z.data_ptr(), g.data_ptr(), out.data_ptr(),
                                                      p.shape[0], p.shape[1] - 1,
                                                      torch.cuda.current_stream().cuda_stream))
            return out, None, None, None, None, None
        deg = p.shape[1] - 1
        der = torch.zeros_like(z)
        for k in range(deg, 0, -1):
            der = der * z + k * p[:, k:k + 1]
        g = grad_roots / torch.conj(-der)
        zc = torch.conj(z)
        cols, zk = [], torch.ones_like(z)
        for k in range(deg + 1):
            cols.append((g * zk).sum(dim=1))
            zk = zk * zc
        return torch.stack(cols, dim=1), None, None, None, None, None


def poly_roots(coeffs, itmax=2000, compensated=False, custom_init=False, roots_init=None, flags=0,
               out=None, check=False):
    """Roots of complex polynomials; the last axis of `coeffs` holds the coefficients starting from
    the HIGHEST order term (reference docstring, ehrlich_aberth_primitive.py:49-65).  Returns the
    same shape with the last axis shrunk by one.
    `check=True` (no reference counterpart: the reference prints "not all roots converged" from C++) also
    fetches the per-polynomial sweep counts and warns when a polynomial ran into `itmax`."""
    ncoeffs = coeffs.shape[-1]
    out_shape = tuple(coeffs.shape[:-1]) + (ncoeffs - 1,)
    flat = coeffs.reshape(-1, ncoeffs)
    ri = None
    if custom_init:
        if roots_init is None:
            raise ValueError("custom_init=True requires roots_init")
        ri = roots_init.reshape(flat.shape[0], ncoeffs - 1)
    if isinstance(flat, torch.Tensor) and flat.requires_grad:
        roots = _PolyRoots.apply(torch.flip(flat, dims=[1]), ri, itmax, compensated, custom_init, flags)
    else:
        # the kernel reads the rows back to front instead of materialising coeffs[:, ::-1]
        if check and out is None:
            roots, sweeps = _solve_flat(flat, ri, itmax, compensated, custom_init,
                                        flags | _lib.FLAG_COEFFS_HIGH_FIRST, return_sweeps=True)
            bad = int((sweeps < 0).sum())
            if bad:
                import warnings
                warnings.warn(f"poly_roots: {bad} of {flat.shape[0]} polynomials did not converge within itmax={itmax}",
                              RuntimeWarning, stacklevel=2)
        else:
            roots = _solve_flat(flat, ri, itmax, compensated, custom_init,
                                flags | _lib.FLAG_COEFFS_HIGH_FIRST, out=out)
    return roots.reshape(out_shape)
