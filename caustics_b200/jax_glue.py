"""JAX registration of the sm_100a kernels (UNEXECUTED in this image: jax is not installable here,
SURVEY 0.4).  It is the reference's `ehrlich_aberth_primitive.py` with the translation rule
re-targeted at `caustics_ea_xla`; the abstract-eval, JVP and batching rules are the reference's and
stay in Python, so jit / vmap / grad keep working (north_star).

Reference lines replaced:
  ehrlich_aberth_primitive.py:16-28    capsule registration  -> _register() below
  ehrlich_aberth_primitive.py:131-244  XLA translation rule  -> _lowering() (jax.ffi.ffi_lowering,
                                        legacy custom-call api_version=0 with an opaque descriptor)
  kernels.h:11-17 / gpu_ops.cc:24-25   descriptor packing    -> _descriptor()
"""
import ctypes

import numpy as np

from . import _lib

_TARGET = "caustics_b200_ehrlich_aberth"


def _capsule(fn_ptr):
    """PyCapsule named like the reference's (pybind11_kernel_helpers.h:21-24)."""
    new = ctypes.pythonapi.PyCapsule_New
    new.restype = ctypes.py_object
    new.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
    return new(fn_ptr, b"xla._CUSTOM_CALL_TARGET", None)


def _descriptor(size, deg, itmax, compensated, custom_init, flags=0):
    d = _lib.EADescriptor()
    _lib.lib().caustics_ea_make_descriptor(ctypes.byref(d), size, deg, itmax, int(compensated),
                                           int(custom_init), flags)
    return bytes(d)


def register():
    """Build the jax primitive.  Returns (poly_roots, ehrlich_aberth) with the reference signatures."""
    import jax
    import jax.numpy as jnp
    from jax import core, lax, vmap
    from jax.interpreters import ad, batching, mlir

    L = _lib.lib()
    ptr = ctypes.cast(L.caustics_ea_xla, ctypes.c_void_p).value
    jax.ffi.register_ffi_target(_TARGET, _capsule(ptr), platform="CUDA", api_version=0)

    prim = core.Primitive("ehrlich_aberth")

    def ehrlich_aberth(coeffs, roots_init, itmax=None, compensated=None, custom_init=False):
        return prim.bind(coeffs, roots_init, itmax=itmax, compensated=compensated, custom_init=custom_init)

    def _abstract(coeffs, roots_init, **kw):          # ehrlich_aberth_primitive.py:115-125
        ncoeffs = coeffs.shape[-1]
        return core.ShapedArray((coeffs.shape[0] * (ncoeffs - 1),), coeffs.dtype)

    def _lowering(ctx, coeffs, roots_init, *, itmax, compensated, custom_init):
        aval = ctx.avals_in[0]
        if aval.dtype != np.complex128:
            raise NotImplementedError(f"Unsupported dtype {aval.dtype}")     # :187-190
        size, deg = aval.shape[0], aval.shape[1] - 1
        opaque = _descriptor(size, deg, itmax, compensated, custom_init)
        # legacy (untyped) custom call: XLA's API_VERSION_ORIGINAL = 1 is the signature without a status
        # argument, void(stream, buffers, opaque, opaque_len), which is what the reference's gpu_ehrlich_aberth
        # and caustics_ea_xla have; ffi_lowering defaults to 4 (typed FFI), which takes a dict, not bytes
        return jax.ffi.ffi_lowering(_TARGET, backend_config=opaque, api_version=1)(ctx, coeffs, roots_init)

    def _jvp(args, tangents, itmax=None, compensated=False, custom_init=False):   # :254-324
        p, roots_init = args
        dp = tangents[0]
        size, deg = p.shape[0], p.shape[1] - 1
        z = prim.bind(p, roots_init, itmax=itmax, compensated=compensated,
                      custom_init=custom_init).reshape((size, deg))
        dp = jnp.zeros_like(p) if type(dp) is ad.Zero else dp
        num = jnp.zeros_like(z)
        der = jnp.zeros_like(z)
        for k in range(deg, -1, -1):                  # Horner: no (size, deg, deg+1) temporary
            num = num * z + dp[:, k:k + 1]
            if k >= 1:
                der = der * z + k * p[:, k:k + 1]
        return z.reshape(-1), (-num / der).reshape(-1)

    def _batch(args, axes, **kw):                     # :330-353
        coeffs, roots_init = args
        ncoeffs, nroots = coeffs.shape[-1], roots_init.shape[-1]
        out_shape = coeffs.shape[:-1] + (ncoeffs - 1,)
        res = ehrlich_aberth(coeffs.reshape(-1, ncoeffs), roots_init.reshape(-1, nroots), **kw)
        return res.reshape(out_shape), axes[0]

    prim.def_impl(lambda *a, **k: jax.jit(lambda *aa: prim.bind(*aa, **k))(*a))
    prim.def_abstract_eval(_abstract)
    mlir.register_lowering(prim, _lowering, platform="cuda")
    ad.primitive_jvps[prim] = _jvp
    batching.primitive_batchers[prim] = _batch

    def poly_roots(coeffs, itmax=2000, compensated=False, custom_init=False, roots_init=None):
        ncoeffs = coeffs.shape[-1]
        out_shape = coeffs.shape[:-1] + (ncoeffs - 1,)
        flat = coeffs.reshape((-1, ncoeffs))[:, ::-1]                          # :73
        if custom_init:
            ri = roots_init.reshape((flat.shape[0], ncoeffs - 1))
        else:
            ri = jnp.zeros((flat.shape[0], ncoeffs - 1), dtype=jnp.complex128)
        return ehrlich_aberth(flat, ri, itmax=itmax, compensated=compensated,
                              custom_init=custom_init).reshape(out_shape)

    return poly_roots, ehrlich_aberth
