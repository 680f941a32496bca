"""CPU tier: caustics_b200/jax_glue.py executed against a STAND-IN for jax (jax is not installable in this
image).  The stand-in records what the glue registers and lets a NumPy root finder play the
device, so this checks the glue's own logic -- capsule, descriptor bytes, shape plumbing, coefficient flip,
JVP and batching rules -- not JAX and not the kernels (the `-m gpu` tier covers those through the same C ABI)."""
import ctypes
import sys
import types
from collections import namedtuple

import numpy as np
import pytest


class _Zero:
    pass


class _Primitive:
    backend = None          # set by the test: (coeffs_low_high, roots_init, **params) -> flat roots

    def __init__(self, name):
        self.name, self.impl, self.abstract = name, None, None

    def def_impl(self, f):
        self.impl = f

    def def_abstract_eval(self, f):
        self.abstract = f

    def bind(self, *args, **kw):
        out = _Primitive.backend(*args, **kw)
        aval = self.abstract(*[namedtuple("A", "shape dtype")(a.shape, a.dtype) for a in args], **kw)
        assert out.shape == aval.shape and out.dtype == aval.dtype      # abstract eval agrees with the "device"
        return out


@pytest.fixture
def fake_jax(monkeypatch):
    rec = {"targets": [], "lowerings": []}
    jax = types.ModuleType("jax")
    jax.ffi = types.SimpleNamespace(
        register_ffi_target=lambda name, capsule, platform, api_version: rec["targets"].append((name, capsule, platform, api_version)),
        ffi_lowering=lambda target, backend_config, api_version=4: (lambda ctx, *ops: ("custom_call", target, backend_config, len(ops), api_version)))
    jax.jit = lambda f: f
    jnp = types.ModuleType("jax.numpy")
    for n in ("zeros", "zeros_like", "complex128"):
        setattr(jnp, n, getattr(np, n))
    core = types.ModuleType("jax.core")
    core.Primitive = _Primitive
    core.ShapedArray = namedtuple("ShapedArray", "shape dtype")
    interp = types.ModuleType("jax.interpreters")
    ad = types.SimpleNamespace(Zero=_Zero, primitive_jvps={})
    batching = types.SimpleNamespace(primitive_batchers={})
    mlir = types.SimpleNamespace(register_lowering=lambda prim, rule, platform: rec["lowerings"].append((prim, rule, platform)))
    interp.ad, interp.batching, interp.mlir = ad, batching, mlir
    jax.numpy, jax.core, jax.interpreters, jax.lax, jax.vmap = jnp, core, interp, types.SimpleNamespace(), None
    for name, mod in (("jax", jax), ("jax.numpy", jnp), ("jax.core", core), ("jax.interpreters", interp)):
        monkeypatch.setitem(sys.modules, name, mod)
    rec["ad"], rec["batching"] = ad, batching
    return rec


def _np_backend(coeffs, roots_init, itmax=None, compensated=None, custom_init=False):
    # the primitive's contract: coeffs (size, deg+1) LOW -> HIGH, result flat (size * deg,)
    return np.concatenate([np.roots(row[::-1]) for row in coeffs]).astype(np.complex128)


def test_register_and_shapes(fake_jax, built_lib):
    from caustics_b200 import jax_glue, _lib
    _Primitive.backend = staticmethod(_np_backend)
    poly_roots, ehrlich_aberth = jax_glue.register()
    # one target, legacy custom-call signature, named capsule around the exported symbol
    (name, capsule, platform, api), = fake_jax["targets"]
    assert (name, platform, api) == ("caustics_b200_ehrlich_aberth", "CUDA", 0)
    get = ctypes.pythonapi.PyCapsule_GetPointer
    get.restype, get.argtypes = ctypes.c_void_p, [ctypes.py_object, ctypes.c_char_p]
    assert get(capsule, b"xla._CUSTOM_CALL_TARGET") == ctypes.cast(built_lib.caustics_ea_xla, ctypes.c_void_p).value
    # poly_roots: any leading shape, coefficients HIGH -> LOW like the reference (ehrlich_aberth_primitive.py:34-96)
    rng = np.random.default_rng(0)
    c = rng.standard_normal((3, 4, 6)) + 1j * rng.standard_normal((3, 4, 6))
    z = poly_roots(c)
    assert z.shape == (3, 4, 5)
    res = np.abs(sum(c[..., k:k + 1] * z ** (5 - k) for k in range(6)))
    assert res.max() < 1e-9
    # lowering rule: descriptor bytes = caustics_ea_descriptor (include/caustics_b200.h), complex128 only
    (prim, rule, plat), = fake_jax["lowerings"]
    assert plat == "cuda"
    ctx = types.SimpleNamespace(avals_in=[types.SimpleNamespace(shape=(12, 6), dtype=np.dtype(np.complex128))])
    kind, target, opaque, nops, api = rule(ctx, "coeffs", "roots_init", itmax=2500, compensated=True, custom_init=False)
    assert api == 1 and isinstance(opaque, bytes)        # legacy custom call with an opaque byte string
    assert (kind, target, nops) == ("custom_call", "caustics_b200_ehrlich_aberth", 2) and len(opaque) == ctypes.sizeof(_lib.EADescriptor) == 24
    d = _lib.EADescriptor.from_buffer_copy(opaque)
    assert (d.size, d.deg, d.itmax, d.compensated, d.custom_init, d.flags) == (12, 5, 2500, 1, 0, 0)
    ctx.avals_in[0].dtype = np.dtype(np.complex64)
    with pytest.raises(NotImplementedError):
        rule(ctx, "coeffs", "roots_init", itmax=2500, compensated=False, custom_init=False)


def test_jvp_and_batching_rules(fake_jax, built_lib):
    from caustics_b200 import jax_glue, roots_jvp
    _Primitive.backend = staticmethod(_np_backend)
    poly_roots, ehrlich_aberth = jax_glue.register()
    prim = fake_jax["lowerings"][0][0]
    rng = np.random.default_rng(1)
    p = rng.standard_normal((7, 6)) + 1j * rng.standard_normal((7, 6))          # low -> high
    dp = rng.standard_normal((7, 6)) + 1j * rng.standard_normal((7, 6))
    ri = np.zeros((7, 5), complex)
    z, dz = fake_jax["ad"].primitive_jvps[prim]((p, ri), (dp, _Zero()), itmax=2500, compensated=False, custom_init=False)
    assert z.shape == dz.shape == (35,)
    assert np.allclose(dz.reshape(7, 5), roots_jvp(p, z.reshape(7, 5), dp), rtol=1e-12)
    # against finite differences of the "device" (ehrlich_aberth_primitive.py:254-324, tests/test_ehrlich_aberth_primitive.py:59-64)
    h = 1e-7
    zp = np.concatenate([np.roots(r[::-1]) for r in p + h * dp]).reshape(7, 5)
    z0 = z.reshape(7, 5)
    for n in range(7):                                   # np.roots order may differ between calls: match nearest
        fd = np.array([zp[n][np.argmin(np.abs(zp[n] - x))] - x for x in z0[n]]) / h
        assert np.allclose(fd, dz.reshape(7, 5)[n], rtol=1e-4, atol=1e-5)
    # symbolic-zero tangent
    z2, dz2 = fake_jax["ad"].primitive_jvps[prim]((p, ri), (_Zero(), _Zero()), itmax=2500, compensated=False, custom_init=False)
    assert np.all(dz2 == 0)
    # batching rule (ehrlich_aberth_primitive.py:330-353): extra leading axes are flattened and restored
    pb = np.stack([p, 2 * p, p[::-1]])
    out, axis = fake_jax["batching"].primitive_batchers[prim]((pb, np.zeros((3, 7, 5), complex)), (0, 0), itmax=2500, compensated=False, custom_init=False)
    assert out.shape == (3, 7, 5) and axis == 0
