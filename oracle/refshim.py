"""TEST INFRASTRUCTURE ONLY -- runs the reference's OWN Python (unmodified, imported from
/root/reference/src) in this container, where jax is not installable.

How: a small NumPy stand-in for the handful of jax entry points the reference uses is installed in
``sys.modules`` as ``jax`` (eager semantics: ``jit`` is the identity, ``lax.scan``/``lax.map`` are
Python loops, ``lax.cond``/``switch`` are Python branches, ``vmap`` loops over the leading axis),
and ``caustics.ehrlich_aberth_primitive`` -- the only module that needs jax *internals* -- is
replaced by a stub whose ``poly_roots`` has the reference's signature and semantics
(ehrlich_aberth_primitive.py:34-94) and calls the reference's own C++ solver through oracle/_ref.
Everything else (point_source.py, extended_source.py, integrate.py, multipole.py, lightcurve.py,
utils.py) is the reference's code, byte for byte.

Differences from real jax that can matter, all replicated or documented:
  * out-of-range integer indexing is clamped like XLA gather (utils.py:93-94 reads x[tail+1]);
  * ``lax.dynamic_slice`` clamps its start index;
  * argsort is stable (XLA's sort is);
  * ``jax.random`` IS JAX's default threefry2x32 generator (oracle/jaxprng.py, float64 as under
    jax_enable_x64): the reference's jitters of warm starts and duplicate roots come from fixed keys
    (extended_source.py:76-85,146) and decide which limb intervals get refined at caustic crossings, so
    they are reproduced exactly rather than stood in for.

This module only works where /root/reference is mounted; it exists to generate tests/golden/*
(see tests/golden/make_golden.py) and to pin oracle/lens.py and oracle/extended.py.
"""
import importlib
import os
import sys
import types

import numpy as np
import scipy.special

from . import jaxprng as _prng

REF_SRC = "/root/reference/src"


class FArr(np.ndarray):
    """ndarray with jax's ``.at[idx].set()`` and clamped integer indexing."""

    def __array_finalize__(self, obj):
        pass

    @property
    def at(self):
        return _At(self)

    def _clamp(self, idx):
        def c1(i, n):
            if isinstance(i, (int, np.integer)) and not isinstance(i, (bool, np.bool_)):
                i = int(i)
                if i >= n:
                    return n - 1
                if i < -n:
                    return 0
            elif isinstance(i, np.ndarray) and i.ndim == 0 and np.issubdtype(i.dtype, np.integer):
                return c1(int(i), n)
            return i

        if isinstance(idx, tuple):
            if any(x is None or x is Ellipsis for x in idx):
                return idx
            out, ax = [], 0
            for x in idx:
                out.append(c1(x, self.shape[ax]) if ax < self.ndim else x)
                ax += 1
            return tuple(out)
        if self.ndim >= 1:
            return c1(idx, self.shape[0])
        return idx

    def __getitem__(self, idx):
        r = super().__getitem__(self._clamp(idx))
        return r

    def __iter__(self):
        for i in range(self.shape[0]):
            yield super().__getitem__(i)

    def argsort(self, axis=-1, kind=None, order=None, **kw):
        return np.asarray(self).argsort(axis=axis, kind="stable").view(FArr)

    def astype(self, dtype, *a, **k):
        return np.asarray(self).astype(dtype, *a, **k).view(FArr)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, val, *ignored, **kw):
        out = np.array(self.arr, copy=True)
        idx = self.idx
        # jax drops out-of-bounds scatter updates
        if isinstance(idx, (int, np.integer)) or (isinstance(idx, np.ndarray) and idx.ndim == 0):
            if int(idx) >= out.shape[0] or int(idx) < -out.shape[0]:
                return out.view(FArr)
        out[idx] = val
        return out.view(FArr)


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, FArr):
        return x.view(FArr)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    if isinstance(x, np.generic):
        return np.asarray(x).view(FArr)
    return x


def _unwrap(x):
    if isinstance(x, FArr):
        return np.asarray(x)
    if isinstance(x, (tuple, list)):
        return type(x)(_unwrap(v) for v in x)
    return x


class _JnpModule(types.ModuleType):
    """jax.numpy stand-in: numpy, with results wrapped in FArr and a few jax-only kwargs."""

    def __getattr__(self, name):
        over = _JNP_OVERRIDES.get(name)
        if over is not None:
            return over
        obj = getattr(np, name)
        if callable(obj) and not isinstance(obj, type):
            def f(*a, **k):
                return _wrap(obj(*_unwrap(a), **{kk: _unwrap(v) for kk, v in k.items()}))
            f.__name__ = name
            return f
        return obj


def _argsort(a, axis=-1, **kw):
    return _wrap(np.argsort(np.asarray(a), axis=axis, kind="stable"))


def _argwhere(a, size=None, fill_value=0):
    r = np.argwhere(np.asarray(a))
    if size is not None:
        out = np.full((size, r.shape[1]), fill_value, dtype=r.dtype)
        out[: min(size, len(r))] = r[:size]
        r = out
    return _wrap(r)


def _unique(a, return_index=False, size=None, **kw):
    r = np.unique(np.asarray(a), return_index=return_index)
    return _wrap(r)


def _trapz(y, x=None, dx=1.0, axis=-1):
    return _wrap(np.trapezoid(np.asarray(y), x=None if x is None else np.asarray(x), dx=dx, axis=axis))


def _isin(a, b, assume_unique=False, **kw):
    return _wrap(np.isin(np.asarray(a), np.asarray(b)))


def _array(x, dtype=None, **kw):
    return _wrap(np.array(_unwrap(x), dtype=dtype))


_JNP_OVERRIDES = {
    "argsort": _argsort, "argwhere": _argwhere, "unique": _unique, "trapz": _trapz,
    "isin": _isin, "array": _array, "bool_": np.bool_, "complex128": np.complex128,
    "float64": np.float64, "int64": np.int64, "ndarray": np.ndarray, "pi": np.pi,
}


# ---- tree helpers for scan / vmap -------------------------------------------------------------
def _tree_index(xs, i):
    if isinstance(xs, (tuple, list)):
        return type(xs)(_tree_index(x, i) for x in xs)
    return _wrap(np.asarray(xs)[i])


def _tree_len(xs):
    if isinstance(xs, (tuple, list)):
        return _tree_len(xs[0])
    return len(xs)


def _tree_stack(ys):
    y0 = ys[0]
    if isinstance(y0, (tuple, list)):
        return type(y0)(_tree_stack([y[k] for y in ys]) for k in range(len(y0)))
    return _wrap(np.stack([np.asarray(y) for y in ys]))


def _scan(f, init, xs, length=None):
    carry, ys = init, []
    n = _tree_len(xs) if xs is not None else length
    for i in range(n):
        carry, y = f(carry, _tree_index(xs, i) if xs is not None else None)
        ys.append(y)
    return carry, _tree_stack(ys)


def _cond(pred, true_fn, false_fn, *operands):
    return true_fn(*operands) if bool(pred) else false_fn(*operands)


def _switch(index, branches, *operands):
    i = int(np.clip(int(index), 0, len(branches) - 1))
    return branches[i](*operands)


def _dynamic_slice(x, starts, sizes):
    x = np.asarray(x)
    sl = []
    for s, n, dim in zip(starts, sizes, x.shape):
        s = int(np.clip(int(s), 0, dim - n))
        sl.append(slice(s, s + n))
    return _wrap(x[tuple(sl)])


def _lax_map(f, xs):
    n = _tree_len(xs)
    return _tree_stack([f(_tree_index(xs, i)) for i in range(n)])


def _vmap(f, in_axes=0, out_axes=0):
    def g(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = np.asarray(a).shape[ax]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else _wrap(np.take(np.asarray(a), i, axis=ax))
                  for a, ax in zip(args, axes)]
            outs.append(f(*sl))
        return _tree_stack(outs)
    return g


def _jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


# jax.random: JAX's default threefry2x32 generator restated in oracle/jaxprng.py (pinned to the Random123 and
# JAX known answers), so the reference draws exactly the jitters it draws under real JAX with x64 enabled
def _prng_key(seed):
    return _prng.prng_key(seed)


def _split(key, num=2):
    return tuple(_prng.split(np.asarray(key, dtype=np.uint32), num))


def _uniform(key, shape=(), dtype=np.float64, minval=0.0, maxval=1.0):
    return _wrap(_prng.uniform(np.asarray(key, dtype=np.uint32), tuple(shape), np.float64, minval, maxval))


def install():
    """Install the stand-in jax and import the reference package; returns the module."""
    if "caustics" in sys.modules and getattr(sys.modules["caustics"], "_is_refshim", False):
        return sys.modules["caustics"]
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("reference tree not mounted; refshim only works in the build container")
    from . import solver

    jax = types.ModuleType("jax")
    jnp = _JnpModule("jax.numpy")
    lax = types.ModuleType("jax.lax")
    lax.scan, lax.cond, lax.switch, lax.map = _scan, _cond, _switch, _lax_map
    lax.dynamic_slice = _dynamic_slice
    rnd = types.ModuleType("jax.random")
    rnd.PRNGKey, rnd.split, rnd.uniform = _prng_key, _split, _uniform
    jsp = types.ModuleType("jax.scipy")
    jsps = types.ModuleType("jax.scipy.special")
    jsps.gammaln = lambda x: _wrap(np.asarray(scipy.special.gammaln(np.asarray(x))))
    jsp.special = jsps
    cfgm = types.ModuleType("jax.config")

    class _Cfg:
        def update(self, *a, **k):
            pass
    cfgm.config = _Cfg()
    jax.numpy, jax.lax, jax.random, jax.scipy, jax.config = jnp, lax, rnd, jsp, cfgm.config
    jax.jit, jax.vmap = _jit, _vmap
    for name, mod in [("jax", jax), ("jax.numpy", jnp), ("jax.lax", lax), ("jax.random", rnd),
                      ("jax.scipy", jsp), ("jax.scipy.special", jsps), ("jax.config", cfgm)]:
        sys.modules[name] = mod

    # package skeleton so that the reference's relative imports resolve without running its
    # __init__ (which is what pulls in the jax-internals-only primitive module first)
    pkg = types.ModuleType("caustics")
    pkg.__path__ = [os.path.join(REF_SRC, "caustics")]
    pkg._is_refshim = True
    sys.modules["caustics"] = pkg

    prim = types.ModuleType("caustics.ehrlich_aberth_primitive")

    def poly_roots(coeffs, itmax=2000, compensated=False, custom_init=False, roots_init=None):
        # shape plumbing of ehrlich_aberth_primitive.py:66-94, solver = reference C++ (oracle/_ref)
        coeffs = np.asarray(coeffs, dtype=np.complex128)
        ncoeffs = coeffs.shape[-1]
        out_shape = coeffs.shape[:-1] + (ncoeffs - 1,)
        flat = coeffs.reshape(-1, ncoeffs)[:, ::-1]
        ri = None
        if custom_init:
            ri = np.asarray(roots_init, dtype=np.complex128).reshape(flat.shape[0], ncoeffs - 1)
        r = solver.ref_solve(flat, itmax=itmax, compensated=compensated, custom_init=custom_init,
                             roots_init=ri)
        return _wrap(r.reshape(out_shape))

    prim.poly_roots = poly_roots
    sys.modules["caustics.ehrlich_aberth_primitive"] = prim

    for sub in ["utils", "point_source", "integrate", "extended_source", "multipole"]:
        m = importlib.import_module("caustics." + sub)
        setattr(pkg, sub, m)
    pkg.lens_eq = pkg.point_source.lens_eq
    pkg.mag_point_source = pkg.point_source.mag_point_source
    pkg.critical_and_caustic_curves = pkg.point_source.critical_and_caustic_curves
    pkg.mag_extended_source = pkg.extended_source.mag_extended_source
    m = importlib.import_module("caustics.lightcurve")
    pkg.lightcurve = m
    pkg.mag = m.mag
    return pkg


def arr(x):
    """Wrap an input for the reference functions (they expect jax-like arrays)."""
    return _wrap(np.array(x))
