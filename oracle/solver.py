"""TEST INFRASTRUCTURE ONLY -- ctypes front ends to the two CPU root solvers.

``ref_solve``  calls the reference's own XLA CPU custom call ``cpu_ehrlich_aberth``
               (/root/reference/lib/ehrlich_aberth/cpu_ops.cc:15-81) through the PyCapsule its
               pybind11 module exports (pybind11_kernel_helpers.h:21-24), with exactly the operand
               list the reference's translation rule passes (ehrlich_aberth_primitive.py:196-221):
               size, deg, itmax as int64 scalars, compensated, custom_init as bool scalars,
               coeffs (size, deg+1) c128 low->high, roots_init (size, deg) c128; output flat
               (size*deg,) c128.
``port_solve`` calls oracle/ea_oracle.c (our C restatement).
"""
import ctypes
import importlib.util
import os
import glob
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lock = threading.Lock()
_ref_fn = None
_port = None


def build(quiet=True):
    """Compile libea_oracle.so and (when /root/reference exists) oracle/_ref."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def ref_available():
    return bool(glob.glob(os.path.join(_HERE, "_ref", "ehrlich_aberth_cpu_op*.so")))


def _load_ref():
    global _ref_fn
    with _lock:
        if _ref_fn is not None:
            return _ref_fn
        paths = glob.glob(os.path.join(_HERE, "_ref", "ehrlich_aberth_cpu_op*.so"))
        if not paths:
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` where "
                               "/root/reference is mounted)")
        spec = importlib.util.spec_from_file_location("ehrlich_aberth_cpu_op", paths[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        cap = mod.registrations()["cpu_ehrlich_aberth"]
        get = ctypes.pythonapi.PyCapsule_GetPointer
        get.restype = ctypes.c_void_p
        get.argtypes = [ctypes.py_object, ctypes.c_char_p]
        ptr = get(cap, b"xla._CUSTOM_CALL_TARGET")
        proto = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p))
        _ref_fn = (proto(ptr), mod)  # keep the module alive
        return _ref_fn


def _load_port():
    global _port
    with _lock:
        if _port is not None:
            return _port
        path = os.path.join(_HERE, "libea_oracle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.ea_oracle_solve.restype = ctypes.c_int
        lib.ea_oracle_solve.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _port = lib
        return lib


def _prep(coeffs, roots_init, custom_init):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.complex128)
    assert coeffs.ndim == 2
    size, deg = coeffs.shape[0], coeffs.shape[1] - 1
    if custom_init:
        roots_init = np.ascontiguousarray(roots_init, dtype=np.complex128).reshape(size, deg)
    else:
        roots_init = np.zeros((size, deg), dtype=np.complex128)
    return coeffs, roots_init, size, deg


def ref_solve(coeffs, itmax=2500, compensated=False, custom_init=False, roots_init=None):
    """Reference solver.  coeffs: (size, deg+1) complex128, LOW -> HIGH order."""
    fn, _ = _load_ref()
    coeffs, roots_init, size, deg = _prep(coeffs, roots_init, custom_init)
    out = np.empty(size * deg, dtype=np.complex128)
    scal = [np.array(size, dtype=np.int64), np.array(deg, dtype=np.int64),
            np.array(itmax, dtype=np.int64), np.array(bool(compensated), dtype=np.bool_),
            np.array(bool(custom_init), dtype=np.bool_)]
    ins = (ctypes.c_void_p * 7)(*[a.ctypes.data for a in scal], coeffs.ctypes.data,
                                 roots_init.ctypes.data)
    fn(out.ctypes.data, ins)
    return out.reshape(size, deg)


def port_solve(coeffs, itmax=2500, compensated=False, custom_init=False, roots_init=None,
               return_stats=False):
    """C restatement (ea_oracle.c).  Same contract as ``ref_solve``."""
    lib = _load_port()
    coeffs, roots_init, size, deg = _prep(coeffs, roots_init, custom_init)
    out = np.empty((size, deg), dtype=np.complex128)
    sweeps = np.zeros(size, dtype=np.int32)
    stats = np.zeros(3, dtype=np.int64)
    rc = lib.ea_oracle_solve(coeffs.ctypes.data, roots_init.ctypes.data, out.ctypes.data, size,
                             deg, itmax, int(compensated), int(custom_init), sweeps.ctypes.data,
                             stats.ctypes.data)
    if rc != 0:
        raise ValueError("ea_oracle_solve: bad arguments")
    if return_stats:
        return out, sweeps, stats
    return out


def solve(coeffs, **kw):
    """Best available CPU solver: the reference when built, else the port."""
    return ref_solve(coeffs, **kw) if ref_available() else port_solve(coeffs, **kw)


def threaded(fn, coeffs, nthreads, **kw):
    """Run ``fn`` on disjoint slices in ``nthreads`` threads (ctypes releases the GIL; both
    solvers are re-entrant: cpu_ops.cc:36-40 allocates its scratch per call)."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.complex128)
    n = coeffs.shape[0]
    bounds = np.linspace(0, n, nthreads + 1).astype(int)
    out = [None] * nthreads
    ri = kw.pop("roots_init", None)

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        out[i] = fn(coeffs[lo:hi], roots_init=None if ri is None else ri[lo:hi], **kw)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return np.concatenate(out, axis=0)
