"""CPU tier: the C-ABI library builds, loads, and exports every symbol include/caustics_b200.h
declares (no compute calls).  Also the host-side argument checking and the loud failure without
a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "caustics_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(caustics_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from caustics_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/caustics_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in caustics_b200/_lib.py"


def test_struct_layouts(built_lib):
    from caustics_b200 import _lib
    d = _lib.EADescriptor()
    assert built_lib.caustics_ea_make_descriptor(ctypes.byref(d), 123456789012, 10, 2500, 1, 0, 3) == ctypes.sizeof(d) == 24
    assert (d.size, d.deg, d.itmax, d.compensated, d.custom_init, d.flags) == (123456789012, 10, 2500, 1, 0, 3)
    assert ctypes.sizeof(_lib.Lens) == 56
    assert ctypes.sizeof(_lib.MagPSDescriptor) == 72 and ctypes.sizeof(_lib.MagExtDescriptor) == 112


def test_argument_errors_without_compute(built_lib):
    from caustics_b200 import _lib
    L = built_lib
    assert L.caustics_ea_solve(None, None, None, None, 10, 17, 100, 0, 0, 0, None) == 2  # unsupported degree
    assert L.caustics_ea_solve(None, None, None, None, -1, 5, 100, 0, 0, 0, None) == 1
    assert L.caustics_ea_solve(None, None, None, None, 0, 5, 100, 0, 0, 0, None) == 0    # empty batch: no launch
    assert L.caustics_ea_solve(None, None, None, None, 4, 5, 100, 0, 0, 0, None) == 1    # null buffers
    for deg in range(2, 17):
        assert L.caustics_ea_degree_supported(deg) == 1
    assert L.caustics_ea_degree_supported(17) == 0 and L.caustics_ea_degree_supported(1) == 0
    bad = _lib.Lens(); bad.nlenses = 4
    assert L.caustics_mag_point_source(None, None, None, 0, ctypes.byref(bad), 10, 0, 0, None) == 1
    # bad opaque descriptor: nothing launched, sticky error instead of the reference's C++ throw
    bufs = (ctypes.c_void_p * 3)()
    L.caustics_ea_xla(None, bufs, b"xx", 2)
    assert L.caustics_last_xla_error() == 3
    for fn in (L.caustics_mag_ps_xla, L.caustics_mag_ext_xla):
        L.caustics_ea_xla(None, bufs, bytes(_lib.EADescriptor(deg=5)), 24)     # size 0: clears the sticky error
        assert L.caustics_last_xla_error() == 0
        fn(None, bufs, b"xx", 2)
        assert L.caustics_last_xla_error() == 3
    assert b"degree" in L.caustics_error_string(2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built_lib):
    import caustics_b200 as cb
    from caustics_b200._lib import CausticsError
    with pytest.raises(CausticsError):
        cb.poly_roots(np.ones((3, 6), dtype=np.complex128))
    with pytest.raises(CausticsError):
        cb.mag_point_source(np.zeros(4, dtype=np.complex128) + 0.1, nlenses=2, s=0.9, q=0.2)
    with pytest.raises(CausticsError):
        cb.mag_point_source_map(-1.5, -1.5, 1e-3, 1e-3, 10, 10, s=0.9, q=0.2)
    with pytest.raises(ValueError):          # argument errors come before the device check
        cb.mag_point_source_map(-1.5, -1.5, 1e-3, 1e-3, 10, 10, nlenses=1)


def test_product_never_imports_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "caustics_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libea_oracle" not in src and "oracle/_ref" not in src, f


def test_jvp_rule_host():
    """roots_jvp (the Python JVP rule that stays in Python) against finite differences with
    numpy.roots as the solver -- runs without a GPU."""
    import caustics_b200 as cb
    rng = np.random.default_rng(3)
    c = rng.standard_normal((20, 6)) + 1j * rng.standard_normal((20, 6))
    dc = rng.standard_normal((20, 6)) + 1j * rng.standard_normal((20, 6))
    def roots(cc, ref=None):
        out = np.array([np.roots(row[::-1]) for row in cc])
        if ref is not None:  # align with ref ordering
            out = np.array([[o[np.argmin(abs(o - r))] for r in rr] for o, rr in zip(out, ref)])
        return out
    z = roots(c)
    h = 1e-6
    fd = (roots(c + h * dc, z) - roots(c - h * dc, z)) / (2 * h)
    assert np.allclose(cb.roots_jvp(c, z, dc), fd, rtol=1e-4, atol=1e-4)
    zt = cb.roots_jvp(torch.from_numpy(c), torch.from_numpy(z), torch.from_numpy(dc))
    assert np.allclose(zt.numpy(), cb.roots_jvp(c, z, dc))


def test_round2_entry_points_without_compute(built_lib):
    """the round-2 additions: tuning overrides (no getenv anywhere in the library), the gated workspace
    sizes, and the peer-buffer entry points' argument checks / loud failure without a device"""
    L = built_lib
    assert L.caustics_set_tuning(b"path_run", 8) == 0 and L.caustics_set_tuning(b"path_run", -1) == 0
    assert L.caustics_set_tuning(b"no_such_knob", 1) == 1 and L.caustics_set_tuning(None, 1) == 1
    for key in (b"grid_run", b"path_run", b"grid_extrap", b"ext_variants", b"open_wsmall", b"host_slots", b"host_chunk_log2",
                b"ext_split", b"ext_windows"):      # every key include/caustics_b200.h documents
        assert L.caustics_set_tuning(key, 1) == 0 and L.caustics_set_tuning(key, -1) == 0, key
    for f in os.listdir(os.path.join(ROOT, "caustics_b200", "csrc")):
        assert "getenv" not in open(os.path.join(ROOT, "caustics_b200", "csrc", f)).read().replace("no getenv", ""), f
    # a gated call's workspace: per-source arrays for max_full sources + a list of n points
    full = L.caustics_ext_workspace_bytes(1_000_000, 2, 200, 1, 100)
    small = L.caustics_mag_workspace_bytes(1_000_000, 60_000, 2, 200, 1, 100)
    one = L.caustics_mag_workspace_bytes(1_000_000, 1, 2, 200, 1, 100)
    assert one < small < full and small < 4 << 30 and one >= 4_000_000
    assert L.caustics_mag_workspace_bytes(1000, 5000, 2, 200, 0, 100) == L.caustics_ext_workspace_bytes(1000, 2, 200, 4, 100)   # 4 = CAUSTICS_WS_MAG_ONLY
    assert L.caustics_mag_workspace_bytes(1000, 10, 7, 200, 0, 100) == 0          # bad nlenses
    # peer buffers
    p = ctypes.c_void_p()
    assert L.caustics_peer_alloc(None, 16) == 1 and L.caustics_peer_alloc(ctypes.byref(p), 0) == 1
    assert L.caustics_peer_export(None, None) == 1 and L.caustics_peer_open(None, None) == 1
    assert L.caustics_peer_free(None) == 0 and L.caustics_peer_close(None) == 0
    if not torch.cuda.is_available():
        assert L.caustics_peer_alloc(ctypes.byref(p), 1024) >= 1000            # 1000 + cudaError: no CPU stand-in
        lens = __import__("caustics_b200")._lib.Lens(); lens.nlenses = 2; lens.a = 0.45; lens.e1 = 0.8
        out = np.zeros(64)
        assert L.caustics_mag_point_source_grid_host(0.0, 0.0, 0.1, 0.1, 8, 0, 8, out.ctypes.data, ctypes.byref(lens),
                                                     100, 0, 0) >= 1000
    assert L.caustics_mag_point_source_grid_host(0.0, 0.0, 0.1, 0.1, 0, 0, 8, None, None, 100, 0, 0) == 1
