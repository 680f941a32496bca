"""ctypes binding of libcaustics_b200.so (the C ABI declared in include/caustics_b200.h).

There is no CPU implementation behind this module: if the library is missing it must be built
(`python -m caustics_b200.build`), and every compute call needs a CUDA device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CAUSTICS_B200_LIB", os.path.join(_HERE, "libcaustics_b200.so"))

FLAG_INIT_BINI = 1
FLAG_COEFFS_HIGH_FIRST = 2


class CausticsError(RuntimeError):
    pass


class Lens(ctypes.Structure):
    """caustics_lens: the reference's low-level `_params` + centre-of-mass shift."""
    _fields_ = [("nlenses", ctypes.c_int32), ("reserved", ctypes.c_int32), ("a", ctypes.c_double),
                ("e1", ctypes.c_double), ("e2", ctypes.c_double), ("r3_re", ctypes.c_double),
                ("r3_im", ctypes.c_double), ("x_cm", ctypes.c_double)]


class MagPSDescriptor(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("lens", Lens), ("itmax", ctypes.c_int32), ("compensated", ctypes.c_uint8),
                ("flags", ctypes.c_uint8), ("reserved", ctypes.c_uint8 * 2)]


class MagExtDescriptor(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("lens", Lens), ("rho", ctypes.c_double), ("u1", ctypes.c_double),
                ("q", ctypes.c_double), ("workspace_bytes", ctypes.c_uint64), ("npts_limb", ctypes.c_int32),
                ("npts_ld", ctypes.c_int32), ("itmax", ctypes.c_int32), ("limb_darkening", ctypes.c_uint8),
                ("compensated", ctypes.c_uint8), ("gate", ctypes.c_uint8), ("reserved", ctypes.c_uint8)]


class EADescriptor(ctypes.Structure):
    _fields_ = [("size", ctypes.c_int64), ("deg", ctypes.c_int32), ("itmax", ctypes.c_int32),
                ("compensated", ctypes.c_uint8), ("custom_init", ctypes.c_uint8),
                ("flags", ctypes.c_uint8), ("reserved", ctypes.c_uint8), ("pad", ctypes.c_int32)]


_vp, _i, _i64, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
_LP = ctypes.POINTER(Lens)

# name -> (restype, argtypes): every symbol include/caustics_b200.h declares
SIGNATURES = {
    "caustics_version": (ctypes.c_char_p, []),
    "caustics_device_count": (_i, []),
    "caustics_ea_degree_supported": (_i, [_i]),
    "caustics_error_string": (ctypes.c_char_p, [_i]),
    "caustics_ea_solve": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "caustics_ea_jvp": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "caustics_ea_vjp": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "caustics_ea_solve_host": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i]),
    "caustics_release_workspace": (None, []),
    "caustics_ea_xla": (None, [_vp, ctypes.POINTER(_vp), ctypes.c_char_p, ctypes.c_size_t]),
    "caustics_mag_ps_xla": (None, [_vp, ctypes.POINTER(_vp), ctypes.c_char_p, ctypes.c_size_t]),
    "caustics_mag_ext_xla": (None, [_vp, ctypes.POINTER(_vp), ctypes.c_char_p, ctypes.c_size_t]),
    "caustics_last_xla_error": (_i, []),
    "caustics_ea_make_descriptor": (ctypes.c_size_t, [ctypes.POINTER(EADescriptor), _i64, _i, _i, _i, _i, _i]),
    "caustics_images_point_source": (_i, [_vp, _vp, _vp, _vp, _i64, _LP, _i, _i, _i, _i, _vp]),
    "caustics_images_point_source_sequential": (_i, [_vp, _vp, _vp, _i64, _i64, _LP, _i, _i, _vp]),
    "caustics_trajectory": (_i, [_vp, _vp, _i64, _d, _d, _d, _d, _d, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "caustics_marginalized_log_likelihood": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "caustics_mag_point_source": (_i, [_vp, _vp, _vp, _i64, _LP, _i, _i, _i, _vp]),
    "caustics_mag_point_source_grid": (_i, [_d, _d, _d, _d, _i64, _i64, _i64, _vp, _LP, _i, _i, _i, _vp]),
    "caustics_mag_point_source_host": (_i, [_vp, _vp, _i64, _LP, _i, _i, _i]),
    "caustics_mag_point_source_grid_host": (_i, [_d, _d, _d, _d, _i64, _i64, _i64, _vp, _LP, _i, _i, _i]),
    "caustics_peer_alloc": (_i, [ctypes.POINTER(_vp), ctypes.c_size_t]),
    "caustics_peer_free": (_i, [_vp]),
    "caustics_peer_export": (_i, [_vp, _vp]),
    "caustics_peer_open": (_i, [_vp, ctypes.POINTER(_vp)]),
    "caustics_peer_close": (_i, [_vp]),
    "caustics_peer_enable": (_i, [_i]),
    "caustics_set_tuning": (_i, [ctypes.c_char_p, _i]),
    "caustics_mag_workspace_bytes": (ctypes.c_size_t, [_i64, _i64, _i, _i, _i, _i]),
    "caustics_mag_extended_source_list": (_i, [_vp, _vp, _vp, _vp, _i64, _d, _LP, _i, _i, _d, _i, _i, _i, _vp,
                                               ctypes.c_size_t, _vp]),
    "caustics_mag_extended_source_grad": (_i, [_vp, _vp, _vp, _i64, _d, _LP, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "caustics_bench_fp64_peak": (_i, [_vp, _i, _i, _vp]),
    "caustics_bench_fp64_peak3": (_i, [_vp, _i, _i, _vp]),
    "caustics_mag_gate": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _d, _LP, _d, _i, _i, _vp]),
    "caustics_ext_contour_capacity": (_i, [_i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "caustics_ext_contours": (_i, [_vp, _vp, _i64, _d, _LP, _i, _i, _i, _vp, ctypes.c_size_t,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "caustics_match_tracks": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "caustics_ext_workspace_bytes": (ctypes.c_size_t, [_i64, _i, _i, _i, _i]),
    "caustics_mag_extended_source": (_i, [_vp, _vp, _i64, _d, _LP, _i, _i, _d, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "caustics_mag": (_i, [_vp, _vp, _vp, _i64, _d, _LP, _d, _i, _i, _d, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CausticsError(
                f"{LIB_PATH} is not built; run `python -m caustics_b200.build` (needs nvcc). "
                "caustics_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise CausticsError(f"caustics_b200 error {rc}: {lib().caustics_error_string(rc).decode()}")


def require_cuda():
    if lib().caustics_device_count() < 1:
        raise CausticsError("no CUDA device: caustics_b200 runs only on a GPU (no CPU fallback)")
