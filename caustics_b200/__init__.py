"""caustics_b200 -- B200 (sm_100a) implementation of the data-parallel hot path of
fbartolic/caustics behind the reference's own API names.

The compute lives in hand-written CUDA kernels behind the C ABI of include/caustics_b200.h
(libcaustics_b200.so, built in-tree by `python -m caustics_b200.build`).  There is no CPU fallback.
"""
from .primitive import poly_roots, ehrlich_aberth, roots_jvp
from .point_source import (mag_point_source, mag_point_source_map, lens_eq, lens_eq_det_jac, lens_params,
                           critical_and_caustic_curves)
from .extended_source import mag_extended_source, mag, mag_gate
from .lightcurve import AnnualParallaxTrajectory, marginalized_log_likelihood, light_curve_log_likelihood

__all__ = ["poly_roots", "ehrlich_aberth", "roots_jvp", "mag_point_source", "mag_point_source_map", "lens_eq",
           "lens_eq_det_jac", "lens_params", "mag_extended_source", "mag",
           "critical_and_caustic_curves", "mag_gate", "AnnualParallaxTrajectory",
           "marginalized_log_likelihood", "light_curve_log_likelihood"]
__version__ = "0.1.0"
