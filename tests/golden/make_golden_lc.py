"""Golden vectors for the trajectory and the marginalised likelihood from the REFERENCE's own code
-> tests/golden/lc_golden.npz.  Build container only:  python tests/golden/make_golden_lc.py
  * linalg.marginalized_log_likelihood: imported under oracle/refshim.py (NumPy stand-in for jax).
  * trajectory.AnnualParallaxTrajectory: the module imports astropy (absent), so the two methods on the
    path (trajectory.py:106-158) are executed from the module's source text with a namespace `self`
    holding synthetic ephemeris tables."""
import ast
import os
import sys
import textwrap
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
warnings.filterwarnings("ignore")
REF = "/root/reference/src/caustics"

out = {}
rng = np.random.default_rng(7)

# ---- trajectory ----
src = open(os.path.join(REF, "trajectory.py")).read()
cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "AnnualParallaxTrajectory")
ns = {"jnp": np, "np": np}
for fn in cls.body:
    if isinstance(fn, ast.FunctionDef) and fn.name in ("_compute_delta_sun_position_and_velocity", "compute"):
        exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
t = np.sort(rng.uniform(2458000.0, 2458400.0, 300))
t_jpl = np.arange(t[0], t[-1] + 1, 1)                      # trajectory.py:36
ph = 2 * np.pi * (t_jpl - t_jpl[0]) / 365.25
s_e, s_n = 0.9 * np.cos(ph + 0.3), 0.4 * np.sin(ph + 1.1)  # synthetic projected Sun position (au)
s_e_dot, s_n_dot = np.gradient(s_e, t_jpl), np.gradient(s_n, t_jpl)
self = types.SimpleNamespace(t_jpl=t_jpl, s_e=s_e, s_n=s_n, s_e_dot=s_e_dot, s_n_dot=s_n_dot)
self._compute_delta_sun_position_and_velocity = lambda tt, t0: ns["_compute_delta_sun_position_and_velocity"](self, tt, t0)
tp = dict(t0=2458210.3, tE=35.0, u0=0.12, piEE=0.21, piEN=-0.13)
out.update(traj_t=t, traj_t_jpl=t_jpl, traj_s_e=s_e, traj_s_n=s_n, traj_s_e_dot=s_e_dot, traj_s_n_dot=s_n_dot)
out["traj_params"] = np.array([tp[k] for k in ("t0", "tE", "u0", "piEE", "piEN")])
out["traj_w_cartesian"] = np.asarray(ns["compute"](self, t, "cartesian", **dict(tp)))
pol = dict(t0=tp["t0"], tE=tp["tE"], u0=tp["u0"], psi=0.7, piE=0.3)
out["traj_w_polar"] = np.asarray(ns["compute"](self, t, "polar", **dict(pol)))

# ---- likelihood ----
from oracle import refshim  # noqa: E402
refshim.install()
A_ = refshim.arr
try:
    from caustics import linalg as LA
    mll = LA.marginalized_log_likelihood
    mll = getattr(mll, "__wrapped__", mll)
except Exception as e:  # pragma: no cover
    raise SystemExit(f"cannot import the reference's linalg under the shim: {e!r}")
As, fs, cs = [], [], []
for n in (300, 1200):
    A = 1 + 5 * np.exp(-0.5 * (np.linspace(-3, 3, n)) ** 2) * (1 + 0.2 * np.sin(np.linspace(0, 40, n)))
    sig = 0.02 * (1 + rng.uniform(0, 1, n))
    f = 3.7 * A + 1.9 + sig * rng.standard_normal(n)
    As.append(A); fs.append(f); cs.append(1 / sig**2)
betas, ll = mll([A_(x) for x in As], [A_(x) for x in fs], [A_(x) for x in cs], dense_covariance=False)
for i in range(2):
    out[f"ll_A{i}"], out[f"ll_f{i}"], out[f"ll_c{i}"] = As[i], fs[i], cs[i]
    out[f"ll_beta{i}"] = np.asarray(A_(betas[i]), dtype=np.float64).reshape(-1)
out["ll_total"] = np.float64(np.asarray(A_(ll)).real)
print("ll", out["ll_total"], "betas", out["ll_beta0"], out["ll_beta1"])
np.savez_compressed(os.path.join(HERE, "lc_golden.npz"), **out)
