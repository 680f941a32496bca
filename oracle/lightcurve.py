"""TEST INFRASTRUCTURE (CPU oracle): NumPy restatement of the reference's source trajectory and
flux-marginalised likelihood.  Pinned against the reference's own code by tests/golden/make_golden_lc.py
(linalg.py imported under oracle/refshim.py; trajectory.py cannot be imported -- it needs astropy at
import time -- so the two methods on the path are executed from their source text).
Never imported by the product."""
import numpy as np


def delta_sun(t, t0, t_jpl, s_e, s_n, s_e_dot, s_n_dot):
    """trajectory.py:106-120 _compute_delta_sun_position_and_velocity"""
    s_e_t, s_n_t = np.interp(t, t_jpl, s_e), np.interp(t, t_jpl, s_n)
    s_e_t0, s_n_t0 = np.interp(t0, t_jpl, s_e), np.interp(t0, t_jpl, s_n)
    s_e_dot_t0, s_n_dot_t0 = np.interp(t0, t_jpl, s_e_dot), np.interp(t0, t_jpl, s_n_dot)
    return s_e_t - s_e_t0 - (t - t0) * s_e_dot_t0, s_n_t - s_n_t0 - (t - t0) * s_n_dot_t0


def trajectory(t, tables=None, parametrization="cartesian", **params):
    """trajectory.py:122-158 AnnualParallaxTrajectory.compute; tables = (t_jpl, s_e, s_n, s_e_dot, s_n_dot)
    or None for rectilinear motion"""
    t = np.asarray(t, dtype=np.float64)
    if parametrization == "polar":
        psi, piE = params["psi"], params["piE"]
    elif parametrization == "cartesian":
        psi = np.arctan2(params["piEE"], params["piEN"])
        piE = np.sqrt(params["piEN"] ** 2 + params["piEE"] ** 2)
    else:
        raise ValueError("Invalid parametrization.")
    de, dn = (0.0, 0.0) if tables is None else delta_sun(t, params["t0"], *tables)
    tau = (t - params["t0"]) / params["tE"]
    u_e = params["u0"] * np.cos(psi) + tau * np.sin(psi) + piE * de
    u_n = -params["u0"] * np.sin(psi) + tau * np.cos(psi) + piE * dn
    return u_e + 1j * u_n


def marginalized_log_likelihood(A_list, fobs_list, C_inv_list):
    """linalg.py:55-70, dense_covariance=False.  Like the reference it materialises diag(C_inv)
    (n x n doubles): keep n to a few thousand."""
    ll, betas = 0.0, []
    for A, fobs, C_inv in zip(A_list, fobs_list, C_inv_list):
        if len(A) > 20000:
            raise MemoryError("oracle restatement builds an n x n matrix; use n <= 20000")
        M = np.stack([A, np.ones_like(A)]).T
        MTCinvM = M.T @ np.diag(C_inv) @ M
        Sigma = np.linalg.solve(MTCinvM, np.eye(2))
        beta = Sigma @ M.T @ np.diag(C_inv) @ fobs[:, None]
        fpred = (M @ beta).reshape(-1)
        ll += -0.5 * np.sum((fobs - fpred) ** 2 * C_inv) + 0.5 * np.log(np.linalg.det(2 * np.pi * Sigma))
        betas.append(beta.reshape(-1))
    return betas, ll
