"""Either side of the magnification in a light-curve likelihood: the source
trajectory (reference src/caustics/trajectory.py) and the flux-marginalised log-likelihood
(src/caustics/linalg.py), as stream-ordered kernels so that one likelihood evaluation --
trajectory -> mag -> likelihood -- is enqueued without a host round trip."""
import numpy as np
import torch

from . import _lib

__all__ = ["AnnualParallaxTrajectory", "marginalized_log_likelihood", "light_curve_log_likelihood"]


def _dev(x, device=None):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))
    return t.to(device=device or (t.device if t.is_cuda else "cuda"), dtype=torch.float64).contiguous()


class AnnualParallaxTrajectory:
    """trajectory.py:12-158.  The reference builds the Sun's projected position and velocity tables from
    the JPL ephemeris through astropy (`__init__(t, coords)`, :24-104); astropy is outside the hot path,
    so the tables are passed in: `t_jpl, s_e, s_n, s_e_dot, s_n_dot` (East/North components, daily
    sampling in the reference).  Without tables the motion is rectilinear (no parallax)."""

    def __init__(self, t=None, coords=None, *, t_jpl=None, s_e=None, s_n=None, s_e_dot=None, s_n_dot=None):
        if coords is not None:
            raise NotImplementedError(
                "ephemeris lookup needs astropy (trajectory.py:60-104); pass the tables "
                "t_jpl, s_e, s_n, s_e_dot, s_n_dot instead")
        self.t = t
        tabs = (t_jpl, s_e, s_n, s_e_dot, s_n_dot)
        if any(x is not None for x in tabs):
            if any(x is None for x in tabs):
                raise ValueError("t_jpl, s_e, s_n, s_e_dot and s_n_dot have to be given together")
            _lib.require_cuda()
            self.t_jpl, self.s_e, self.s_n, self.s_e_dot, self.s_n_dot = (_dev(x) for x in tabs)
            if len({x.numel() for x in (self.t_jpl, self.s_e, self.s_n, self.s_e_dot, self.s_n_dot)}) != 1:
                raise ValueError("ephemeris tables differ in length")
        else:
            self.t_jpl = None

    def compute(self, t, parametrization="cartesian", **params):
        """trajectory.py:122-158: complex source position u_e + i u_n at the times `t`."""
        if parametrization == "polar":
            psi, piE = float(params["psi"]), float(params["piE"])
        elif parametrization == "cartesian":
            piEE, piEN = float(params.get("piEE", 0.0)), float(params.get("piEN", 0.0))
            psi, piE = float(np.arctan2(piEE, piEN)), float(np.hypot(piEN, piEE))
        else:
            raise ValueError(
                "Invalid parametrization. Choose from 'polar' (piE, psi) or 'cartesian' (piEE, piEN).")
        _lib.require_cuda()
        is_t = isinstance(t, torch.Tensor)
        td = _dev(t)
        w = torch.empty(td.shape, dtype=torch.complex128, device=td.device)
        # the tables moved to t's device stay referenced until after the launch (a temporary freed before the
        # kernel is enqueued could be handed out again by the caching allocator)
        moved = [] if self.t_jpl is None else [x.to(td.device) for x in
                                               (self.t_jpl, self.s_e, self.s_n, self.s_e_dot, self.s_n_dot)]
        tabs = [None] * 5 if self.t_jpl is None else [x.data_ptr() for x in moved]
        with torch.cuda.device(td.device):
            _lib.check(_lib.lib().caustics_trajectory(
                td.data_ptr(), w.data_ptr(), td.numel(), float(params["t0"]), float(params["tE"]),
                float(params["u0"]), psi, piE, *tabs, 0 if self.t_jpl is None else self.t_jpl.numel(),
                torch.cuda.current_stream().cuda_stream))
        del moved
        if is_t:
            return w if t.is_cuda else w.cpu()
        return w.cpu().numpy()


def _loglike_device(A, fobs, c_inv):
    """one light curve -> device tensor (F_s, F_b, ll); no synchronisation"""
    A = _dev(A)
    fobs, c_inv = _dev(fobs, A.device), _dev(c_inv, A.device)
    if not (A.dim() == fobs.dim() == c_inv.dim() == 1 and A.numel() == fobs.numel() == c_inv.numel()):
        raise ValueError("A, fobs and the diagonal of C_inv have to be 1-D arrays of one length")
    out = torch.empty(3, dtype=torch.float64, device=A.device)
    with torch.cuda.device(A.device):
        _lib.check(_lib.lib().caustics_marginalized_log_likelihood(
            A.data_ptr(), fobs.data_ptr(), c_inv.data_ptr(), A.numel(), out.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
    return out


def marginalized_log_likelihood(A_list, fobs_list, C_inv_list, dense_covariance=False, Lam_sd=1e04):
    """linalg.py:11-95.  Diagonal covariance: (beta_list, ll) with beta = (F_s, F_b) per light curve and
    ll summed over the light curves.  The dense branch of the reference cannot run (`jnp.linals`,
    linalg.py:80) and is not provided."""
    if dense_covariance:
        raise NotImplementedError("dense covariance: the reference's own branch is not runnable (linalg.py:80)")
    _lib.require_cuda()
    outs = [_loglike_device(A, f, c) for A, f, c in zip(A_list, fobs_list, C_inv_list)]
    res = torch.stack(outs).cpu().numpy()          # the only synchronisation
    host = not any(isinstance(x, torch.Tensor) for x in A_list)
    betas = [r[:2].copy() if host else torch.from_numpy(r[:2].copy()) for r in res]
    return betas, float(res[:, 2].sum())


def light_curve_log_likelihood(t, fobs, c_inv, trajectory, rho, traj_params, lens_params_hl, nlenses=2,
                               parametrization="cartesian", **mag_kwargs):
    """trajectory -> `mag` -> marginalised likelihood for one light curve, enqueued back to back on the
    current stream; returns (beta, ll) after a single device->host read of three doubles.
    `traj_params`: t0, tE, u0 and the parallax parameters; `lens_params_hl`: s, q[, q3, r3, psi]."""
    from .extended_source import mag
    _lib.require_cuda()
    w = trajectory.compute(_dev(t), parametrization, **traj_params)
    A = mag(w, rho, nlenses=nlenses, **mag_kwargs, **lens_params_hl)
    out = _loglike_device(A, fobs, c_inv).cpu().numpy()
    return out[:2].copy(), float(out[2])
