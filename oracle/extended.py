"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's contour-integration
extended-source magnification and of the `mag` light-curve dispatcher.

Follows (paths relative to /root/reference/src/caustics):
  images_of_source_limb   extended_source.py:66-153  (+ _permute_images :34-53, utils.match_points
                          utils.py:15-40, _images_point_source_sequential point_source.py:1711-1759)
  get_segments            extended_source.py:251-309, _process_segments :220-248, _split_segment :156-216
  merge_open_segments     extended_source.py:495-667, _connection_condition :324-442,
                          _merge_two_segments :445-492, _get_segment_length :317-321
  integrate_unif / _ld    integrate.py:18-121, utils.trapz_zero_avoiding utils.py:89-99
  mag_extended_source     extended_source.py:741-904
  mag                     lightcurve.py:99-254

The reference works on zero-padded fixed-shape arrays (a jit requirement); this restatement keeps
the same decisions but represents a segment as an index range [lo, hi) of one image track and a
merged contour as an ordered list of (segment, reversed) pieces -- the representation the CUDA
kernel uses.  Where the padded formulation has observable side effects they are reproduced and
flagged "padding:" below.  The jitters the reference draws from jax.random (1e-6 on warm starts,
1e-9 on exact duplicate roots, extended_source.py:76-85,139-148) are replaced by fixed
deterministic values of the same size.

Pinned by tests/golden/ext_golden.npz (generated from the reference's own Python through
oracle/refshim.py) to rtol 1e-4 or better, and by the reference's self-contained known-answer
test for segment splitting (tests/test_extended_source.py:186-206).
"""
import numpy as np

from . import lens as _lens
from . import solver as _solver

from . import jaxprng as _prng   # the reference's own jitter stream (fixed jax.random keys), :76-85,146


def _solve_images(w, nlenses, p, z_init=None, compensated=False, itmax=2500):
    """roots (deg, n), mask (deg, n) for source points w (n,); z_init (deg, n) or None"""
    zi = None if z_init is None else np.ascontiguousarray(z_init.T)
    return _lens.images_point_source(w, nlenses, itmax, compensated, custom_init=z_init is not None,
                                     z_init=zi, **p)


def match_points(a, b):
    """utils.py:15-40: for a_i in order pick the nearest not-yet-used b (ties: lowest index)"""
    used, out = np.zeros(len(b), bool), []
    for ai in a:
        d = np.abs(b - ai)
        d[used] = np.inf
        k = int(np.argmin(d))
        used[k] = True
        out.append(k)
    return np.array(out)


def images_of_source_limb(w0, rho, nlenses=2, npts=300, niter=10, roots_itmax=2500,
                          roots_compensated=False, **p):
    """extended_source.py:66-153 -> z (deg, N'), mask, parity with rows = continuous image tracks"""
    npts_init = int(0.5 * npts)
    theta = np.concatenate([np.linspace(-np.pi, np.pi, npts_init - 1, endpoint=False), [np.pi - 1e-8]])
    w = rho * np.exp(1j * theta) + w0
    if nlenses == 1:
        z, mask = _lens.images_point_source(w, 1)
    else:
        # sequential walk, each point warm-started from the previous one (point_source.py:1711-1759);
        # NB roots_compensated is not forwarded here (extended_source.py:104-106)
        cols = [_solve_images(w[:1], nlenses, p, itmax=roots_itmax)]
        for k in range(1, len(w)):
            cols.append(_solve_images(w[k:k + 1], nlenses, p, z_init=cols[-1][0], itmax=roots_itmax))
        z = np.concatenate([c[0] for c in cols], axis=1)
        mask = np.concatenate([c[1] for c in cols], axis=1)
    parity = np.sign(_lens.lens_eq_det_jac(z, nlenses, **p))

    n = int(int(0.5 * npts) / niter)
    for _ in range(niter):
        dz = np.abs(z[:, 1:] - z[:, :-1])
        dz = np.where(mask[:, 1:] | mask[:, :-1], dz, 0.0)
        dmax = dz.max(axis=0)
        idc = np.argsort(dmax, kind="stable")[::-1][:n]
        th_new = 0.5 * (theta[idc] + theta[idc + 1])
        w_new = rho * np.exp(1j * th_new) + w0
        if nlenses == 1:
            z_new, m_new = _lens.images_point_source(w_new, 1)
        else:
            z_new, m_new = _solve_images(w_new, nlenses, p, z_init=z[:, idc] + _prng.limb_jitters(z.shape[0], n),
                                         compensated=roots_compensated, itmax=roots_itmax)
        p_new = np.sign(_lens.lens_eq_det_jac(z_new, nlenses, **p))
        theta = np.insert(theta, idc + 1, th_new)
        z = np.insert(z, idc + 1, z_new, axis=1)
        mask = np.insert(mask, idc + 1, m_new, axis=1)
        parity = np.insert(parity, idc + 1, p_new, axis=1)

    # exact duplicates get a tiny real offset (extended_source.py:139-148)
    flat = z.reshape(-1)
    _, first = np.unique(flat, return_index=True)
    dup = np.ones(flat.shape, bool)
    dup[first] = False
    if dup.any():
        z = np.where(dup.reshape(z.shape), z + _prng.duplicate_jitters(*z.shape), z)

    # order every column so that row i continues row i of the previous column (:34-53)
    carry = z[:, 0].copy()
    for k in range(z.shape[1]):
        idx = match_points(carry, z[:, k])
        z[:, k], mask[:, k], parity[:, k] = z[idx, k], mask[idx, k], parity[idx, k]
        carry = z[:, k].copy()
    return z, mask, parity


# ------------------------------------------------------------------------------------------------
def split_track(z, parity, mask, max_parts=10):
    """_split_segment (extended_source.py:156-216) as index ranges: maximal runs [lo, hi) of
    consecutive real images of one parity with no jump > 0.1 between neighbours.  The reference
    keeps at most 2*n_parts = 10 ranges per track (argwhere size, :202-203)."""
    z = np.where(mask, z, 0.0)
    par = np.where(mask, parity, 0.0)
    n = len(z)
    dzz = z[1:] - z[:-1]
    jump = dzz.real**2 + dzz.imag**2 > 0.1**2
    dpar = par[1:] - par[:-1]
    dmask = mask[1:].astype(float) - mask[:-1].astype(float)
    change = jump | (dpar != 0) | (dmask != 0)
    change = np.concatenate([[bool(mask[0])], change, [bool(mask[-1])]])
    dmask = np.concatenate([[1.0 if mask[0] else 0.0], dmask, [-1.0 if mask[-1] else 0.0]])
    starts = np.flatnonzero(change & (dmask >= 0))[:max_parts]
    ends = np.flatnonzero(change & (dmask <= 0))[:max_parts]
    out = []
    for k in range(max_parts):
        lo = starts[k] if k < len(starts) else 0
        hi = ends[k] if k < len(ends) else 0
        out.append((int(lo), int(hi)) if not (lo == 0 and hi == 0) else (0, 0))
    return out


class Segment:
    """points of one track over [lo, hi): z (len >= 2), parity of its head"""
    __slots__ = ("z", "parity")

    def __init__(self, z, parity):
        self.z, self.parity = np.asarray(z, dtype=np.complex128), float(parity)

    @property
    def empty(self):
        return len(self.z) == 0


def get_segments(z, mask, parity, nlenses=2):
    """extended_source.py:251-309 -> (closed [(track z, parity)], open [Segment] padded with empty
    segments to 3*(nlenses^2+1) entries, all_closed)"""
    nseg = 3 * (nlenses**2 + 1)
    zm = z * mask
    closed_flag = (np.abs(zm[:, 0] - zm[:, -1]) < 1e-5) & mask.all(axis=1)
    closed = [(zm[i], (parity[i] * mask[i])[0]) for i in range(len(z)) if closed_flag[i]]
    all_closed = bool(closed_flag.all())
    segs = []
    if not all_closed:
        parts = []
        for i in range(len(z)):
            if closed_flag[i]:
                parts += [None] * 10          # padding: zero rows of a closed track yield empty parts
                continue
            for lo, hi in split_track(z[i], parity[i], mask[i]):
                zz = np.where(mask[i], z[i], 0.0)[lo:hi]
                # fewer than 2 (non-zero) points: dropped (:232-233)
                parts.append(Segment(zz, (parity[i] * mask[i])[lo]) if np.count_nonzero(zz) >= 2 else None)
        # non-empty parts first, in REVERSED original order (argsort of a bool, reversed, :236)
        segs = [s for s in parts[::-1] if s is not None][:nseg]
    segs += [Segment([], 0.0)] * (nseg - len(segs))
    return closed, segs, all_closed


# ---- stitching -----------------------------------------------------------------------------------
def _seg_len(s):
    return 0.0 if s.empty else float(np.abs(np.diff(s.z)).sum())


def _head_line(s):
    """two points at the head, connection point LAST; skips a near-duplicate end vertex (:368-378)"""
    x, t = s.z, len(s.z) - 1
    g = lambda k: x[k] if 0 <= k < len(x) else 0j            # padding: reads beyond the tail see 0
    if (abs(g(1) - g(0)) > 1e-5) or (t <= 1):
        return g(1), g(0)
    return g(2), g(1)


def _tail_line(s):
    x, t = s.z, len(s.z) - 1
    if t < 0:
        return 0j, 0j
    g = lambda k: x[min(max(k, 0), len(x) - 1)] if len(x) else 0j
    if (abs(g(t) - g(t - 1)) > 1e-5) or (t <= 1):
        # lax.dynamic_slice(x, (t-1,), (2,)) clamps its start at 0 (:383)
        s0 = max(t - 1, 0)
        return x[s0], (x[s0 + 1] if s0 + 1 < len(x) else 0j)
    s0 = max(t - 2, 0)
    return x[s0], (x[s0 + 1] if s0 + 1 < len(x) else 0j)


def connection_condition(s1, s2, ctype, min_dist=1e-5, max_dist=1e-1, max_ang=60.0):
    """extended_source.py:324-442.  ctype 0 T-H, 1 H-T, 2 H-H, 3 T-T"""
    same = s1.parity * s2.parity > 0.0
    cond_parity = same if ctype in (0, 1) else not same
    l1 = _tail_line(s1) if ctype in (0, 3) else _head_line(s1)
    l2 = _head_line(s2) if ctype in (0, 2) else _tail_line(s2)
    l1 = np.array(l1, dtype=np.complex128)   # numpy semantics (nan/inf), like the reference's jnp
    l2 = np.array(l2, dtype=np.complex128)
    dist = abs(l1[1] - l2[1])
    with np.errstate(all="ignore"):
        v1 = (l1[1] - l1[0]) / abs(l1[1] - l1[0])
        v2 = (l2[1] - l2[0]) / abs(l2[1] - l2[0])
        alpha = np.arccos(v1.real * v2.real + v1.imag * v2.imag)
        c2 = (180.0 - np.rad2deg(alpha)) < max_ang
    c3 = abs(l1[1] - l2[1]) < abs(l1[0] - l2[0])
    geom = (dist < max_dist and bool(c2) and c3) or dist < min_dist
    return bool(cond_parity and geom)


def _merge_two(s1, s2, ctype):
    """extended_source.py:445-492; the merged segment keeps the parity of its new head"""
    if ctype == 0:
        return Segment(np.concatenate([s1.z, s2.z]), s1.parity)
    if ctype == 1:
        return Segment(np.concatenate([s2.z, s1.z]), s2.parity)
    if ctype == 2:
        return Segment(np.concatenate([s2.z[::-1], s1.z]), -s2.parity)
    return Segment(np.concatenate([s1.z, s2.z[::-1]]), s1.parity)


def merge_open_segments(segs, max_contours=3, max_in_contour=20):
    """extended_source.py:495-667 -> list of merged Segments (one per contour round)"""
    segs = list(segs)
    merged = []
    for _ in range(max_contours):
        # shortest non-empty first; empty ones (length 0 -> NaN) last, stable (:656-658)
        lens_ = np.array([_seg_len(s) for s in segs])
        order = np.argsort(np.where(lens_ != 0, lens_, np.nan), kind="stable")
        segs = [segs[i] for i in order]
        active, pool = segs[0], segs[1:]
        for _step in range(max_in_contour):
            if not any((not s.empty) and s.z[0] != 0 for s in pool):
                break
            end = lambda s, tail: (0j if s.empty else (s.z[-1] if tail else s.z[0]))
            a_h, a_t = end(active, False), end(active, True)
            d = np.array([[abs(a_t - end(s, False)) for s in pool],     # T-H
                          [abs(a_h - end(s, True)) for s in pool],      # H-T
                          [abs(a_h - end(s, False)) for s in pool],     # H-H
                          [abs(a_t - end(s, True)) for s in pool]])     # T-T
            best = np.argsort(d.reshape(-1), kind="stable")[:4]
            done = False
            for f in best:
                ctype, idx = divmod(int(f), len(pool))
                if connection_condition(active, pool[idx], ctype):
                    active = _merge_two(active, pool[idx], ctype)
                    pool[idx] = Segment([], 0.0)
                    done = True
                    break
            if not done:
                break   # nothing changes in the remaining scan steps
        merged.append(active)
        segs = pool
        max_in_contour -= 2
    return merged


# ---- Green's integrals -----------------------------------------------------------------------------
def integrate_unif(c):
    """1/2 closed-integral (x dy - y dx) by the trapezoid rule, integrate.py:23-27"""
    x, y = c.real, c.imag
    return float(np.trapezoid(0.5 * x, y) + np.trapezoid(-0.5 * y, x))


def _brightness(z, rho, w0, u1, nlenses, p):
    """integrate.py:29-44"""
    r = np.abs(_lens.lens_eq(z, nlenses, **p) - w0) / rho
    with np.errstate(all="ignore"):
        B = np.where(r <= 1.0, 1 + np.sqrt(np.maximum(1 - r**2, 0.0)),
                     1 - np.sqrt(np.maximum(1 - 1.0 / r**2, 0.0)))
    return 3.0 / (3.0 - u1) * (u1 * B + 1.0 - 2.0 * u1)


def _two_panel(f, a, b, rho, n1, n2):
    """two Gauss-Legendre panels split at b -/+ 2 rho (or a + |b-a|/2 for short intervals; when
    b < a that point lies outside [b, a] -- reproduced), integrate.py:56-75"""
    ad = np.abs(b - a)
    split = np.where(b > a, b - 2 * rho, b + 2 * rho)
    split = np.where(0.5 * ad <= 2 * rho, a + 0.5 * ad, split)
    out = 0.0
    for lo, hi, n in ((a, split, n1), (split, b, n2)):
        x, wgt = np.polynomial.legendre.leggauss(n)
        pts = 0.5 * (hi - lo) * x[:, None] + 0.5 * (hi + lo)
        out = out + np.sum(0.5 * (hi - lo) * f(pts) * wgt[:, None], axis=0)
    return out


def integrate_ld(c, w0, rho, u1, nlenses, p, npts=100):
    """Dominik (1998) limb-darkened Green integrals, integrate.py:47-121.  c: closed contour
    (last vertex == first)."""
    z0 = c.sum() / len(c)
    n1 = int(npts / 2)
    n2 = npts - n1
    x, y = c.real, c.imag
    P = -0.5 * _two_panel(lambda yy: _brightness(x + 1j * yy, rho, w0, u1, nlenses, p),
                          np.full_like(x, z0.imag), y, rho, n1, n2)
    Q = 0.5 * _two_panel(lambda xx: _brightness(xx + 1j * y, rho, w0, u1, nlenses, p),
                         np.full_like(y, z0.real), x, rho, n1, n2)
    return float(np.trapezoid(P, x) + np.trapezoid(Q, y))


# ------------------------------------------------------------------------------------------------
def contours(w0, rho, nlenses=2, npts_limb=150, roots_itmax=2500, roots_compensated=False, **p):
    """all closed contours [(vertices incl. closing point, parity)] of the images of the limb"""
    z, mask, parity = images_of_source_limb(w0, rho, nlenses, npts_limb, roots_itmax=roots_itmax,
                                            roots_compensated=roots_compensated, **p)
    out = []
    if nlenses == 1:
        for i in range(2):
            out.append((np.append(z[i], z[i][0]), parity[i][0]))
        return out
    closed, segs, all_closed = get_segments(z, mask, parity, nlenses)
    for zz, par in closed:
        out.append((np.append(zz, zz[0]), par))
    if not all_closed:
        for s in merge_open_segments(segs):
            if not s.empty:
                out.append((np.append(s.z, s.z[0]), s.parity))
    return out


def mag_extended_source(w0, rho, nlenses=2, npts_limb=150, limb_darkening=False, u1=0.0, npts_ld=100,
                        roots_itmax=2500, roots_compensated=False, **hp):
    """extended_source.py:741-904 (high-level parameters s, q[, q3, r3, psi])"""
    p, x_cm = _lens.lens_params(nlenses, **hp)
    w0 = complex(w0) + x_cm
    tot = 0.0
    for c, par in contours(w0, rho, nlenses, npts_limb, roots_itmax, roots_compensated, **p):
        I = integrate_ld(c, w0, rho, u1, nlenses, p, npts_ld) if limb_darkening else integrate_unif(c)
        tot += I * par
    return abs(tot) / (np.pi * rho**2)


def mag(w_points, rho, nlenses=2, npts_limb=200, limb_darkening=False, u1=0.0, npts_ld=100,
        roots_itmax=2500, roots_compensated=False, return_test=False, **hp):
    """lightcurve.py:99-254.  nlenses == 2: hexadecapole where the gate passes, else full contour
    integration; nlenses == 3: full integration everywhere (the reference leaves mu_multi
    unassigned there and cannot run)."""
    w_points = np.asarray(w_points, dtype=np.complex128)
    if nlenses == 2:
        mu, test = _lens.gate(w_points, rho, hp["s"], hp["q"], roots_itmax, roots_compensated)
    else:
        mu, test = np.zeros(w_points.shape), np.zeros(w_points.shape, bool)
    out = np.array(mu, dtype=float)
    for i in np.flatnonzero(~test):
        out[i] = mag_extended_source(w_points[i], rho, nlenses, npts_limb, limb_darkening, u1, npts_ld,
                                     roots_itmax, roots_compensated, **hp)
    return (out, test) if return_test else out
