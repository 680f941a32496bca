"""Differentiable (torch) form of the hexadecapole approximation, used only when gradients of `mag`
are requested; the forward/no-grad path is the CUDA gate kernel (csrc/multipole.cuh, k_gate).

Reference: /root/reference/src/caustics/multipole.py:8-244 (Cassan 2017).  Same construction as the
device code: truncated bivariate Taylor inversion of the lens map around each image
(d_1 = mu0 [(xi + i eta) + conj(W2)(xi - i eta)], d_p = mu0 [conj(R_p) + conj(W2) R_p],
R = sum_{m>=2} W_{m+1} d^m / m!), then the disk average of Im(conj(d_xi) d_eta)."""
import math

import torch

_ORD = 5


def _ser_mul(A, B):
    """product of bivariate series stored as dicts {(k, l): tensor}, truncated at total order 5"""
    C = {}
    for (k1, l1), a in A.items():
        for (k2, l2), b in B.items():
            if k1 + l1 + k2 + l2 <= _ORD:
                key = (k1 + k2, l1 + l2)
                C[key] = C[key] + a * b if key in C else a * b
    return C


def hexadecapole_terms(z, rho, u1, r, eps):
    """per-image (mu0, delta_quad, delta_hex); z any shape, r/eps lists of lens positions / masses"""
    W = {}
    for k in range(2, 7):
        acc = 0
        for rj, ej in zip(r, eps):
            acc = acc + ej / (z - rj) ** k
        W[k] = (-1) ** (k - 1) * math.factorial(k - 1) * acc
    W2c = torch.conj(W[2])
    mu0 = 1.0 / (1.0 - torch.abs(W[2]) ** 2)
    d = {(1, 0): mu0 * (1.0 + W2c), (0, 1): mu0 * 1j * (1.0 - W2c)}
    for order in range(2, _ORD + 1):
        R = {}
        pw = dict(d)
        for m in range(2, order + 1):
            pw = _ser_mul(pw, d)
            for key, v in pw.items():
                if sum(key) == order:
                    t = W[m + 1] / math.factorial(m) * v
                    R[key] = R[key] + t if key in R else t
        for k in range(order + 1):
            Rk = R.get((k, order - k))
            if Rk is not None:
                d[(k, order - k)] = mu0 * (torch.conj(Rk) + W2c * Rk)
    zx = {(k - 1, l): k * v for (k, l), v in d.items() if k >= 1}
    zy = {(k, l - 1): l * v for (k, l), v in d.items() if l >= 1}
    F = {}
    for (k1, l1), a in zx.items():
        for (k2, l2), b in zy.items():
            key = (k1 + k2, l1 + l2)
            if key in ((0, 0), (2, 0), (0, 2), (4, 0), (0, 4), (2, 2)):
                t = (torch.conj(a) * b).imag
                F[key] = F[key] + t if key in F else t
    mu2 = 0.5 * (F[(2, 0)] + F[(0, 2)])
    mu4 = 3.0 * (F[(4, 0)] + F[(0, 4)]) + F[(2, 2)]
    Gamma = 2 * u1 / (3.0 - u1)
    dq = 0.5 * mu2 * (1.0 - Gamma / 5.0) * rho**2
    dh = mu4 / 24.0 * (1.0 - 11.0 * Gamma / 35.0) * rho**4
    return F[(0, 0)], dq, dh
