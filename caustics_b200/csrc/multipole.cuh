// Hexadecapole (Cassan 2017) approximation of the finite-source magnification of one image.
//
// Reference: /root/reference/src/caustics/multipole.py:8-244, which transcribes the expanded
// a_pq / Q_pq recursion.  Here the same quantities come from the Taylor inversion of the lens
// mapping: with W_k = W1^(k-1), W1 = sum_j eps_j / (z - r_j), the image displacement d(xi, eta) for
// a source offset (xi, eta) solves
//     xi + i eta = d - conj( sum_{m>=1} W_{m+1} d^m / m! ),
// order by order:  d_1 = mu0 [(xi + i eta) + conj(W2)(xi - i eta)],
//                  d_p = mu0 [conj(R_p) + conj(W2) R_p],  R = sum_{m>=2} W_{m+1} d^m / m!
// (homogeneous degree-p parts; d^m's degree-p part only needs d_1..d_{p-1}).  The area element
// F = Im(conj(d_xi) d_eta), averaged over the disk, gives
//     mu = F00 + (F20 + F02) rho^2/4 (1 - Gamma/5) + [3 (F40 + F04) + F22] rho^4/24 (1 - 11 Gamma/35).
#pragma once
#include "cplx.cuh"
#include "lens_core.cuh"

namespace cb200 {

template <int NL>
__device__ void hexadecapole_terms(const LensConst& L, cd z, double rho, double u1, double& mu0,
                                   double& dquad, double& dhex) {
  // W[k], k = 2..6:  (-1)^(k-1) (k-1)! sum_j eps_j / (z - r_j)^k   (multipole.py:212-231)
  cd W[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) W[k] = mk(0, 0);
  for (int j = 0; j < (NL == 1 ? 1 : NL); ++j) {
    const cd u = NL == 1 ? crecip(z) : crecip(z - L.r[j]);
    const double e = NL == 1 ? 1.0 : L.eps[j];
    cd pw = u;
    double fact = 1.0, sgn = 1.0;
    for (int k = 2; k <= 6; ++k) {
      pw = pw * u;
      fact *= (double)(k - 1);
      sgn = -sgn;
      W[k] = W[k] + (sgn * fact * e) * pw;
    }
  }
  const cd W2c = conj(W[2]);
  mu0 = 1.0 / (1.0 - norm2(W[2]));
  // P[m][q][k]: coefficient of xi^k eta^(q-k) in the degree-q part of d^m
  cd P[6][6][6];
  P[1][1][1] = mu0 * (mk(1, 0) + W2c);                 // d/d xi
  P[1][1][0] = mu0 * (mk(0, 1) * (mk(1, 0) - W2c));    // d/d eta
  const double invfact[6] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0};
  for (int p = 2; p <= 5; ++p) {
    cd R[6];
    for (int k = 0; k <= p; ++k) R[k] = mk(0, 0);
    for (int m = 2; m <= p; ++m) {
      // degree-p part of d^m = sum_i d_i (x) (d^(m-1))_(p-i)
      for (int k = 0; k <= p; ++k) P[m][p][k] = mk(0, 0);
      for (int i = 1; i <= p - m + 1; ++i) {
        const int q2 = p - i;
        for (int k1 = 0; k1 <= i; ++k1)
          for (int k2 = 0; k2 <= q2; ++k2)
            P[m][p][k1 + k2] = cfma(P[1][i][k1], P[m - 1][q2][k2], P[m][p][k1 + k2]);
      }
      const cd cw = invfact[m] * W[m + 1];
      for (int k = 0; k <= p; ++k) R[k] = cfma(cw, P[m][p][k], R[k]);
    }
    for (int k = 0; k <= p; ++k) P[1][p][k] = mu0 * (conj(R[k]) + W2c * R[k]);
  }
  // derivative series: zx_{k,l} = (k+1) d_{k+1,l}, zy_{k,l} = (l+1) d_{k,l+1};  F = Im(conj(zx) zy)
  auto dco = [&](int k, int l) { return P[1][k + l][k]; };   // d_{k,l}
  auto Fco = [&](int K, int Lq) {
    double acc = 0.0;
    for (int k1 = 0; k1 <= K; ++k1)
      for (int l1 = 0; l1 <= Lq; ++l1) {
        const int k2 = K - k1, l2 = Lq - l1;
        if (k1 + 1 + l1 > 5 || k2 + l2 + 1 > 5) continue;
        const cd a = (double)(k1 + 1) * dco(k1 + 1, l1);
        const cd b = (double)(l2 + 1) * dco(k2, l2 + 1);
        acc += a.re * b.im - a.im * b.re;   // Im(conj(a) b)
      }
    return acc;
  };
  const double F00 = Fco(0, 0);
  const double mu2 = 0.5 * (Fco(2, 0) + Fco(0, 2));
  const double mu4 = 3.0 * (Fco(4, 0) + Fco(0, 4)) + Fco(2, 2);
  const double Gamma = 2.0 * u1 / (3.0 - u1);
  mu0 = F00;
  dquad = 0.5 * mu2 * (1.0 - Gamma / 5.0) * rho * rho;
  dhex = mu4 / 24.0 * (1.0 - 11.0 * Gamma / 35.0) * rho * rho * rho * rho;
}

}  // namespace cb200
