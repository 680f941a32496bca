// Lens-equation device code: polynomial coefficients, lens mapping, Jacobian, image filter.
//
// Reference behaviour (paths under /root/reference/src/caustics):
//   coefficients  point_source.py:17-86 (binary), :89-1479 (triple).  The reference ships the
//                 expanded monomials (1 340 for the triple lens); a kernel forms the same polynomial
//                 from its product form (notebooks/ComplexPolynomialCoefficients.ipynb cells 2-5):
//                   H = prod(z-r_i), G = sum_j eps_j prod_{i!=j}(z-r_i), v_j = conj(r_j)-conj(w)
//                   0 = (z-w) prod_j (G - v_j H) - H sum_j eps_j prod_{i!=j} (G - v_i H)
//                 H and G depend on the lens only and are built once on the host (LensConst).
//   lens_eq       point_source.py:1536-1556      det J    point_source.py:1558-1580
//   image filter  |lens_eq(z) - w| < 1e-6, point_source.py:1704-1707
#pragma once
#include "cplx.cuh"

namespace cb200 {

// Lens constants, passed to kernels by value.  Positions r_j, mass fractions eps_j, and the
// lens-only polynomials H (degree NL, monic) and G (degree NL-1), coefficients low->high.
struct LensConst {
  int nlenses;
  double eps[3];
  cd r[3];
  cd H[4];
  cd G[3];
  double x_cm;  // centre-of-mass shift added to source positions (point_source.py:1805,1816)
};

template <int NA, int NB>
__device__ __forceinline__ void pmul(const cd (&a)[NA], const cd (&b)[NB], cd (&out)[NA + NB - 1]) {
#pragma unroll
  for (int i = 0; i < NA + NB - 1; ++i) out[i] = mk(0, 0);
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) out[i + j] = cfma(a[i], b[j], out[i + j]);
}

// Coefficients (low->high) of the degree NL^2+1 lens polynomial for source position w.
template <int NL>
__device__ __forceinline__ void lens_poly(const LensConst& L, cd w, cd (&p)[NL * NL + 2]);

template <>
__device__ __forceinline__ void lens_poly<2>(const LensConst& L, cd w, cd (&p)[6]) {
  const cd wb = conj(w);
  cd A[2][3];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const cd v = conj(L.r[j]) - wb;
#pragma unroll
    for (int k = 0; k < 3; ++k) A[j][k] = (k < 2 ? L.G[k] : mk(0, 0)) - v * L.H[k];
  }
  cd A12[5], first[6], lin[2] = {-w, mk(1, 0)};
  pmul<3, 3>(A[0], A[1], A12);
  pmul<5, 2>(A12, lin, first);
  cd S[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) S[k] = L.eps[0] * A[1][k] + L.eps[1] * A[0][k];
  cd Hc[3] = {L.H[0], L.H[1], L.H[2]}, second[5];
  pmul<3, 3>(S, Hc, second);
#pragma unroll
  for (int k = 0; k < 5; ++k) p[k] = first[k] - second[k];
  p[5] = first[5];
}

template <>
__device__ __forceinline__ void lens_poly<3>(const LensConst& L, cd w, cd (&p)[11]) {
  const cd wb = conj(w);
  cd A[3][4];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const cd v = conj(L.r[j]) - wb;
#pragma unroll
    for (int k = 0; k < 4; ++k) A[j][k] = (k < 3 ? L.G[k] : mk(0, 0)) - v * L.H[k];
  }
  cd A12[7], A13[7], A23[7];
  pmul<4, 4>(A[0], A[1], A12);
  pmul<4, 4>(A[0], A[2], A13);
  pmul<4, 4>(A[1], A[2], A23);
  cd A123[10], first[11], lin[2] = {-w, mk(1, 0)};
  pmul<7, 4>(A12, A[2], A123);
  pmul<10, 2>(A123, lin, first);
  cd S[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) S[k] = L.eps[0] * A23[k] + (L.eps[1] * A13[k] + L.eps[2] * A12[k]);
  cd Hc[4] = {L.H[0], L.H[1], L.H[2], L.H[3]}, second[10];
  pmul<7, 4>(S, Hc, second);
#pragma unroll
  for (int k = 0; k < 10; ++k) p[k] = first[k] - second[k];
  p[10] = first[10];
}

// sum_j eps_j / (conj(z) - conj(r_j))  and  sum_j eps_j / (conj(z) - conj(r_j))^2
template <int NL>
__device__ __forceinline__ void lens_sums(const LensConst& L, cd z, cd& s1, cd& s2) {
  s1 = mk(0, 0);
  s2 = mk(0, 0);
  const cd zb = conj(z);
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    const cd u = crecip(zb - conj(L.r[j]));
    s1 = s1 + L.eps[j] * u;
    s2 = s2 + L.eps[j] * (u * u);
  }
}

// Image test and signed Jacobian determinant of one candidate image z of source position w.
template <int NL>
__device__ __forceinline__ void image_eval(const LensConst& L, cd z, cd w, bool& real_image,
                                           double& detj) {
  cd s1, s2;
  lens_sums<NL>(L, z, s1, s2);
  const cd d = (z - s1) - w;
  real_image = norm2(d) < 1e-12;  // |lens_eq(z) - w| < 1e-6
  detj = 1.0 - norm2(s2);
}

}  // namespace cb200
