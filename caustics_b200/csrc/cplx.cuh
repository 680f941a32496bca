// Complex-double helpers for the sm_100a kernels.  Plain structs (no thrust): everything lives in
// registers, arithmetic is written out so that nvcc can contract to DFMA where contraction is
// harmless, and the error-free transformations use the _rn intrinsics so it never can.
#pragma once
#ifndef CB200_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace cb200 {

struct cd {
  double re, im;
};

__device__ __forceinline__ cd mk(double re, double im) { cd r; r.re = re; r.im = im; return r; }
__device__ __forceinline__ cd operator+(cd a, cd b) { return mk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cd operator-(cd a, cd b) { return mk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cd operator-(cd a) { return mk(-a.re, -a.im); }
// Products and sums-of-products are written with explicit fma / _rn intrinsics, so the rounding of
// every complex operation is fixed by the source and not by nvcc's contraction choices: a
// polynomial's result is then independent of which template variant its warp executes (the variants
// differ only in selects), i.e. of how the batch happens to be laid out over warps.
__device__ __forceinline__ cd operator*(cd a, cd b) {
  return mk(fma(a.re, b.re, -__dmul_rn(a.im, b.im)), fma(a.re, b.im, __dmul_rn(a.im, b.re)));
}
__device__ __forceinline__ cd operator*(double s, cd a) { return mk(s * a.re, s * a.im); }
__device__ __forceinline__ cd conj(cd a) { return mk(a.re, -a.im); }
__device__ __forceinline__ double norm2(cd a) { return fma(a.re, a.re, __dmul_rn(a.im, a.im)); }
__device__ __forceinline__ double cabs_fast(cd a) { return sqrt(norm2(a)); }
// a*b + c with all four products fused
__device__ __forceinline__ cd cfma(cd a, cd b, cd c) {
  return mk(fma(a.re, b.re, fma(-a.im, b.im, c.re)), fma(a.re, b.im, fma(a.im, b.re, c.im)));
}
// Fast double reciprocal / reciprocal square root: the 20-bit hardware seed (MUFU.RCP64H / RSQ64H via
// the PTX .approx forms) polished by two Newton steps -> <= 2 ulp, branch-free (the compiler's own
// 1.0/x and sqrt() carry a slow-path call for denormals that costs BSSY/BSYNC/BRA on every use).
// Valid for normal, finite, non-zero arguments, which is all the path produces: coefficients are
// normalised by a power of two on load, so |.|^2 neither overflows nor underflows.
#ifdef CB200_HOSTSIM
__device__ __forceinline__ double rcp_fast(double x) { return 1.0 / x; }
__device__ __forceinline__ double rsqrt_fast(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ double rcp_fast1(double x) { return 1.0 / x; }
#else
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
// one Newton step: relative error ~2^-44, for quadrature integrands
__device__ __forceinline__ double rcp_fast1(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double h = 0.5 * y;
  double e = fma(-x * y, y, 1.0);
  y = fma(h, e, y);
  h = 0.5 * y;
  e = fma(-x * y, y, 1.0);
  return fma(h, e, y);
}
#endif
// sqrt(x) for x >= 0 (exact zero allowed), branch-free
__device__ __forceinline__ double sqrt_fast(double x) {
  const double r = rsqrt_fast(x);
  return x > 0.0 ? x * r : 0.0;
}
__device__ __forceinline__ cd crecip(cd a) {
  double inv = rcp_fast(norm2(a));
  return mk(a.re * inv, -a.im * inv);
}
__device__ __forceinline__ cd cdiv(cd a, cd b) {
  double inv = rcp_fast(norm2(b));
  return mk(__dmul_rn(fma(a.re, b.re, __dmul_rn(a.im, b.im)), inv), __dmul_rn(fma(a.im, b.re, -__dmul_rn(a.re, b.im)), inv));
}
__device__ __forceinline__ cd csel(bool c, cd a, cd b) { return mk(c ? a.re : b.re, c ? a.im : b.im); }

// ---- error-free transformations (reference: lib/ehrlich_aberth/horner.h:44-84) ---------------
// _rn intrinsics are never contracted or re-associated by nvcc.
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
  double x = __dadd_rn(a, b);
  double t = __dsub_rn(x, a);
  e = __dadd_rn(__dsub_rn(a, __dsub_rn(x, t)), __dsub_rn(b, t));
  s = x;
}
__device__ __forceinline__ void two_prod(double a, double b, double& p, double& e) {
  double x = __dmul_rn(a, b);
  e = __fma_rn(a, b, -x);
  p = x;
}

}  // namespace cb200
