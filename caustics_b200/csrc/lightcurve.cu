// Either side of the magnification in a light-curve likelihood: the source
// trajectory that produces w_points and the flux-marginalised Gaussian log-likelihood that consumes
// the magnifications.  Both are stream-ordered with no host round trip, so
//   caustics_trajectory -> caustics_mag -> caustics_marginalized_log_likelihood
// is one enqueue per likelihood evaluation; the host reads three doubles at the end.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/caustics_b200.h"

namespace {

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? CAUSTICS_OK : CAUSTICS_ERR_CUDA_BASE + (int)e; }

// numpy/jax `interp` semantics: piecewise linear, clamped to the end values outside [xp[0], xp[m-1]]
__device__ __forceinline__ double interp1(double x, const double* __restrict__ xp, const double* __restrict__ fp, int m) {
  if (x <= xp[0]) return fp[0];
  if (x >= xp[m - 1]) return fp[m - 1];
  int lo = 0, hi = m - 1;  // xp[lo] <= x < xp[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid; else hi = mid;
  }
  const double dx = xp[hi] - xp[lo];
  return fp[lo] + (x - xp[lo]) / dx * (fp[hi] - fp[lo]);
}

// trajectory.py:106-158.  Annual parallax: the Sun's projected position (s_e, s_n) and velocity tabulated
// on t_jpl; delta = s(t) - s(t0) - (t - t0) * sdot(t0) is the departure from rectilinear motion.
__global__ void k_trajectory(const double* __restrict__ t, double2* __restrict__ w, int64_t n, double t0, double tE,
                             double u0, double psi, double piE, const double* __restrict__ t_jpl,
                             const double* __restrict__ s_e, const double* __restrict__ s_n,
                             const double* __restrict__ s_e_dot, const double* __restrict__ s_n_dot, int m) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ti = t[i];
  double de = 0.0, dn = 0.0;
  if (m > 0) {
    const double e0 = interp1(t0, t_jpl, s_e, m), n0 = interp1(t0, t_jpl, s_n, m);
    const double ed0 = interp1(t0, t_jpl, s_e_dot, m), nd0 = interp1(t0, t_jpl, s_n_dot, m);
    de = interp1(ti, t_jpl, s_e, m) - e0 - (ti - t0) * ed0;
    dn = interp1(ti, t_jpl, s_n, m) - n0 - (ti - t0) * nd0;
  }
  double sp, cp;
  sincos(psi, &sp, &cp);
  const double tau = (ti - t0) / tE;
  const double ue = u0 * cp + tau * sp + piE * de;
  const double un = -u0 * sp + tau * cp + piE * dn;
  w[i] = make_double2(ue, un);
}

// linalg.py:55-70, diagonal covariance: least squares for beta = (F_s, F_b) in f = F_s A + F_b, then
//   ll = -1/2 sum (f - f_pred)^2 C_inv + 1/2 log det(2 pi Sigma),  Sigma = (M^T C_inv M)^-1.
// One CTA, fixed summation order: the result does not depend on launch shape or timing.
constexpr int LL_NT = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = LL_NT / 2; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(LL_NT)
k_marg_loglike(const double* __restrict__ A, const double* __restrict__ f, const double* __restrict__ ci, int64_t n,
               double* __restrict__ out) {
  __shared__ double sh[LL_NT];
  double sAA = 0, sA = 0, s1 = 0, sAf = 0, sf = 0;
  for (int64_t i = threadIdx.x; i < n; i += LL_NT) {
    const double a = A[i], c = ci[i], y = f[i];
    sAA = fma(c * a, a, sAA);
    sA = fma(c, a, sA);
    s1 += c;
    sAf = fma(c * a, y, sAf);
    sf = fma(c, y, sf);
  }
  sAA = block_sum(sAA, sh); sA = block_sum(sA, sh); s1 = block_sum(s1, sh);
  sAf = block_sum(sAf, sh); sf = block_sum(sf, sh);
  // Sigma = inverse of [[sAA, sA], [sA, s1]]
  const double det = sAA * s1 - sA * sA;
  const double i00 = s1 / det, i01 = -sA / det, i11 = sAA / det;
  const double b0 = i00 * sAf + i01 * sf, b1 = i01 * sAf + i11 * sf;
  double chi = 0;
  for (int64_t i = threadIdx.x; i < n; i += LL_NT) {
    const double r = f[i] - fma(b0, A[i], b1);
    chi = fma(r * r, ci[i], chi);
  }
  chi = block_sum(chi, sh);
  if (threadIdx.x == 0) {
    const double two_pi = 6.283185307179586476925286766559;
    const double det_sigma = two_pi * two_pi * (i00 * i11 - i01 * i01);  // det(2 pi Sigma)
    out[0] = b0;
    out[1] = b1;
    out[2] = -0.5 * chi + 0.5 * log(det_sigma);
  }
}

}  // namespace

extern "C" {

int caustics_trajectory(const double* t, void* w, int64_t n, double t0, double tE, double u0, double psi,
                        double piE, const double* t_jpl, const double* s_e, const double* s_n,
                        const double* s_e_dot, const double* s_n_dot, int n_jpl, void* stream) {
  if (n < 0 || n_jpl < 0 || !(tE != 0.0)) return CAUSTICS_ERR_BAD_ARG;
  if (n == 0) return CAUSTICS_OK;
  if (!t || !w) return CAUSTICS_ERR_BAD_ARG;
  if (n_jpl > 0 && (!t_jpl || !s_e || !s_n || !s_e_dot || !s_n_dot)) return CAUSTICS_ERR_BAD_ARG;
  const int64_t nblk = (n + 255) / 256;
  if (nblk > 0x7fffffffLL) return CAUSTICS_ERR_BAD_ARG;
  k_trajectory<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(t, (double2*)w, n, t0, tE, u0, psi, piE, t_jpl, s_e,
                                                                 s_n, s_e_dot, s_n_dot, n_jpl);
  return cuda_rc(cudaGetLastError());
}

int caustics_marginalized_log_likelihood(const double* mag, const double* fobs, const double* c_inv, int64_t n,
                                         double* out, void* stream) {
  if (n < 2) return CAUSTICS_ERR_BAD_ARG;  // two linear parameters
  if (!mag || !fobs || !c_inv || !out) return CAUSTICS_ERR_BAD_ARG;
  k_marg_loglike<<<1, LL_NT, 0, (cudaStream_t)stream>>>(mag, fobs, c_inv, n, out);
  return cuda_rc(cudaGetLastError());
}

}  // extern "C"
