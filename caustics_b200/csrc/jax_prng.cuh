// The reference's warm-start jitters, exactly.  extended_source.py:76-85,146 draws them from a FIXED
// jax.random key (PRNGKey(0) and its two children), so they are constants of the algorithm:
//   refinement warm start of root j, new point of rank r:  + U1[j][r] + i U2[j][r],  U in (-1e-6, 1e-6),
//       U1 = uniform(key1, (deg, n)), U2 = uniform(key2, (deg, n)) -- the same table in every round
//   exact duplicates:  + V[j][p],  V = uniform(key, (deg, npts)) in (-1e-9, 1e-9), real
// JAX's default generator is counter based (threefry2x32, Salmon et al. SC'11; 20 rounds), so element i of
// an N-element float64 draw is a pure function of (key, i, N):  block (x0, x1) = (i, N + i) gives the high
// and low 32 bits, the top 52 become the mantissa of a double in [1, 2).  Each thread computes the few
// values it needs (~110 integer instructions each) instead of reading a table.  Host twin for the tests:
// oracle/jaxprng.py, pinned to the Random123 and JAX known answers.
#pragma once
#include <stdint.h>

#include "cplx.cuh"

namespace cb200 {

// jax.random.split(jax.random.PRNGKey(0)): PRNGKey(0) = (0, 0)
constexpr uint32_t JAX_KEY0_A = 0u, JAX_KEY0_B = 0u;
constexpr uint32_t JAX_KEY1_A = 4146024105u, JAX_KEY1_B = 967050713u;
constexpr uint32_t JAX_KEY2_A = 2718843009u, JAX_KEY2_B = 1272950319u;

__host__ __device__ __forceinline__ uint32_t tf_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1,
                                                      uint32_t& y0, uint32_t& y1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0]; x1 += ks[1];
#define CB200_TF4(a, b, c, d)                         \
  x0 += x1; x1 = tf_rotl(x1, a) ^ x0; x0 += x1; x1 = tf_rotl(x1, b) ^ x0; \
  x0 += x1; x1 = tf_rotl(x1, c) ^ x0; x0 += x1; x1 = tf_rotl(x1, d) ^ x0;
  CB200_TF4(13, 15, 26, 6)  x0 += ks[1]; x1 += ks[2] + 1u;
  CB200_TF4(17, 29, 16, 24) x0 += ks[2]; x1 += ks[0] + 2u;
  CB200_TF4(13, 15, 26, 6)  x0 += ks[0]; x1 += ks[1] + 3u;
  CB200_TF4(17, 29, 16, 24) x0 += ks[1]; x1 += ks[2] + 4u;
  CB200_TF4(13, 15, 26, 6)  x0 += ks[2]; x1 += ks[0] + 5u;
#undef CB200_TF4
  y0 = x0; y1 = x1;
}

// element i of jax.random.uniform(key, shape with N elements, float64, minval, maxval), C order
__host__ __device__ __forceinline__ double jax_uniform_f64(uint32_t k0, uint32_t k1, uint32_t i, uint32_t N,
                                                           double minval, double maxval) {
  uint32_t hi, lo;
  threefry2x32(k0, k1, i, N + i, hi, lo);
  const unsigned long long bits = (((unsigned long long)hi << 32) | lo) >> 12 | 0x3FF0000000000000ull;
  double f;
#ifdef __CUDA_ARCH__
  f = __longlong_as_double((long long)bits);
  const double v = __dadd_rn(__dmul_rn(f - 1.0, maxval - minval), minval);   // XLA does not contract these
#else
  memcpy(&f, &bits, 8);
  volatile double prod = (f - 1.0) * (maxval - minval);
  const double v = prod + minval;
#endif
  return v > minval ? v : minval;
}

// warm-start jitter of root j for the new limb point of rank r (n new points per round)
__device__ __forceinline__ cd limb_jitter(int j, int r, int deg, int n) {
  const uint32_t i = (uint32_t)(j * n + r), N = (uint32_t)(deg * n);
  return mk(jax_uniform_f64(JAX_KEY1_A, JAX_KEY1_B, i, N, -1e-6, 1e-6),
            jax_uniform_f64(JAX_KEY2_A, JAX_KEY2_B, i, N, -1e-6, 1e-6));
}
// offset added to an exact duplicate: root j at position p of the theta-ordered limb (npts points)
__host__ __device__ __forceinline__ double duplicate_jitter(int j, int p, int deg, int npts) {
  return jax_uniform_f64(JAX_KEY0_A, JAX_KEY0_B, (uint32_t)(j * npts + p), (uint32_t)(deg * npts), -1e-9, 1e-9);
}

}  // namespace cb200
