// Kernels 1 and 2 of the hot path and their C-ABI launchers (include/caustics_b200.h).
//   kernel 1  ea_kernel<DEG, COMP>      coeffs -> roots            (the `ehrlich_aberth` primitive)
//   kernel 2  ps_kernel<NL, COMP, MODE> w -> images / magnification (coefficients, solve, lens-
//             equation filter and Jacobian fused: nothing but w and the result touches HBM)
// One thread per polynomial / source position; see ea_core.cuh for the execution shape.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/caustics_b200.h"
#include "ea_core.cuh"
#include "lens_core.cuh"
#include "ps_walk.cuh"
#include "nvtx_range.h"
#include "tuning.h"
#include <atomic>

using namespace cb200;

namespace {

constexpr int NT = 128;  // threads per CTA (4 warps): 20 KB (deg 10) of root planes per CTA

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? CAUSTICS_OK : CAUSTICS_ERR_CUDA_BASE + (int)e; }

// ------------------------------------------------------------------------------------------
#ifndef CB200_EA_MINBLOCKS
#define CB200_EA_MINBLOCKS 5
#endif
// degrees <= 10 (the lens polynomials): 128-thread CTAs, 5 CTAs/SM (96 registers);
// degrees 11..16: 64-thread CTAs (the shared planes stay below 48 KB), no register cap
template <int DEG, bool COMP, int NTK>
__global__ void __launch_bounds__(NTK, DEG <= 10 ? CB200_EA_MINBLOCKS : 1)
ea_kernel(const double2* __restrict__ coeffs, const double2* __restrict__ roots_init,
          double2* __restrict__ roots, int32_t* __restrict__ sweeps, int64_t size, int itmax,
          int custom_init, int flags) {
  __shared__ EASmem<DEG, COMP, NTK> sm;
  const int tid = threadIdx.x;
  const int64_t idx = (int64_t)blockIdx.x * NTK + tid;
  const bool active = idx < size;
  const int init_mode = flags & CAUSTICS_FLAG_INIT_BINI ? EA_INIT_BINI : EA_INIT_REFERENCE;
  const bool high_first = (flags & CAUSTICS_FLAG_COEFFS_HIGH_FIRST) != 0;
  cd p[DEG + 1];
  if (active) {
    const double2* src = coeffs + idx * (DEG + 1);
#pragma unroll
    for (int k = 0; k <= DEG; ++k) {
      const double2 v = __ldg(src + (high_first ? DEG - k : k));
      p[k] = mk(v.x, v.y);
    }
    ea_normalise<DEG>(p);
    if (custom_init) {
      const double2* ri = roots_init + idx * DEG;
#pragma unroll
      for (int j = 0; j < DEG; ++j) {
        const double2 v = __ldg(ri + j);
        sm.zre[j][tid] = v.x;
        sm.zim[j][tid] = v.y;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k <= DEG; ++k) p[k] = mk(k == 0 ? -1.0 : (k == DEG ? 1.0 : 0.0), 0.0);
  }
  const EAResult r = ea_solve_thread<DEG, COMP, NTK>(p, sm, tid, active, custom_init != 0, init_mode, itmax,
                                                     init_mode == EA_INIT_BINI);
  if (active) {
    double2* dst = roots + idx * DEG;
#pragma unroll
    for (int j = 0; j < DEG; ++j) dst[j] = make_double2(sm.zre[j][tid], sm.zim[j][tid]);
    if (sweeps) sweeps[idx] = r.converged ? r.sweeps : -r.sweeps;
  }
}

template <int DEG>
int launch_ea(const void* coeffs, const void* roots_init, void* roots, int32_t* sweeps, int64_t size,
              int itmax, int compensated, int custom_init, int flags, cudaStream_t st) {
  constexpr int NTK = DEG <= 10 ? NT : 64;
  const int64_t nblk = (size + NTK - 1) / NTK;
  if (nblk > 0x7fffffffLL) return CAUSTICS_ERR_BAD_ARG;
  dim3 grid((unsigned)nblk), block(NTK);
  if (compensated)
    ea_kernel<DEG, true, NTK><<<grid, block, 0, st>>>((const double2*)coeffs, (const double2*)roots_init,
                                                      (double2*)roots, sweeps, size, itmax, custom_init, flags);
  else
    ea_kernel<DEG, false, NTK><<<grid, block, 0, st>>>((const double2*)coeffs, (const double2*)roots_init,
                                                       (double2*)roots, sweeps, size, itmax, custom_init, flags);
  return cuda_rc(cudaGetLastError());
}

// Implicit-function tangent and cotangent of the roots (ehrlich_aberth_primitive.py:254-324):
//   JVP  dz_j = -(sum_k dp_k z_j^k) / p'(z_j)
//   VJP  gp_k = sum_j conj(-z_j^k / p'(z_j)) gz_j      (the transpose JAX derives from the JVP)
// One thread per polynomial, Horner in registers; memory-bound (no (size, deg, deg+1) temporary).
__global__ void ea_jvp_kernel(const double2* __restrict__ coeffs, const double2* __restrict__ roots,
                              const double2* __restrict__ dcoeffs, double2* __restrict__ droots, int64_t size, int deg) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= size) return;
  const double2* p = coeffs + n * (deg + 1);
  const double2* dp = dcoeffs + n * (deg + 1);
  for (int j = 0; j < deg; ++j) {
    const double2 zz = roots[n * deg + j];
    const cd z = mk(zz.x, zz.y);
    cd num = mk(0, 0), der = mk(0, 0);
    for (int k = deg; k >= 0; --k) {
      num = cfma(num, z, mk(dp[k].x, dp[k].y));
      if (k >= 1) der = cfma(der, z, (double)k * mk(p[k].x, p[k].y));
    }
    const cd r = cdiv(-num, der);
    droots[n * deg + j] = make_double2(r.re, r.im);
  }
}
__global__ void ea_vjp_kernel(const double2* __restrict__ coeffs, const double2* __restrict__ roots,
                              const double2* __restrict__ groots, double2* __restrict__ gcoeffs, int64_t size, int deg) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= size) return;
  const double2* p = coeffs + n * (deg + 1);
  cd acc[33];
  for (int k = 0; k <= deg; ++k) acc[k] = mk(0, 0);
  for (int j = 0; j < deg; ++j) {
    const double2 zz = roots[n * deg + j], gg = groots[n * deg + j];
    const cd z = mk(zz.x, zz.y);
    cd der = mk(0, 0);
    for (int k = deg; k >= 1; --k) der = cfma(der, z, (double)k * mk(p[k].x, p[k].y));
    const cd g = cdiv(mk(gg.x, gg.y), conj(-der));
    const cd zc = conj(z);
    cd zk = mk(1, 0);
    for (int k = 0; k <= deg; ++k) { acc[k] = cfma(g, zk, acc[k]); zk = zk * zc; }
  }
  for (int k = 0; k <= deg; ++k) gcoeffs[n * (deg + 1) + k] = make_double2(acc[k].re, acc[k].im);
}

// ------------------------------------------------------------------------------------------
// kernel 2.  MODE 0: magnification only; MODE 1: images + mask (root axis first, like the reference)
enum { PS_MAG = 0, PS_IMAGES = 1 };

struct GridSpec {
  double x0, y0, dx, dy;
  int64_t nx, row_begin;
  int use;  // 1: generate w from the grid instead of reading it
};

template <int NL, bool COMP, int MODE>
__global__ void __launch_bounds__(NT)
ps_kernel(const double2* __restrict__ w_in, GridSpec g, const double2* __restrict__ z_init,
          double2* __restrict__ z_out, uint8_t* __restrict__ mask_out, double* __restrict__ mag,
          uint8_t* __restrict__ nimg, int64_t n, LensConst L, int itmax, int custom_init, int flags) {
  constexpr int DEG = NL * NL + 1;
  __shared__ EASmem<DEG, COMP, NT> sm;
  const int tid = threadIdx.x;
  const int64_t idx = (int64_t)blockIdx.x * NT + tid;
  const bool active = idx < n;
  // The magnification is a sum over images: the root ORDER is irrelevant, so the fused magnification
  // always starts from the intended complex Bini estimates (~10 % fewer updates than the reference's
  // real-axis guesses); the images mode keeps the reference-compatible order unless asked otherwise.
  const int init_mode = (MODE == PS_MAG || (flags & CAUSTICS_FLAG_INIT_BINI)) ? EA_INIT_BINI : EA_INIT_REFERENCE;
  cd w = mk(0.3, 0.2);
  if (active) {
    if (g.use) {
      const int64_t iy = idx / g.nx, ix = idx - iy * g.nx;
      w = mk(fma((double)ix, g.dx, g.x0), fma((double)(iy + g.row_begin), g.dy, g.y0));
    } else {
      const double2 v = __ldg(w_in + idx);
      w = mk(v.x, v.y);
    }
    w.re += L.x_cm;
  }
  cd p[DEG + 1];
  lens_poly<NL>(L, w, p);
  ea_normalise<DEG>(p);
  if (MODE == PS_IMAGES) {
    if (active && custom_init) {
      const double2* ri = z_init + idx * DEG;
#pragma unroll
      for (int j = 0; j < DEG; ++j) {
        const double2 v = __ldg(ri + j);
        sm.zre[j][tid] = v.x;
        sm.zim[j][tid] = v.y;
      }
    }
  }
  ea_solve_thread<DEG, COMP, NT>(p, sm, tid, active, MODE == PS_IMAGES && custom_init != 0, init_mode, itmax,
                                 init_mode == EA_INIT_BINI);
  if (!active) return;
  double mu = 0.0;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < DEG; ++j) {
    const cd z = mk(sm.zre[j][tid], sm.zim[j][tid]);
    bool real_image;
    double detj;
    image_eval<NL>(L, z, w, real_image, detj);
    if (MODE == PS_IMAGES) {
      z_out[(int64_t)j * n + idx] = make_double2(z.re, z.im);
      mask_out[(int64_t)j * n + idx] = real_image ? 1 : 0;
    } else {
      // (1/|det J|) * mask, point_source.py:1829 -- a false image contributes exactly 0
      if (real_image) { mu += 1.0 / fabs(detj); ++cnt; }
    }
  }
  if (MODE == PS_MAG) {
    mag[idx] = mu;
    if (nimg) nimg[idx] = (uint8_t)cnt;
  }
}

// _images_point_source_sequential (point_source.py:1711-1759): a path of source positions is a scan --
// position 0 starts from the default initial estimates, position k from the images of position k-1
// (custom_init), so row j of the result follows one image along the path.  One thread per path; the
// roots never leave shared memory between positions.
template <int NL, bool COMP>
__global__ void __launch_bounds__(NT)
ps_seq_kernel(const double2* __restrict__ w_in, double2* __restrict__ z_out, uint8_t* __restrict__ mask_out,
              int64_t npaths, int64_t n, LensConst L, int itmax) {
  constexpr int DEG = NL * NL + 1;
  __shared__ EASmem<DEG, COMP, NT> sm;
  const int tid = threadIdx.x;
  const int64_t idx = (int64_t)blockIdx.x * NT + tid;
  const bool active = idx < npaths;
  for (int64_t k = 0; k < n; ++k) {
    cd w = mk(0.3, 0.2);
    if (active) {
      const double2 v = __ldg(w_in + idx * n + k);
      w = mk(v.x + L.x_cm, v.y);
    }
    cd p[DEG + 1];
    lens_poly<NL>(L, w, p);
    ea_normalise<DEG>(p);
    ea_solve_thread<DEG, COMP, NT, false>(p, sm, tid, active, k > 0, EA_INIT_REFERENCE, itmax);
    if (active) {
#pragma unroll
      for (int j = 0; j < DEG; ++j) {
        const cd z = mk(sm.zre[j][tid], sm.zim[j][tid]);
        bool real_image;
        double detj;
        image_eval<NL>(L, z, w, real_image, detj);
        z_out[(idx * DEG + j) * n + k] = make_double2(z.re, z.im);
        mask_out[(idx * DEG + j) * n + k] = real_image ? 1 : 0;
      }
    }
  }
}

// Magnification map as warm-started column walks (ps_walk.cuh; CAUSTICS_FLAG_GRID_WALK).  CTA b covers
// columns (b % ncolblk) * NTW ... and rows (b / ncolblk) * run ... of the requested row block.
// 128-thread CTAs for the binary lens, 64 for the triple (root + previous-root planes stay < 48 KB).
template <int NL> struct WalkNT { static constexpr int value = NL == 2 ? 128 : 64; };

template <int NL, bool COMP>
__global__ void __launch_bounds__(WalkNT<NL>::value)
ps_grid_walk_kernel(GridSpec g, double* __restrict__ mag, int64_t nrows, int run, int64_t ncolblk, LensConst L,
                    int itmax, int extrap) {
  constexpr int DEG = NL * NL + 1, NTW = WalkNT<NL>::value;
  __shared__ EASmem<DEG, COMP, NTW> sm;
  __shared__ double pre[DEG][NTW], pim[DEG][NTW];
  const int tid = threadIdx.x;
  const int64_t cb = (int64_t)blockIdx.x % ncolblk, rg = (int64_t)blockIdx.x / ncolblk;
  const int64_t ix = cb * NTW + tid, row0 = rg * run;
  const bool active = ix < g.nx;
  const int64_t left = nrows - row0;
  const int nrun = left < run ? (int)left : run;
  WalkColumn src;
  src.wx = (active ? fma((double)ix, g.dx, g.x0) : 0.3) + L.x_cm;
  src.y0 = g.y0; src.dy = g.dy; src.row_abs0 = g.row_begin + row0;
  ps_walk_body<NL, COMP, NTW>(src, nrun, active ? nrun : 0, mag + row0 * g.nx + (active ? ix : 0), g.nx,
                              L, itmax, extrap != 0, sm, &pre[0][tid], &pim[0][tid], tid);
}

// A 1-D array of source positions whose consecutive elements are neighbours (CAUSTICS_FLAG_PATH_WALK):
// thread t owns elements [t * run, (t + 1) * run).
template <int NL, bool COMP>
__global__ void __launch_bounds__(WalkNT<NL>::value)
ps_path_walk_kernel(const double* __restrict__ w, double* __restrict__ mag, int64_t n, int run, LensConst L,
                    int itmax, int extrap) {
  constexpr int DEG = NL * NL + 1, NTW = WalkNT<NL>::value;
  __shared__ EASmem<DEG, COMP, NTW> sm;
  __shared__ double pre[DEG][NTW], pim[DEG][NTW];
  const int tid = threadIdx.x;
  const int64_t first = ((int64_t)blockIdx.x * NTW + tid) * run;
  const int64_t left = n - first;
  const int nrun = left <= 0 ? 0 : (left < run ? (int)left : run);
  WalkPath src;
  src.w = w + 2 * (nrun ? first : 0);
  src.x_cm = L.x_cm;
  ps_walk_body<NL, COMP, NTW>(src, run, nrun, mag + (nrun ? first : 0), 1, L, itmax, extrap != 0, sm,
                              &pre[0][tid], &pim[0][tid], tid);
}

// rows per walk: 32 when the map is large enough to fill the machine several times over, shorter walks
// (more CTAs) for small maps; caustics_set_tuning("grid_run", v) overrides (tests, experiments)
int grid_walk_run(int64_t ncolblk, int64_t nrows) {
  { const int v = cb200::tuning_get(cb200::TUNE_GRID_RUN); if (v >= 1 && v <= 4096) return v; }
  int run = 32;
  while (run > 4 && ncolblk * ((nrows + run - 1) / run) < 148 * 16) run >>= 1;
  return run;
}

// elements per thread of a path walk: long enough to amortise the cold first solve, short enough that
// the batch still fills the machine (>= ~16 warps per SM)
int path_walk_run(int64_t n) {
  { const int v = cb200::tuning_get(cb200::TUNE_PATH_RUN); if (v >= 1 && v <= 4096) return v; }
  int run = 32;
  while (run > 1 && n / run < 148 * 16 * 32) run >>= 1;
  return run;
}

template <int NL>
int launch_path_walk(const void* w, double* mag, int64_t n, const LensConst& L, int itmax, int compensated, cudaStream_t st) {
  constexpr int NTW = WalkNT<NL>::value;
  const int run = path_walk_run(n);
  const int64_t nthreads = (n + run - 1) / run;
  const int64_t nblk = (nthreads + NTW - 1) / NTW;
  if (nblk > 0x7fffffffLL) return CAUSTICS_ERR_BAD_ARG;
  const int extrap = cb200::tuning_get(cb200::TUNE_GRID_EXTRAP) == 0 ? 0 : 1;   // default (unset = -1): on
  if (compensated) ps_path_walk_kernel<NL, true><<<(unsigned)nblk, NTW, 0, st>>>((const double*)w, mag, n, run, L, itmax, extrap);
  else ps_path_walk_kernel<NL, false><<<(unsigned)nblk, NTW, 0, st>>>((const double*)w, mag, n, run, L, itmax, extrap);
  return cuda_rc(cudaGetLastError());
}

template <int NL>
int launch_grid_walk(GridSpec g, double* mag, int64_t nrows, const LensConst& L, int itmax, int compensated, cudaStream_t st) {
  constexpr int NTW = WalkNT<NL>::value;
  const int64_t ncolblk = (g.nx + NTW - 1) / NTW;
  const int run = grid_walk_run(ncolblk, nrows);
  const int64_t nblk = ncolblk * ((nrows + run - 1) / run);
  if (nblk > 0x7fffffffLL) return CAUSTICS_ERR_BAD_ARG;
  const int extrap = cb200::tuning_get(cb200::TUNE_GRID_EXTRAP) == 0 ? 0 : 1;   // default (unset = -1): on
  if (compensated) ps_grid_walk_kernel<NL, true><<<(unsigned)nblk, NTW, 0, st>>>(g, mag, nrows, run, ncolblk, L, itmax, extrap);
  else ps_grid_walk_kernel<NL, false><<<(unsigned)nblk, NTW, 0, st>>>(g, mag, nrows, run, ncolblk, L, itmax, extrap);
  return cuda_rc(cudaGetLastError());
}

int make_lens_const(const caustics_lens* lens, LensConst* out) {
  if (!lens) return CAUSTICS_ERR_BAD_ARG;
  LensConst L;
  memset(&L, 0, sizeof(L));
  L.nlenses = lens->nlenses;
  L.x_cm = lens->x_cm;
  struct hc { double re, im; };
  auto mul = [](hc a, hc b) { return hc{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; };
  hc r[3] = {{lens->a, 0.0}, {-lens->a, 0.0}, {lens->r3_re, lens->r3_im}};
  double eps[3];
  int n = lens->nlenses;
  if (n == 2) { eps[0] = lens->e1; eps[1] = 1.0 - lens->e1; eps[2] = 0.0; }
  else if (n == 3) { eps[0] = lens->e1; eps[1] = lens->e2; eps[2] = 1.0 - lens->e1 - lens->e2; }
  else return CAUSTICS_ERR_BAD_ARG;
  // H = prod (z - r_i), low->high
  hc H[4] = {{1, 0}, {0, 0}, {0, 0}, {0, 0}};
  int dh = 0;
  for (int i = 0; i < n; ++i) {
    hc nH[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (int k = 0; k <= dh; ++k) {
      hc t = mul(H[k], hc{-r[i].re, -r[i].im});
      nH[k].re += t.re; nH[k].im += t.im;
      nH[k + 1].re += H[k].re; nH[k + 1].im += H[k].im;
    }
    ++dh;
    for (int k = 0; k <= dh; ++k) H[k] = nH[k];
  }
  // G = sum_j eps_j prod_{i != j} (z - r_i)
  hc G[3] = {{0, 0}, {0, 0}, {0, 0}};
  for (int j = 0; j < n; ++j) {
    hc t[3] = {{eps[j], 0}, {0, 0}, {0, 0}};
    int dt = 0;
    for (int i = 0; i < n; ++i) {
      if (i == j) continue;
      hc nt[3] = {{0, 0}, {0, 0}, {0, 0}};
      for (int k = 0; k <= dt; ++k) {
        hc u = mul(t[k], hc{-r[i].re, -r[i].im});
        nt[k].re += u.re; nt[k].im += u.im;
        nt[k + 1].re += t[k].re; nt[k + 1].im += t[k].im;
      }
      ++dt;
      for (int k = 0; k <= dt; ++k) t[k] = nt[k];
    }
    for (int k = 0; k < n; ++k) { G[k].re += t[k].re; G[k].im += t[k].im; }
  }
  for (int i = 0; i < 3; ++i) { L.eps[i] = eps[i]; L.r[i].re = r[i].re; L.r[i].im = r[i].im; }
  for (int k = 0; k < 4; ++k) { L.H[k].re = H[k].re; L.H[k].im = H[k].im; }
  for (int k = 0; k < 3; ++k) { L.G[k].re = G[k].re; L.G[k].im = G[k].im; }
  *out = L;
  return CAUSTICS_OK;
}

template <int NL, int MODE>
int launch_ps(const double2* w, GridSpec g, const double2* z_init, double2* z, uint8_t* mask, double* mag,
              uint8_t* nimg, int64_t n, const LensConst& L, int itmax, int compensated, int custom_init,
              int flags, cudaStream_t st) {
  const int64_t nblk = (n + NT - 1) / NT;
  if (nblk > 0x7fffffffLL) return CAUSTICS_ERR_BAD_ARG;
  if (nblk == 0) return CAUSTICS_OK;
  dim3 grid((unsigned)nblk), block(NT);
  if (compensated)
    ps_kernel<NL, true, MODE><<<grid, block, 0, st>>>(w, g, z_init, z, mask, mag, nimg, n, L, itmax, custom_init, flags);
  else
    ps_kernel<NL, false, MODE><<<grid, block, 0, st>>>(w, g, z_init, z, mask, mag, nimg, n, L, itmax, custom_init, flags);
  return cuda_rc(cudaGetLastError());
}

// Sticky status of the XLA custom calls, process-wide: XLA runs custom calls on its own launcher
// thread, the Python side that wants to know reads from another one.  The FIRST error sticks until read.
std::atomic<int> g_last_xla_error{0};
inline void set_xla_error(int rc) {
  if (rc == 0) return;
  int expected = 0;
  g_last_xla_error.compare_exchange_strong(expected, rc);
}


// Same, but every DFMA reads three DISTINCT, changing register pairs (no constant or repeated
// operand): the register-file-limited rate that real Horner / Aberth code sees.
__global__ void __launch_bounds__(256) fp64_peak3_kernel(double* sink, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  double b0 = 1.0000001, b1 = 0.9999999, b2 = 1.0000002, b3 = 0.9999998, b4 = 1.0000003, b5 = 0.9999997, b6 = 1.0000004, b7 = 0.9999996;
  double c0 = 1e-9 * threadIdx.x, c1 = c0 + 1e-9, c2 = c0 + 2e-9, c3 = c0 + 3e-9, c4 = c0 + 4e-9, c5 = c0 + 5e-9, c6 = c0 + 6e-9, c7 = c0 + 7e-9;
#pragma unroll 2
  for (int i = 0; i < iters; i += 3) {
    // rotate roles so no operand is loop-invariant: a = a*b+c ; b = b*c+a ; c = c*a+b
    a0 = fma(a0, b0, c0); a1 = fma(a1, b1, c1); a2 = fma(a2, b2, c2); a3 = fma(a3, b3, c3);
    a4 = fma(a4, b4, c4); a5 = fma(a5, b5, c5); a6 = fma(a6, b6, c6); a7 = fma(a7, b7, c7);
    b0 = fma(b0, c0, a0); b1 = fma(b1, c1, a1); b2 = fma(b2, c2, a2); b3 = fma(b3, c3, a3);
    b4 = fma(b4, c4, a4); b5 = fma(b5, c5, a5); b6 = fma(b6, c6, a6); b7 = fma(b7, c7, a7);
    c0 = fma(c0, a0, b0); c1 = fma(c1, a1, b1); c2 = fma(c2, a2, b2); c3 = fma(c3, a3, b3);
    c4 = fma(c4, a4, b4); c5 = fma(c5, a5, b5); c6 = fma(c6, a6, b6); c7 = fma(c7, a7, b7);
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)) + ((b0 + b1) + (b2 + b3)) + ((c4 + c5) + (c6 + c7));
  if (r == 123.456) sink[0] = r;
}

// Greedy nearest-neighbour track matching along the first axis (utils.match_points, utils.py:15-40,
// as used by critical_and_caustic_curves, point_source.py:1637-1641): one thread per curve set.
// z (B, npts, D) complex128 -> out (B, npts, D) with column k ordered so that entry i continues
// entry i of column k-1 (for each i in order: the nearest not-yet-taken root, ties to the lowest).
__global__ void match_tracks_kernel(const double2* __restrict__ z, double2* __restrict__ out, int64_t B, int npts, int D) {
  const int64_t bidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bidx >= B) return;
  const double2* zi = z + bidx * npts * D;
  double2* zo = out + bidx * npts * D;
  double cre[16], cim[16];
  for (int i = 0; i < D; ++i) { cre[i] = zi[i].x; cim[i] = zi[i].y; zo[i] = zi[i]; }
  for (int k = 1; k < npts; ++k) {
    unsigned used = 0;
    double nre[16], nim[16];
    for (int i = 0; i < D; ++i) {
      double bd = 1e300; int best = -1;
      for (int m = 0; m < D; ++m) {
        const double dx = zi[k * D + m].x - cre[i], dy = zi[k * D + m].y - cim[i];
        const double d2 = dx * dx + dy * dy;
        if (!((used >> m) & 1u) && d2 < bd) { bd = d2; best = m; }
      }
      if (best < 0) for (int m = D - 1; m >= 0; --m) if (!((used >> m) & 1u)) best = m;
      used |= 1u << best;
      nre[i] = zi[k * D + best].x; nim[i] = zi[k * D + best].y;
      zo[k * D + i] = zi[k * D + best];
    }
    for (int i = 0; i < D; ++i) { cre[i] = nre[i]; cim[i] = nim[i]; }
  }
}

// DFMA-saturating microbenchmark: 8 independent FMA chains per thread.  Used by bench.py to measure
// the FP64 (non-tensor) roofline denominator on the device the numbers are taken on.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000001, c = 1e-9 * threadIdx.x;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 123.456) sink[0] = r;  // never true: keeps the chains alive
}

}  // namespace

// ============================================================================================
// used by extended.cu (same lens constants for every kernel family); not part of the public ABI
extern "C" __attribute__((visibility("hidden"))) int caustics_internal_lens_const(const caustics_lens* lens, void* out) {
  return make_lens_const(lens, (LensConst*)out);
}

extern "C" {

int caustics_ea_jvp(const void* coeffs, const void* roots, const void* dcoeffs, void* droots, int64_t size, int deg,
                    void* stream) {
  CB200_NVTX("caustics_ea_jvp");
  if (size < 0 || deg < 1 || deg > 32) return CAUSTICS_ERR_BAD_ARG;
  if (size == 0) return CAUSTICS_OK;
  if (!coeffs || !roots || !dcoeffs || !droots) return CAUSTICS_ERR_BAD_ARG;
  ea_jvp_kernel<<<(unsigned)((size + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (const double2*)coeffs, (const double2*)roots, (const double2*)dcoeffs, (double2*)droots, size, deg);
  return cuda_rc(cudaGetLastError());
}

int caustics_ea_vjp(const void* coeffs, const void* roots, const void* groots, void* gcoeffs, int64_t size, int deg,
                    void* stream) {
  CB200_NVTX("caustics_ea_vjp");
  if (size < 0 || deg < 1 || deg > 32) return CAUSTICS_ERR_BAD_ARG;
  if (size == 0) return CAUSTICS_OK;
  if (!coeffs || !roots || !groots || !gcoeffs) return CAUSTICS_ERR_BAD_ARG;
  ea_vjp_kernel<<<(unsigned)((size + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (const double2*)coeffs, (const double2*)roots, (const double2*)groots, (double2*)gcoeffs, size, deg);
  return cuda_rc(cudaGetLastError());
}

int caustics_bench_fp64_peak3(double* sink, int blocks, int iters, void* stream) {
  if (!sink || blocks <= 0 || iters <= 0) return CAUSTICS_ERR_BAD_ARG;
  fp64_peak3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, iters, 0.5);
  return cuda_rc(cudaGetLastError());
}

int caustics_match_tracks(const void* z, void* out, int64_t nsets, int npts, int deg, void* stream) {
  CB200_NVTX("caustics_match_tracks");
  if (nsets < 0 || npts < 0 || deg < 1 || deg > 16) return CAUSTICS_ERR_BAD_ARG;
  if (nsets == 0 || npts == 0) return CAUSTICS_OK;
  if (!z || !out) return CAUSTICS_ERR_BAD_ARG;
  match_tracks_kernel<<<(unsigned)((nsets + 63) / 64), 64, 0, (cudaStream_t)stream>>>((const double2*)z, (double2*)out, nsets, npts, deg);
  return cuda_rc(cudaGetLastError());
}

int caustics_bench_fp64_peak(double* sink, int blocks, int iters, void* stream) {
  if (!sink || blocks <= 0 || iters <= 0) return CAUSTICS_ERR_BAD_ARG;
  fp64_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, iters, 0.5);
  return cuda_rc(cudaGetLastError());
}

const char* caustics_version(void) { return "caustics_b200 0.2 (sm_100a)"; }

int caustics_set_tuning(const char* key, int value) { return cb200::tuning_set(key, value) ? CAUSTICS_OK : CAUSTICS_ERR_BAD_ARG; }

int caustics_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int caustics_ea_degree_supported(int deg) { return deg >= 2 && deg <= 16; }

const char* caustics_error_string(int code) {
  switch (code) {
    case CAUSTICS_OK: return "ok";
    case CAUSTICS_ERR_BAD_ARG: return "bad argument";
    case CAUSTICS_ERR_UNSUPPORTED_DEGREE: return "polynomial degree has no instantiated kernel (supported: 2..16)";
    case CAUSTICS_ERR_BAD_DESCRIPTOR: return "opaque descriptor has the wrong size";
    default:
      if (code >= CAUSTICS_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(code - CAUSTICS_ERR_CUDA_BASE));
      return "unknown error";
  }
}

int caustics_ea_solve(const void* coeffs, const void* roots_init, void* roots, int32_t* sweeps,
                      int64_t size, int deg, int itmax, int compensated, int custom_init,
                      int flags, void* stream) {
  CB200_NVTX("caustics_ea_solve");
  if (size < 0 || itmax < 0) return CAUSTICS_ERR_BAD_ARG;
  if (!caustics_ea_degree_supported(deg)) return CAUSTICS_ERR_UNSUPPORTED_DEGREE;
  if (size == 0) return CAUSTICS_OK;
  if (!coeffs || !roots || (custom_init && !roots_init)) return CAUSTICS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define CB200_CASE(D) \
  case D: return launch_ea<D>(coeffs, roots_init, roots, sweeps, size, itmax, compensated, custom_init, flags, st);
  switch (deg) {
    CB200_CASE(2) CB200_CASE(3) CB200_CASE(4) CB200_CASE(5) CB200_CASE(6)
    CB200_CASE(7) CB200_CASE(8) CB200_CASE(9) CB200_CASE(10) CB200_CASE(11) CB200_CASE(12)
    CB200_CASE(13) CB200_CASE(14) CB200_CASE(15) CB200_CASE(16)
  }
#undef CB200_CASE
  return CAUSTICS_ERR_UNSUPPORTED_DEGREE;
}

size_t caustics_ea_make_descriptor(caustics_ea_descriptor* out, int64_t size, int deg, int itmax,
                                   int compensated, int custom_init, int flags) {
  if (out) {
    memset(out, 0, sizeof(*out));
    out->size = size; out->deg = deg; out->itmax = itmax;
    out->compensated = compensated ? 1 : 0; out->custom_init = custom_init ? 1 : 0;
    out->flags = (uint8_t)flags;
  }
  return sizeof(caustics_ea_descriptor);
}

int caustics_last_xla_error(void) { return g_last_xla_error.exchange(0); }

void caustics_ea_xla(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  if (opaque_len != sizeof(caustics_ea_descriptor) || !opaque || !buffers) {
    set_xla_error(CAUSTICS_ERR_BAD_DESCRIPTOR);  // the reference throws here (kernel_helpers.h:37-42)
    return;
  }
  caustics_ea_descriptor d;
  memcpy(&d, opaque, sizeof(d));
  set_xla_error(caustics_ea_solve(buffers[0], buffers[1], buffers[2], nullptr, d.size, d.deg, d.itmax,
                                  d.compensated, d.custom_init, d.flags, stream));
}

static_assert(sizeof(caustics_mag_ps_descriptor) == 72 && sizeof(caustics_mag_ext_descriptor) == 112 &&
                  sizeof(caustics_ea_descriptor) == 24 && sizeof(caustics_lens) == 56,
              "descriptor layouts are part of the ABI");

void caustics_mag_ps_xla(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  if (opaque_len != sizeof(caustics_mag_ps_descriptor) || !opaque || !buffers) {
    set_xla_error(CAUSTICS_ERR_BAD_DESCRIPTOR);
    return;
  }
  caustics_mag_ps_descriptor d;
  memcpy(&d, opaque, sizeof(d));
  set_xla_error(caustics_mag_point_source(buffers[0], (double*)buffers[1], nullptr, d.n, &d.lens, d.itmax,
                                          d.compensated, d.flags, stream));
}

void caustics_mag_ext_xla(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  if (opaque_len != sizeof(caustics_mag_ext_descriptor) || !opaque || !buffers) {
    set_xla_error(CAUSTICS_ERR_BAD_DESCRIPTOR);
    return;
  }
  caustics_mag_ext_descriptor d;
  memcpy(&d, opaque, sizeof(d));
  set_xla_error(
      d.gate ? caustics_mag(buffers[0], (double*)buffers[1], nullptr, d.n, d.rho, &d.lens, d.q, d.npts_limb,
                            d.limb_darkening, d.u1, d.npts_ld, d.itmax, d.compensated, buffers[2],
                            (size_t)d.workspace_bytes, stream)
             : caustics_mag_extended_source(buffers[0], (double*)buffers[1], d.n, d.rho, &d.lens, d.npts_limb,
                                            d.limb_darkening, d.u1, d.npts_ld, d.itmax, d.compensated, buffers[2],
                                            (size_t)d.workspace_bytes, stream));
}

int caustics_images_point_source(const void* w, const void* z_init, void* z, uint8_t* mask,
                                 int64_t n, const caustics_lens* lens, int itmax, int compensated,
                                 int custom_init, int flags, void* stream) {
  CB200_NVTX("caustics_images_point_source");
  LensConst L;
  int rc = make_lens_const(lens, &L);
  if (rc) return rc;
  if (n < 0 || itmax < 0) return CAUSTICS_ERR_BAD_ARG;
  if (n == 0) return CAUSTICS_OK;
  if (!w || !z || !mask || (custom_init && !z_init)) return CAUSTICS_ERR_BAD_ARG;
  GridSpec g; memset(&g, 0, sizeof(g));
  cudaStream_t st = (cudaStream_t)stream;
  if (L.nlenses == 2)
    return launch_ps<2, PS_IMAGES>((const double2*)w, g, (const double2*)z_init, (double2*)z, mask, nullptr, nullptr, n, L, itmax, compensated, custom_init, flags, st);
  return launch_ps<3, PS_IMAGES>((const double2*)w, g, (const double2*)z_init, (double2*)z, mask, nullptr, nullptr, n, L, itmax, compensated, custom_init, flags, st);
}

int caustics_images_point_source_sequential(const void* w, void* z, uint8_t* mask, int64_t npaths, int64_t n,
                                            const caustics_lens* lens, int itmax, int compensated, void* stream) {
  CB200_NVTX("caustics_images_point_source_sequential");
  LensConst L;
  int rc = make_lens_const(lens, &L);
  if (rc) return rc;
  if (npaths < 0 || n < 0 || itmax < 0) return CAUSTICS_ERR_BAD_ARG;
  if (npaths == 0 || n == 0) return CAUSTICS_OK;
  if (!w || !z || !mask) return CAUSTICS_ERR_BAD_ARG;
  const int64_t nblk = (npaths + NT - 1) / NT;
  if (nblk > 0x7fffffffLL) return CAUSTICS_ERR_BAD_ARG;
  dim3 grid((unsigned)nblk), block(NT);
  cudaStream_t st = (cudaStream_t)stream;
#define CB200_SEQ(NL, COMP) ps_seq_kernel<NL, COMP><<<grid, block, 0, st>>>((const double2*)w, (double2*)z, mask, npaths, n, L, itmax)
  if (L.nlenses == 2) { if (compensated) CB200_SEQ(2, true); else CB200_SEQ(2, false); }
  else                { if (compensated) CB200_SEQ(3, true); else CB200_SEQ(3, false); }
#undef CB200_SEQ
  return cuda_rc(cudaGetLastError());
}

int caustics_mag_point_source(const void* w, double* mag, uint8_t* nimages, int64_t n,
                              const caustics_lens* lens, int itmax, int compensated, int flags,
                              void* stream) {
  CB200_NVTX("caustics_mag_point_source");
  LensConst L;
  int rc = make_lens_const(lens, &L);
  if (rc) return rc;
  if (n < 0 || itmax < 0) return CAUSTICS_ERR_BAD_ARG;
  if (n == 0) return CAUSTICS_OK;
  if (!w || !mag) return CAUSTICS_ERR_BAD_ARG;
  GridSpec g; memset(&g, 0, sizeof(g));
  cudaStream_t st = (cudaStream_t)stream;
  if ((flags & CAUSTICS_FLAG_PATH_WALK) && !nimages && path_walk_run(n) > 1)
    return L.nlenses == 2 ? launch_path_walk<2>(w, mag, n, L, itmax, compensated, st)
                          : launch_path_walk<3>(w, mag, n, L, itmax, compensated, st);
  if (L.nlenses == 2)
    return launch_ps<2, PS_MAG>((const double2*)w, g, nullptr, nullptr, nullptr, mag, nimages, n, L, itmax, compensated, 0, flags, st);
  return launch_ps<3, PS_MAG>((const double2*)w, g, nullptr, nullptr, nullptr, mag, nimages, n, L, itmax, compensated, 0, flags, st);
}

int caustics_mag_point_source_grid(double x0, double y0, double dx, double dy, int64_t nx,
                                   int64_t row_begin, int64_t row_end, double* mag,
                                   const caustics_lens* lens, int itmax, int compensated,
                                   int flags, void* stream) {
  CB200_NVTX("caustics_mag_point_source_grid");
  LensConst L;
  int rc = make_lens_const(lens, &L);
  if (rc) return rc;
  if (nx <= 0 || row_end < row_begin || itmax < 0) return CAUSTICS_ERR_BAD_ARG;
  const int64_t n = (row_end - row_begin) * nx;
  if (n == 0) return CAUSTICS_OK;
  if (!mag) return CAUSTICS_ERR_BAD_ARG;
  GridSpec g; g.x0 = x0; g.y0 = y0; g.dx = dx; g.dy = dy; g.nx = nx; g.row_begin = row_begin; g.use = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (flags & CAUSTICS_FLAG_GRID_WALK)
    return L.nlenses == 2 ? launch_grid_walk<2>(g, mag, row_end - row_begin, L, itmax, compensated, st)
                          : launch_grid_walk<3>(g, mag, row_end - row_begin, L, itmax, compensated, st);
  if (L.nlenses == 2)
    return launch_ps<2, PS_MAG>(nullptr, g, nullptr, nullptr, nullptr, mag, nullptr, n, L, itmax, compensated, 0, flags, st);
  return launch_ps<3, PS_MAG>(nullptr, g, nullptr, nullptr, nullptr, mag, nullptr, n, L, itmax, compensated, 0, flags, st);
}

}  // extern "C"
