// Kernel family 3: contour-integration extended-source magnification (and the `mag` gate).
//
// Reference behaviour (paths under /root/reference/src/caustics): extended_source.py (all),
// integrate.py, utils.py:15-99, multipole.py, lightcurve.py:24-254.  The reference evaluates one
// source position at a time through ~200 size-1 custom calls and padded XLA scans; here a batch of
// S source positions flows through a short sequence of embarrassingly parallel phases whose state
// lives in HBM as structure-of-arrays with the SOURCE index fastest, so every phase is coalesced:
//
//   k_limb_walk      thread / source        N0 warm-started solves along the limb (:100-107)
//   k_refine_select  thread / source   x10  top-n widest track gaps -> new theta (:114-127)
//   k_refine_solve   thread / (source, new point) x10   warm-started solves (:76-98,127)
//   k_tracks         thread / source        theta order, duplicate guard, greedy track matching
//                                           (:34-53,139-151, utils.py:15-40)
//   k_contours       thread / source        closed/open tracks, splitting, stitching, Green's
//                                           theorem for a uniform disk (:156-727, integrate.py:23-27)
//   k_ld_pq          thread / (source, contour vertex)  Dominik P/Q Gauss-Legendre integrals
//   k_ld_sum         thread / source        trapezoid of P dx + Q dy (integrate.py:47-121)
//   k_gate           thread / point         images + hexadecapole + validity tests + compaction
//
// Every solve is the same Gauss-Seidel Ehrlich-Aberth device function as kernels 1/2 (ea_core.cuh)
// so root labelling along the limb follows the reference's iteration.  Nothing here allocates: the
// caller provides one workspace (caustics_ext_workspace_bytes).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <mutex>

#include "../../include/caustics_b200.h"
#include "extended_core.cuh"
#include "extended_host.h"
#include "nvtx_range.h"
#include "tuning.h"

using namespace cb200;

namespace {

constexpr int NT = 128;
// resident CTAs per SM the register allocation aims at (measured, profiles/r02_ext_variants.txt): the
// one-pass track kernel is fastest with all 255 registers (2 CTAs); the triple-lens refinement solves with 128
// (4 CTAs: 17.0 ms against 17.4 on C4)
#ifndef SW_MINB
#define SW_MINB 2
#endif
// Phase-variant mask (caustics_set_tuning("ext_variants", mask) overrides the rules; tests, experiments):
//    2  stitching on a shared-memory copy of the tracks, one warp per source (track-array path)
//    4  warp-per-source limb-darkened sum
//    8  lane-per-root limb walk (k_limb_walk_group)
//   32  thread-per-source open-track pass instead of the staged one
//   64  the whole-record staged open pass (k_open_staged) instead of the compact one (k_open_compact)
// (1, 16 belonged to earlier pipelines and are ignored.)
// `n` is the number of sources that are integrated (a gated call passes its estimate).
inline int small_mask(int64_t n, int nlenses) {
  const int k = tuning_get(TUNE_EXT_VARIANTS);
  if (k >= 0) return k & 127;
  // measured on one B200 at npts_limb = 200 (batches of 3 000 ... 100 000)
  int m = 0;
  if (n <= (nlenses == 2 ? 16384 : 8192)) m |= 6 | 8;
  return m;
}

// un-gated uniform-disk calls of at least this many sources run as two windows on two streams (split_rule below)
constexpr int64_t SPLIT_MIN = 4096;

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? CAUSTICS_OK : CAUSTICS_ERR_CUDA_BASE + (int)e; }

// ---- kernels: thin wrappers around the phase bodies of extended_core.cuh ------------------------
template <int NL>
__global__ void __launch_bounds__(NT) k_limb_walk(ExtCfg cfg, ExtBuf b, LensConst L) {
  __shared__ EASmem<NL * NL + 1, false, NT> sm;
  limb_walk_body<NL, NT>(cfg, b, L, sm, threadIdx.x, (int64_t)blockIdx.x * NT + threadIdx.x);
}
template <int NL>
__global__ void __launch_bounds__(NT) k_limb_walk_group(ExtCfg cfg, ExtBuf b, LensConst L) {
  __shared__ EASmem<NL * NL + 1, false, NT> sm;
  limb_walk_group_body<NL, NT>(cfg, b, L, sm, threadIdx.x, (int64_t)blockIdx.x * (NT / 32) + threadIdx.x / 32);
}
__global__ void __launch_bounds__(NT) k_limb_walk_single(ExtCfg cfg, ExtBuf b, LensConst L) {
  limb_walk_single_body(cfg, b, L, (int64_t)blockIdx.x * NT + threadIdx.x);
}
// the refinement: a selection kernel + a solve kernel per round
// the selection streams 3.6 KB of state per source and waits on it (long_scoreboard 6.9 warps per issue at 36 warps/SM):
// full occupancy (32 registers) 0.146 ms per launch on C4 against 0.156 at 48 warps and 0.170 at 36
#ifndef SEL_MINB
#define SEL_MINB 16
#endif
template <int D>
__global__ void __launch_bounds__(NT, SEL_MINB) k_round_select(ExtCfg cfg, ExtBuf b, int round) {
  round_select_body<D>(cfg, b, round, threadIdx.x & 31, (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5));
}
template <int NL, bool COMP>
__global__ void __launch_bounds__(NT, (NL == 3 && !COMP) ? 4 : 1) k_round_solve(ExtCfg cfg, ExtBuf b, LensConst L, int round) {
  __shared__ EASmem<(NL == 1 ? 2 : NL * NL + 1), COMP, NT> sm;
  round_solve_body<NL, COMP, NT>(cfg, b, L, sm, threadIdx.x, (int64_t)blockIdx.x * NT + threadIdx.x, round);
}
template <int D>
__global__ void __launch_bounds__(NT) k_tracks(ExtCfg cfg, ExtBuf b) {
  __shared__ double2 stage[D > 5 ? D * NT : 1];
  tracks_body<D>(cfg, b, (int64_t)blockIdx.x * NT + threadIdx.x, stage + (D > 5 ? threadIdx.x : 0), NT);
}
template <int D>
__global__ void __launch_bounds__(NT, SW_MINB) k_sweep(ExtCfg cfg, ExtBuf b) {
  __shared__ double2 stage[D > 5 ? D * NT : 1];
  sweep_body<D>(cfg, b, (int64_t)blockIdx.x * NT + threadIdx.x, stage + (D > 5 ? threadIdx.x : 0), NT);
}
// the sources sweep_body listed: open tracks -> segments -> stitched contours, added to the closed-track sum
template <int D>
__global__ void __launch_bounds__(NT) k_open(ExtCfg cfg, ExtBuf b, LensConst L) {
  const int64_t g = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (g >= *b.open_count) return;
  contours_body<D, false>(cfg, b, L, b.open_list[g], nullptr, true);
}
// The same with one warp per listed source, persistent over the list: the lanes gather the source's tracks
// (through order and permutation; a lane takes a limb point and issues its D image loads together) into
// shared memory, lane i splits track i into segments, lane 0 stitches and integrates on the copy.  The
// listed sources are few (the limb crosses a caustic) and their processing is a chain of dependent reads:
// from shared memory, and spread over every SM instead of count / 128 CTAs.
inline size_t open_stage_bytes(int NP, int D) {
  return (size_t)NP * D * 17 + (size_t)(D * MAXPARTS + MAXSEG) * sizeof(Seg) + 64 + 128 + 64;
}
template <int D>
__global__ void __launch_bounds__(32) k_open_staged(ExtCfg cfg, ExtBuf b, LensConst L) {
  extern __shared__ __align__(16) double stage_mem[];
  const int lane = threadIdx.x, n = cfg.NP * D;
  double* re = stage_mem;
  double* im = re + n;
  Seg* tp = (Seg*)(im + n);
  Seg* parts = tp + D * MAXPARTS;
  int* tn = (int*)(parts + MAXSEG);
  int8_t* chain_mem = (int8_t*)(tn + 16);
  uint8_t* f = (uint8_t*)(chain_mem + 128);
  const int count = *b.open_count;
  for (int g = blockIdx.x; g < count; g += gridDim.x) {
    const int64_t s = b.open_list[g];
    const uint64_t* perm = b.perm + s * cfg.NP;
    for (int pth = lane; pth < cfg.NP; pth += 32) {
      const int slot = b.order[s * cfg.NP + pth];
      const uint64_t pm = perm[pth];
      const uint32_t fw = b.fw[s * cfg.NP + slot];
      const double2* col = b.z + (s * cfg.NP + slot) * D;
      double2 v[D];
#pragma unroll
      for (int t = 0; t < D; ++t) v[t] = col[(int)((pm >> (4 * t)) & 15u)];
#pragma unroll
      for (int t = 0; t < D; ++t) {
        re[pth * D + t] = v[t].x; im[pth * D + t] = v[t].y;
        f[pth * D + t] = (uint8_t)((fw >> (3 * (int)((pm >> (4 * t)) & 15u))) & 7u);
      }
    }
    __syncwarp();
    const TrackStage st{re, im, f, nullptr, D, 1};
    const Tracks T{cfg, b, s, &st, nullptr};
    const int np = build_parts_warp<D>(cfg, T, b.sw_closed[s] & ((1u << D) - 1u), tp, tn, parts, lane);
    if (lane == 0) contours_body<D, false>(cfg, b, L, s, &st, true, parts, np, chain_mem);
    __syncwarp();
  }
}
// The open pass as built at the end of round 2: one warp per listed source, persistent, staging ONLY the tracks that
// hold real images and are not closed (typically 2-4 of a triple lens's 10: the pair born at a fold and whatever the
// matching tangled with it), track-major and renumbered 0..W-1 in track order -- the reference's segment order only
// ever compares track indices, so the renumbering is invisible.  The two scans of a track are spread over the lanes:
// a lane takes a limb point, decides start / end of a run from the point and its predecessor exactly as
// track_parts does (ballots give the run boundaries in increasing order) and stores the chord length to the
// predecessor; a lane per run then adds its run's chord lengths in limb order (the same additions as part_segment's
// serial loop: same bits).  Lane 0 stitches and integrates on the copy.  Sources with wlo < W <= whi are taken by
// this launch (the staging buffer holds whi tracks); a second launch takes the rest.
inline size_t open_compact_bytes(int NP, int whi) {
  return (size_t)NP * whi * 25 + (size_t)(whi * MAXPARTS + MAXSEG) * sizeof(Seg) + (size_t)whi * MAXPARTS * 4 + 128 + 128 + 64;
}
template <int D>
__global__ void __launch_bounds__(32) k_open_compact(ExtCfg cfg, ExtBuf b, LensConst L, int wlo, int whi) {
  extern __shared__ __align__(16) double stage_mem[];
  const int lane = threadIdx.x, NP = cfg.NP, n = NP * whi;
  double* re = stage_mem;
  double* im = re + n;
  double* dl = im + n;                              // dl[t * NP + p] = |pt(t, p) - pt(t, p - 1)|
  Seg* tp = (Seg*)(dl + n);
  Seg* parts = tp + whi * MAXPARTS;
  int16_t* lo = (int16_t*)(parts + MAXSEG);         // [whi][MAXPARTS]
  int16_t* hi = lo + whi * MAXPARTS;
  int* tn = (int*)(hi + whi * MAXPARTS);            // [whi] (<= 16 tracks)
  int8_t* chain_mem = (int8_t*)(tn + 16);           // 128 bytes: the active chain's pieces
  uint8_t* f = (uint8_t*)(chain_mem + 128);
  const int count = *b.open_count;
  for (int g = blockIdx.x; g < count; g += gridDim.x) {
    const int64_t s = b.open_list[g];
    const unsigned work = b.sw_closed[s] >> 16;
    const int W = __popc(work);
    if (W <= wlo || W > whi) continue;              // warp-uniform
    // ---- gather the W tracks through order and permutation
    const uint64_t* perm = b.perm + s * NP;
    for (int pth = lane; pth < NP; pth += 32) {
      const int slot = b.order[s * NP + pth];
      const uint64_t pm = perm[pth];
      const uint32_t fw = b.fw[s * NP + slot];
      const double2* col = b.z + (s * NP + slot) * D;
      unsigned wk = work;
      for (int t = 0; t < W; ++t) {
        const int tr = __ffs(wk) - 1; wk &= wk - 1;
        const int j = (int)((pm >> (4 * tr)) & 15u);
        const double2 v = col[j];
        re[t * NP + pth] = v.x; im[t * NP + pth] = v.y;
        f[t * NP + pth] = (uint8_t)((fw >> (3 * j)) & 7u);
      }
    }
    __syncwarp();
    // ---- runs of real images of one parity without jumps (track_parts, lane per limb point)
    for (int t = 0; t < W; ++t) {
      int np_ = 0, nend = 0;
      for (int base = 0; base <= NP; base += 32) {
        const int p = base + lane;
        bool start = false, end = false;
        if (p <= NP) {
          const int pc = p < NP ? p : NP - 1, pq = p > 0 ? p - 1 : 0;
          const uint8_t fc = f[t * NP + pc], fq = f[t * NP + pq];
          const cd zc = mk(re[t * NP + pc], im[t * NP + pc]), zq = mk(re[t * NP + pq], im[t * NP + pq]);
          const bool real = p < NP && (fc & 1), prev_real = p > 0 && (fq & 1);
          const double par = real ? ((fc & 4) ? 0.0 : ((fc & 2) ? 1.0 : -1.0)) : 0.0;
          const double prev_par = prev_real ? ((fq & 4) ? 0.0 : ((fq & 2) ? 1.0 : -1.0)) : 0.0;
          const cd z = real ? zc : mk(0, 0), prev_z = prev_real ? zq : mk(0, 0);
          if (p == 0) start = real;
          else if (p == NP) end = prev_real;
          else {
            const double dm = (real ? 1.0 : 0.0) - (prev_real ? 1.0 : 0.0);
            const bool change = norm2(z - prev_z) > 0.01 || par != prev_par || dm != 0.0;
            start = change && dm >= 0.0;
            end = change && dm <= 0.0;
            dl[t * NP + p] = sqrt(norm2(zc - zq));
          }
        }
        unsigned sm = __ballot_sync(0xffffffffu, start), em = __ballot_sync(0xffffffffu, end);
        if (lane == 0) {
          while (em && nend < MAXPARTS) { hi[t * MAXPARTS + nend++] = (int16_t)(base + __ffs(em) - 1); em &= em - 1; }
          while (sm && np_ < MAXPARTS) { lo[t * MAXPARTS + np_++] = (int16_t)(base + __ffs(sm) - 1); sm &= sm - 1; }
        }
      }
      if (lane == 0) tn[t] = np_ < nend ? np_ : nend;
    }
    __syncwarp();
    // ---- a lane per run: the segment record (part_segment)
    for (int item = lane; item < W * MAXPARTS; item += 32) {
      const int t = item / MAXPARTS, k = item - t * MAXPARTS;
      if (k >= tn[t]) continue;
      const int a = lo[item], e = hi[item];
      Seg gsg;
      gsg.track = -1; gsg.lo = (int16_t)a; gsg.hi = (int16_t)e; gsg.par = 0; gsg.len = 0.0;
      if (!(e - a < 2 || (a == 0 && e == 0))) {
        gsg.track = (int16_t)t;
        const uint8_t fa = f[t * NP + a];
        gsg.par = (fa & 4) ? 0 : ((fa & 2) ? 1 : -1);
        double len = 0.0;
        for (int p = a + 1; p < e; ++p) len += dl[t * NP + p];
        gsg.len = len;
      }
      tp[item] = gsg;
    }
    __syncwarp();
    if (lane == 0) {
      int np = 0;
      const int nseg_max = 3 * (cfg.nl * cfg.nl + 1);
      for (int i = W - 1; i >= 0 && np < nseg_max; --i)
        for (int k = tn[i] - 1; k >= 0 && np < nseg_max; --k)
          if (tp[i * MAXPARTS + k].track >= 0) parts[np++] = tp[i * MAXPARTS + k];
      const TrackStage st{re, im, f, nullptr, 1, NP};
      contours_body<D, false>(cfg, b, L, s, &st, true, parts, np, chain_mem);
    }
    __syncwarp();
  }
}
// the tangent variant is memory-latency bound (long_scoreboard on top): 5 CTAs/SM (96 registers) 3.19 ms on C4, against
// 3.72 unbounded (164 registers, 3 CTAs), 3.5 with 4, 3.67 with 6, 3.81 with 8; the plain variant does not care
template <int D, bool GRAD>
__global__ void __launch_bounds__(NT, GRAD ? 5 : 1) k_contours(ExtCfg cfg, ExtBuf b, LensConst L) {
  contours_body<D, GRAD>(cfg, b, L, (int64_t)blockIdx.x * NT + threadIdx.x);
}
// small batches: one warp per source; the lanes copy the source's tracks into shared memory, lane 0
// runs the (scalar) stitching logic on the copy
template <int D, bool GRAD>
__global__ void __launch_bounds__(32) k_contours_staged(ExtCfg cfg, ExtBuf b, LensConst L) {
  extern __shared__ double stage_mem[];
  const int64_t s = blockIdx.x;
  if (s >= nsrc(cfg, b)) return;
  const int n = cfg.NP * D;
  double* re = stage_mem;
  double* im = re + n;
  double* th = im + n;
  int8_t* chain_mem = (int8_t*)(th + cfg.NP);
  uint8_t* f = (uint8_t*)(chain_mem + 128);
  for (int k = threadIdx.x; k < n; k += 32) {
    const int64_t g = (int64_t)k * cfg.S + s;    // k = p * D + track, same order as the global planes
    re[k] = b.sre[g]; im[k] = b.sim[g]; f[k] = b.sflg[g];
  }
  if (b.vth)
    for (int pth = threadIdx.x; pth < cfg.NP; pth += 32)
      th[pth] = b.theta[s * cfg.NP + b.order[s * cfg.NP + pth]];
  __syncwarp();
  if (threadIdx.x == 0) {
    const TrackStage st{re, im, f, th, D, 1};
    contours_body<D, GRAD>(cfg, b, L, s, &st, false, nullptr, 0, chain_mem);
  }
}
template <int NL>
__global__ void __launch_bounds__(NT) k_ld_pq(ExtCfg cfg, ExtBuf b, LensConst L) {
  // grid-stride over the (source, vertex) pairs of the sources that are actually integrated (with the
  // gate on, a grid sized for every point of the light curve would start ~95 % of its CTAs only to
  // exit), vertices fastest: a warp takes 32 consecutive vertices of ONE source, so it is either full
  // or past that source's vertex count -- the arithmetic (2 x npts_ld lens-equation evaluations per
  // vertex) dwarfs the strided 16-byte loads.  (VMAX * S < 2^31 is checked by the driver.)
  const unsigned ns = (unsigned)nsrc(cfg, b), vmax = (unsigned)cfg.VMAX;
  const unsigned total = vmax * ns;
  for (unsigned g = blockIdx.x * NT + threadIdx.x; g < total; g += gridDim.x * NT) {
    const unsigned s = g / vmax;
    ld_pq_item<NL>(cfg, b, L, (int)(g - s * vmax), (int64_t)s);
  }
}
__global__ void __launch_bounds__(NT) k_ld_sum(ExtCfg cfg, ExtBuf b) {
  ld_sum_body(cfg, b, (int64_t)blockIdx.x * NT + threadIdx.x);
}
// small batches: one warp per source; lanes share the edges of each contour (trapezoid of P dx + Q dy)
__global__ void __launch_bounds__(NT) k_ld_sum_warp(ExtCfg cfg, ExtBuf b) {
  const int64_t s = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (s >= nsrc(cfg, b)) return;   // warp-uniform
  const int lane = threadIdx.x & 31;
  const int nc = b.ncont[s];
  double total = 0.0;
  for (int c = 0; c < nc; ++c) {
    const int v0 = b.cstart[(int64_t)c * cfg.S + s], v1 = b.cstart[(int64_t)(c + 1) * cfg.S + s];
    double acc = 0.0;
    for (int v = v0 + lane; v + 1 < v1; v += 32) {
      const double2 za = b.vz[(int64_t)v * cfg.S + s], zb = b.vz[(int64_t)(v + 1) * cfg.S + s];
      const double Pa = b.vP[(int64_t)v * cfg.S + s], Pb = b.vP[(int64_t)(v + 1) * cfg.S + s];
      const double Qa = b.vQ[(int64_t)v * cfg.S + s], Qb = b.vQ[(int64_t)(v + 1) * cfg.S + s];
      acc += 0.5 * (Pa + Pb) * (zb.x - za.x) + 0.5 * (Qa + Qb) * (zb.y - za.y);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    total += acc * b.cpar[(int64_t)c * cfg.S + s];
  }
  if (lane == 0) {
    b.mag[src_index(b, s)] = fabs(total) / (3.14159265358979323846 * cfg.rho * cfg.rho);
  }
}
template <bool COMP>
__global__ void __launch_bounds__(NT) k_gate(const double2* w_in, double* mag, uint8_t* test_out, int32_t* list,
                                              int32_t* count, int64_t n, LensConst L, double rho, double q, int itmax) {
  __shared__ EASmem<5, COMP, NT> sm;
  gate_body<COMP, NT>(w_in, mag, test_out, list, count, n, L, rho, q, itmax, sm, threadIdx.x,
                      (int64_t)blockIdx.x * NT + threadIdx.x);
}

// a piece of a host-computed table passed by value (kernel arguments are copied at enqueue / capture)
struct TabChunk { static constexpr int N = 256; double v[N]; };
__global__ void k_store_table(TabChunk c, double* dst, int m) {
  if ((int)threadIdx.x < m) dst[threadIdx.x] = c.v[threadIdx.x];
}

__global__ void k_jitter_table(int D, int nadd, double* out) {
  jitter_table_body(D, nadd, out, (int)(blockIdx.x * blockDim.x + threadIdx.x));
}

// fills list = 0..n-1 and count = n (nlenses != 2: full integration everywhere, lightcurve.py:226-227)
__global__ void k_iota(int32_t* list, int32_t* count, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) list[i] = (int32_t)i;
  if (i == 0) *count = (int32_t)n;
}

template <int NL>
int run_pipeline(const ExtCfg& cfg, ExtBuf b, const LensConst& L, cudaStream_t st) {
  constexpr int D = NL == 1 ? 2 : NL * NL + 1;
  const unsigned gs = (unsigned)((cfg.S + NT - 1) / NT);
  if (NL != 1 && b.list_off == 0)      // the same table serves every window of a gated call
    k_jitter_table<<<(D * cfg.nadd + 127) / 128, 128, 0, st>>>(D, cfg.nadd, const_cast<double*>(b.jit));
  if (NL == 1) k_limb_walk_single<<<gs, NT, 0, st>>>(cfg, b, L);
  else if (cfg.small & 8) {
    constexpr int G = 32 / (NL == 1 ? 5 : NL * NL + 1);   // sources per warp
    const int64_t warps = (cfg.S + G - 1) / G;
    k_limb_walk_group<(NL == 1 ? 2 : NL)><<<(unsigned)((warps + NT / 32 - 1) / (NT / 32)), NT, 0, st>>>(cfg, b, L);
  } else k_limb_walk<(NL == 1 ? 2 : NL)><<<gs, NT, 0, st>>>(cfg, b, L);
  {
    const unsigned gsel = (unsigned)((cfg.S + NT / 32 - 1) / (NT / 32));
    const unsigned gsol = (unsigned)((cfg.S * cfg.nadd + NT - 1) / NT);
    for (int round = 0; round < NITER; ++round) {
      k_round_select<D><<<gsel, NT, 0, st>>>(cfg, b, round);
      if (cfg.comp && NL != 1) k_round_solve<NL, true><<<gsol, NT, 0, st>>>(cfg, b, L, round);
      else k_round_solve<NL, false><<<gsol, NT, 0, st>>>(cfg, b, L, round);
    }
  }
  if (!cfg.ld && !cfg.tracks) {
    // plain uniform-disk magnification: one pass (matching + closed tracks), then the caustic-crossing sources
    cudaError_t e = cudaMemsetAsync(b.open_count, 0, 4, st);
    if (e != cudaSuccess) return cuda_rc(e);
    k_sweep<D><<<gs, NT, 0, st>>>(cfg, b);
    const size_t open_bytes = open_stage_bytes(cfg.NP, D);
    const int tw = tuning_get(TUNE_OPEN_WSMALL);
    const int wsmall = tw > 0 ? (tw < D ? tw : D) : 2;
    if (!(cfg.small & 64) && open_compact_bytes(cfg.NP, D) <= 200 * 1024) {
      // sources with at most wsmall tracks to look at first (small staging buffer, many warps per SM), then the rest
      for (int pass = 0; pass < (wsmall < D ? 2 : 1); ++pass) {
        const int wlo = pass == 0 ? 0 : wsmall, whi = (pass == 0 && wsmall < D) ? wsmall : D;
        const size_t nb = open_compact_bytes(cfg.NP, whi);
        if (nb > 48 * 1024) cudaFuncSetAttribute(k_open_compact<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)open_compact_bytes(cfg.NP, D));
        const int64_t per_sm = (int64_t)(220 * 1024 / (nb + 1024));
        const int64_t slots = 148 * (per_sm < 1 ? 1 : (per_sm > 24 ? 24 : per_sm));
        k_open_compact<D><<<(unsigned)(cfg.S < slots ? cfg.S : slots), 32, nb, st>>>(cfg, b, L, wlo, whi);
      }
    } else if (open_bytes <= 200 * 1024 && !(cfg.small & 32)) {
      if (open_bytes > 48 * 1024) cudaFuncSetAttribute(k_open_staged<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)open_bytes);
      const int64_t per_sm = (int64_t)(220 * 1024 / (open_bytes + 1024));
      const int64_t slots = 148 * (per_sm < 1 ? 1 : (per_sm > 16 ? 16 : per_sm));
      k_open_staged<D><<<(unsigned)(cfg.S < slots ? cfg.S : slots), 32, open_bytes, st>>>(cfg, b, L);
    } else {
      k_open<D><<<gs, NT, 0, st>>>(cfg, b, L);
    }
    return cuda_rc(cudaGetLastError());
  }
  k_tracks<D><<<gs, NT, 0, st>>>(cfg, b);
  const size_t stage_bytes = (size_t)cfg.NP * D * 17 + (size_t)cfg.NP * 8 + 128 + 64;
  if ((cfg.small & 2) && stage_bytes <= 200 * 1024) {
    if (stage_bytes > 48 * 1024) {
      cudaFuncSetAttribute(k_contours_staged<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
      cudaFuncSetAttribute(k_contours_staged<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
    }
    if (b.grad) k_contours_staged<D, true><<<(unsigned)cfg.S, 32, stage_bytes, st>>>(cfg, b, L);
    else k_contours_staged<D, false><<<(unsigned)cfg.S, 32, stage_bytes, st>>>(cfg, b, L);
  } else {
    if (b.grad) k_contours<D, true><<<gs, NT, 0, st>>>(cfg, b, L);
    else k_contours<D, false><<<gs, NT, 0, st>>>(cfg, b, L);
  }
  if (cfg.ld) {
    const int64_t gv_all = ((int64_t)cfg.VMAX * cfg.S + NT - 1) / NT;
    const unsigned gv = (unsigned)(gv_all < 148 * 64 ? gv_all : 148 * 64);
    k_ld_pq<NL><<<gv, NT, 0, st>>>(cfg, b, L);
    if (cfg.small & 4) k_ld_sum_warp<<<(unsigned)((cfg.S + NT / 32 - 1) / (NT / 32)), NT, 0, st>>>(cfg, b);
    else k_ld_sum<<<gs, NT, 0, st>>>(cfg, b);
  }
  return cuda_rc(cudaGetLastError());
}

}  // namespace

// defined in kernels.cu (anonymous-namespace helper re-exported for this TU)
extern "C" int caustics_internal_lens_const(const caustics_lens* lens, void* out);

extern "C" {

size_t caustics_ext_workspace_bytes(int64_t n, int nlenses, int npts_limb, int limb_darkening, int npts_ld) {
  ExtCfg c;
  if (make_cfg(n, 1.0, nlenses, npts_limb, limb_darkening & 3, 0.0, npts_ld, 1, 0, &c)) return 0;
  // a uniform-disk workspace serves the tangent and contour-export entry points too (they keep the image
  // tracks as arrays) unless the caller says it will only ask for magnifications
  c.tracks = (limb_darkening & CAUSTICS_WS_MAG_ONLY) ? 0 : 1;
  // + the alignment slack of a second window (ext_driver cuts large un-gated calls in two)
  return make_layout(c).total + (n >= SPLIT_MIN ? 64 * 1024 : 0);
}

size_t caustics_mag_workspace_bytes(int64_t n, int64_t max_full, int nlenses, int npts_limb, int limb_darkening,
                                    int npts_ld) {
  if (max_full > n) max_full = n;
  if (max_full < 1) max_full = 1;
  ExtCfg c;
  if (n < 0 || make_cfg(max_full, 1.0, nlenses, npts_limb, limb_darkening, 0.0, npts_ld, 1, 0, &c)) return 0;
  return make_layout(c, n).total;
}

}  // extern "C"

namespace {

// The largest number of sources (<= n) whose per-source arrays fit `bytes` next to a compact list of n
// points; 0 when not even one fits.  The layout is monotone in S, so bisect.
int64_t capacity_for(const ExtCfg& proto, int64_t n, size_t bytes) {
  auto fits = [&](int64_t S) { ExtCfg c = proto; c.S = S; return make_layout(c, n).total <= bytes; };
  if (fits(n)) return n;
  if (!fits(1)) return 0;
  int64_t lo = 1, hi = n;               // fits(lo), !fits(hi)
  while (hi - lo > 1) { const int64_t mid = lo + (hi - lo) / 2; (fits(mid) ? lo : hi) = mid; }
  return lo;
}

// Two windows of one un-gated call on two streams.  The phases of the pipeline are unlike -- FP64-bound solves, a
// memory-bound selection, latency-bound matching and stitching -- and a single stream runs them one after the other, each
// with its own idle resource (and the limb walk of 10^5 triple-lens sources with a second wave that fills 3/4 of the
// machine).  Cut into two windows with workspaces of their own, on the caller's stream and on a side stream (fork / join
// by events, so the call stays stream-ordered for the caller and can be captured in a CUDA graph), the windows' phases
// overlap.  Sources are independent: the results are bit for bit those of the single window.
constexpr int SIDE_MAXDEV = 16, SIDE_MAXWIN = 4;
struct Side {
  std::mutex mu; bool init = false;
  cudaStream_t st[SIDE_MAXWIN - 1] = {}; cudaEvent_t fork = nullptr, join[SIDE_MAXWIN - 1] = {};
};
Side g_side[SIDE_MAXDEV];
// The windows of an un-gated call: *first = length of the first one (on the caller's stream), the others share the
// rest equally.  Returns their number (1 = one window).  Rule (measured, profiles/r02_ext_variants.txt): uniform disk,
// binary or triple lens, at least 4 096 sources -> two windows, the first max(n / 2, min(65 536, n - 16 384)) long;
// three and four windows are no better.  caustics_set_tuning("ext_split", nA) / ("ext_windows", K) override.
inline int split_rule(int64_t n, const ExtCfg& cfg, int gate, int64_t* first) {
  if (gate || cfg.ld || cfg.nl == 1) return 1;
  const int t = tuning_get(TUNE_EXT_SPLIT), kw = tuning_get(TUNE_EXT_WINDOWS);
  int K = 2;
  if (kw >= 1) K = kw < SIDE_MAXWIN ? kw : SIDE_MAXWIN;
  else if (t == 0 || (t < 0 && n < SPLIT_MIN)) K = 1;
  if (K == 1 || n < 2 * K * NT) return 1;
  int64_t f = (n + K - 1) / K;
  if (K == 2 && t < 0 && kw < 1) {
    const int64_t g = n - 16384 < 65536 ? n - 16384 : 65536;
    if (g > f) f = g;
  }
  if (t > 0 && t < n) f = t;
  *first = ((f + NT - 1) / NT) * NT;
  return *first < n ? K : 1;
}

// Shared driver.
//   gate == 0  every point of w gets the full contour integration (mag_extended_source)
//   gate == 1  lightcurve.py dispatch: hexadecapole where the reference's tests pass, else integration
//   gate == 2  the caller supplies the compact list of points to integrate (ext_list, ext_count on the
//              device, at most n entries): second half of the two-call form caustics_mag_gate -> here
// Gated calls integrate at most `cap` sources at a time, cap = what the workspace holds; with cap < n the
// integration phases are enqueued ceil(n / cap) times over consecutive windows of the list and a window
// past the device-side count exits at once (the count never visits the host).
int ext_driver(const void* w, double* mag, double* grad, uint8_t* test_out, int64_t n, double rho, const caustics_lens* lens,
               double q_for_gate, int gate, const int32_t* ext_list, const int32_t* ext_count, int npts_limb,
               int limb_darkening, double u1, int npts_ld, int itmax, int compensated, void* workspace,
               size_t workspace_bytes, void* stream) {
  if (!lens || n < 0) return CAUSTICS_ERR_BAD_ARG;
  ExtCfg cfg;
  int rc = make_cfg(n, rho, lens->nlenses, npts_limb, limb_darkening, u1, npts_ld, itmax, compensated, &cfg);
  if (rc) return rc;
  if (n == 0) return CAUSTICS_OK;
  if (!w || !mag || !workspace) return CAUSTICS_ERR_BAD_ARG;
  if (gate == 2 && (!ext_list || !ext_count)) return CAUSTICS_ERR_BAD_ARG;
  if (n > 0x7fffffffLL) return CAUSTICS_ERR_BAD_ARG;
  const int64_t idx_cap = 0x7fffffffLL / (cfg.VMAX > NADD_MAX ? cfg.VMAX : NADD_MAX);   // index range of one pass
  cfg.tracks = grad ? 1 : 0;
  int64_t cap = n;
  if (gate) {
    cap = capacity_for(cfg, n, workspace_bytes);
    if (cap > idx_cap) cap = idx_cap;
    if (cap < 1) return CAUSTICS_ERR_BAD_ARG;
  } else if (n > idx_cap) return CAUSTICS_ERR_BAD_ARG;
  cfg.S = cap;
  cfg.tracks = grad ? 1 : 0;
  const Layout lay = make_layout(cfg, gate ? n : -1);
  if (workspace_bytes < lay.total) return CAUSTICS_ERR_BAD_ARG;
  // Phase variants by the number of sources that are integrated.  A binary-lens gate sends a few per
  // cent of a light curve to the integration (measured rule of round 1: the small-batch bounds apply to
  // 16x the point count); a caller-supplied list is sized by its capacity.
  const int64_t n_int = gate == 1 && lens->nlenses == 2 ? (n + 15) / 16 : (cap < n ? cap : n);
  cfg.small = small_mask(n_int, lens->nlenses);
  LensConst L;
  memset(&L, 0, sizeof(L));
  if (lens->nlenses == 1) { L.nlenses = 1; L.x_cm = 0.0; L.eps[0] = 1.0; }
  else if ((rc = caustics_internal_lens_const(lens, &L))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ExtBuf b = bind(cfg, lay, workspace);
  b.w = (const double2*)w;
  b.mag = mag;
  b.grad = cfg.ld ? nullptr : grad;
  cfg.ngrad_stride = n;
  char* base = (char*)workspace;
  int32_t* list = (int32_t*)(base + lay.list);
  int32_t* count = (int32_t*)(base + lay.count);
  if (cfg.ld) {
    // Gauss-Legendre tables: computed on the host and handed to the device as KERNEL ARGUMENTS, 256
    // doubles per launch.  (A cudaMemcpyAsync from this stack buffer would be fine eagerly, but a
    // captured CUDA graph would replay the copy from a dead host address.)
    double tab[2 * (2048 + 1024 + 512 + 8)];
    const int nn = fill_gl_tables(cfg.n1, cfg.n2, tab);
    for (int off = 0; off < 2 * nn; off += TabChunk::N) {
      TabChunk c;
      const int m = 2 * nn - off < TabChunk::N ? 2 * nn - off : TabChunk::N;
      memcpy(c.v, tab + off, (size_t)m * sizeof(double));
      k_store_table<<<1, TabChunk::N, 0, st>>>(c, (double*)(base + lay.gl) + off, m);
    }
  }
  if (gate == 1 && lens->nlenses == 2) {
    cudaError_t e = cudaMemsetAsync(count, 0, 4, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (compensated)
      k_gate<true><<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>((const double2*)w, mag, test_out, list, count, n, L, rho, q_for_gate, itmax);
    else
      k_gate<false><<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>((const double2*)w, mag, test_out, list, count, n, L, rho, q_for_gate, itmax);
    b.list = list; b.count = count;
  } else if (gate == 1) {
    k_iota<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(list, count, n);
    if (test_out) { cudaError_t e = cudaMemsetAsync(test_out, 0, (size_t)n, st); if (e != cudaSuccess) return cuda_rc(e); }
    b.list = list; b.count = count;
  } else if (gate == 2) {
    b.list = ext_list; b.count = ext_count;
  }
  int64_t first = 0;
  const int K = split_rule(n, cfg, gate, &first);
  if (K > 1) {
    int64_t off[SIDE_MAXWIN + 1];
    off[0] = 0; off[1] = first;
    const int64_t rest = (((n - first) / (K - 1) + NT - 1) / NT) * NT;
    for (int i = 2; i <= K; ++i) off[i] = off[i - 1] + rest < n ? off[i - 1] + rest : n;
    off[K] = n;
    ExtCfg cw[SIDE_MAXWIN]; Layout lw[SIDE_MAXWIN];
    size_t need = 0;
    bool ok = true;
    for (int i = 0; i < K; ++i) {
      if (off[i + 1] <= off[i]) ok = false;
      cw[i] = cfg; cw[i].S = ok ? off[i + 1] - off[i] : 1;
      lw[i] = make_layout(cw[i]);
      need += lw[i].total;
    }
    int dev = 0;
    if (ok && need <= workspace_bytes && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < SIDE_MAXDEV) {
      Side& sd = g_side[dev];
      std::lock_guard<std::mutex> lk(sd.mu);
      if (!sd.init) {
        bool good = cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < SIDE_MAXWIN - 1 && good; ++i)
          good = cudaStreamCreateWithFlags(&sd.st[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&sd.join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!good) return cuda_rc(cudaGetLastError());
        sd.init = true;
      }
      cudaError_t e = cudaEventRecord(sd.fork, st);
      for (int i = 1; i < K && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(sd.st[i - 1], sd.fork, 0);
      if (e != cudaSuccess) return cuda_rc(e);
      size_t at = 0;
      for (int i = 0; i < K; ++i) {
        ExtBuf bw = bind(cw[i], lw[i], base + at);
        at += lw[i].total;
        bw.w = (const double2*)w + off[i]; bw.mag = mag + off[i]; bw.grad = grad ? grad + off[i] : nullptr;
        cudaStream_t si = i == 0 ? st : sd.st[i - 1];
        const int ri = lens->nlenses == 2 ? run_pipeline<2>(cw[i], bw, L, si) : run_pipeline<3>(cw[i], bw, L, si);
        if (ri && !rc) rc = ri;
      }
      // join whatever happened, so that nothing of this call is left running beside the caller's stream
      for (int i = 1; i < K; ++i) {
        cudaError_t ej = cudaEventRecord(sd.join[i - 1], sd.st[i - 1]);
        if (ej == cudaSuccess) ej = cudaStreamWaitEvent(st, sd.join[i - 1], 0);
        if (ej != cudaSuccess && e == cudaSuccess) e = ej;
      }
      if (rc) return rc;
      return e == cudaSuccess ? CAUSTICS_OK : cuda_rc(e);
    }
  }
  for (int64_t off = 0; off < n; off += cap) {
    b.list_off = off;
    switch (lens->nlenses) {
      case 1: rc = run_pipeline<1>(cfg, b, L, st); break;
      case 2: rc = run_pipeline<2>(cfg, b, L, st); break;
      default: rc = run_pipeline<3>(cfg, b, L, st); break;
    }
    if (rc || !gate) break;
  }
  return rc;
}

}  // namespace

extern "C" {

int caustics_mag_extended_source(const void* w, double* mag, int64_t n, double rho, const caustics_lens* lens,
                                 int npts_limb, int limb_darkening, double u1, int npts_ld, int itmax,
                                 int compensated, void* workspace, size_t workspace_bytes, void* stream) {
  CB200_NVTX("caustics_mag_extended_source");
  return ext_driver(w, mag, nullptr, nullptr, n, rho, lens, 0.0, 0, nullptr, nullptr, npts_limb, limb_darkening, u1, npts_ld,
                    itmax, compensated, workspace, workspace_bytes, stream);
}

int caustics_mag_extended_source_grad(const void* w, double* mag, double* grad, int64_t n, double rho,
                                      const caustics_lens* lens, int npts_limb, int itmax, int compensated,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  CB200_NVTX("caustics_mag_extended_source_grad");
  if (!grad) return CAUSTICS_ERR_BAD_ARG;
  return ext_driver(w, mag, grad, nullptr, n, rho, lens, 0.0, 0, nullptr, nullptr, npts_limb, 0, 0.0, 100, itmax,
                    compensated, workspace, workspace_bytes, stream);
}

int caustics_mag_extended_source_list(const void* w, double* mag, const int32_t* list, const int32_t* count,
                                      int64_t max_count, double rho, const caustics_lens* lens, int npts_limb,
                                      int limb_darkening, double u1, int npts_ld, int itmax, int compensated,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  CB200_NVTX("caustics_mag_extended_source_list");
  return ext_driver(w, mag, nullptr, nullptr, max_count, rho, lens, 0.0, 2, list, count, npts_limb, limb_darkening, u1,
                    npts_ld, itmax, compensated, workspace, workspace_bytes, stream);
}

int caustics_ext_contour_capacity(int nlenses, int npts_limb, int* vmax, int* cmax) {
  ExtCfg c;
  const int rc = make_cfg(1, 1.0, nlenses, npts_limb, 0, 0.0, 100, 1, 0, &c);
  if (rc) return rc;
  if (vmax) *vmax = c.VMAX;
  if (cmax) *cmax = c.CMAX;
  return CAUSTICS_OK;
}

int caustics_ext_contours(const void* w, double* mag, int64_t n, double rho, const caustics_lens* lens,
                          int npts_limb, int itmax, int compensated, void* workspace, size_t workspace_bytes,
                          void* vz, double* vtheta, uint8_t* vcid, int32_t* vcount, double* cpar,
                          int32_t* cstart, int32_t* ncont, void* stream) {
  CB200_NVTX("caustics_ext_contours");
  if (!lens || n < 0) return CAUSTICS_ERR_BAD_ARG;
  ExtCfg cfg;
  int rc = make_cfg(n, rho, lens->nlenses, npts_limb, 0, 0.0, 100, itmax, compensated, &cfg);
  if (rc) return rc;
  if (n == 0) return CAUSTICS_OK;
  cfg.small = small_mask(n, lens->nlenses);
  cfg.tracks = 1;
  if (n > 0x7fffffffLL / cfg.VMAX) return CAUSTICS_ERR_BAD_ARG;
  if (!w || !workspace || !vz || !vtheta || !vcid || !vcount || !cpar || !cstart || !ncont) return CAUSTICS_ERR_BAD_ARG;
  const Layout lay = make_layout(cfg);
  if (workspace_bytes < lay.total) return CAUSTICS_ERR_BAD_ARG;
  LensConst L;
  memset(&L, 0, sizeof(L));
  if (lens->nlenses == 1) { L.nlenses = 1; L.eps[0] = 1.0; }
  else if ((rc = caustics_internal_lens_const(lens, &L))) return rc;
  cfg.emit = 1;
  ExtBuf b = bind(cfg, lay, workspace);
  b.w = (const double2*)w;
  b.mag = mag;
  b.vz = (double2*)vz; b.vth = vtheta; b.vcid = vcid; b.vcount = vcount;
  b.cpar = cpar; b.cstart = cstart; b.ncont = ncont; b.cz0 = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  switch (lens->nlenses) {
    case 1: return run_pipeline<1>(cfg, b, L, st);
    case 2: return run_pipeline<2>(cfg, b, L, st);
    default: return run_pipeline<3>(cfg, b, L, st);
  }
}

int caustics_mag_gate(const void* w, double* mag, uint8_t* used_hexadecapole, int32_t* list, int32_t* count,
                      int64_t n, double rho, const caustics_lens* lens, double q, int itmax, int compensated,
                      void* stream) {
  CB200_NVTX("caustics_mag_gate");
  if (!lens || lens->nlenses != 2 || n < 0 || !(rho > 0.0) || itmax < 0) return CAUSTICS_ERR_BAD_ARG;
  if (n == 0) return CAUSTICS_OK;
  if (n > 0x7fffffffLL || !w || !mag || !list || !count) return CAUSTICS_ERR_BAD_ARG;
  LensConst L;
  int rc = caustics_internal_lens_const(lens, &L);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(count, 0, 4, st);
  if (e != cudaSuccess) return cuda_rc(e);
  if (compensated)
    k_gate<true><<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>((const double2*)w, mag, used_hexadecapole, list, count, n, L, rho, q, itmax);
  else
    k_gate<false><<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>((const double2*)w, mag, used_hexadecapole, list, count, n, L, rho, q, itmax);
  return cuda_rc(cudaGetLastError());
}

int caustics_mag(const void* w, double* mag, uint8_t* used_hexadecapole, int64_t n, double rho,
                 const caustics_lens* lens, double q, int npts_limb, int limb_darkening, double u1, int npts_ld,
                 int itmax, int compensated, void* workspace, size_t workspace_bytes, void* stream) {
  CB200_NVTX("caustics_mag");
  return ext_driver(w, mag, nullptr, used_hexadecapole, n, rho, lens, q, 1, nullptr, nullptr, npts_limb, limb_darkening, u1,
                    npts_ld, itmax, compensated, workspace, workspace_bytes, stream);
}

}  // extern "C"
