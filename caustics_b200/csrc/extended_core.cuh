// Device code of kernel family 3 (extended-source magnification).  See extended.cu for the phase
// structure; every phase body is a function of the source index `s` (and, for the solver phases, of
// the thread's lane in the CTA-wide shared root planes), so the same code is driven by the CUDA
// kernels in extended.cu and -- with one "lane" -- by the host logic tests (tests/hostsim).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "ea_core.cuh"
#include "jax_prng.cuh"
#include "lens_core.cuh"
#include "multipole.cuh"

// adaptive far-panel rule (two_panel): lengths in source radii below which the quarter / half order is used.
// Measured on the golden sets (host build, npts_ld = 100): (0, 8) deviates <= 7e-5 from the reference's rule,
// (2, 8) without the distance condition 4e-4, (0, 16) 2e-4 -- the integrand has square-root kinks wherever the
// integration line leaves or enters an image, so low orders are only safe on short panels.
#ifndef CB200_ADAPT_Q
#define CB200_ADAPT_Q 0.0
#define CB200_ADAPT_H 8.0
#endif
#ifndef CB200_EXT_STRAIGHT
#define CB200_EXT_STRAIGHT 0   // warm-started limb solves: the branchy step (see ea_solve_thread)
#endif

namespace cb200 {

#ifdef CB200_HOSTSIM
struct cb200_d2 { double x, y; };
__device__ __forceinline__ cb200_d2 make_cb200_d2(double x, double y) { cb200_d2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ int cb200_atomic_inc(int32_t* p) { return (*p)++; }
__device__ __forceinline__ double rsqrt(double x) { return 1.0 / sqrt(x); }
#else
typedef double2 cb200_d2;
__device__ __forceinline__ cb200_d2 make_cb200_d2(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ int cb200_atomic_inc(int32_t* p) { return atomicAdd(p, 1); }
#endif

constexpr int NADD_MAX = 64;       // new limb points per refinement round (npts_limb <= 1280)
constexpr int NITER = 10;          // refinement rounds, extended_source.py:71
constexpr int MAXSEG = 30;         // 3 * (nlenses^2 + 1), extended_source.py:279-280
constexpr int MAXPARTS = 10;       // per track, extended_source.py:202-203
// warm-start and duplicate jitters: the reference's own values, jax_prng.cuh (limb_jitter, duplicate_jitter)

struct ExtCfg {
  int nl, D, N0, nadd, NP;
  double rho;
  int itmax, comp, ld, n1, n2, VMAX, CMAX;
  int ld_adapt;     // opt-in: fewer Gauss-Legendre nodes on short far panels of the P/Q integrals (two_panel)
  int small;        // mask of small-batch phase variants: 1 warp select, 2 staged contours, 4 warp LD sum, 8 lane-per-root walk
  int tracks;       // 1: the theta-ordered track arrays are materialised (tangent / contour export; always with ld)
  int emit;         // 1: k_contours also writes the contour vertex lists (z, theta, contour id) for the host
  double u1;
  int64_t S;        // capacity (stride) of the source axis
  int64_t ngrad_stride;   // number of points of the call (row length of ExtBuf::grad)
};

struct ExtBuf {
  const cb200_d2* w;       // source positions (all points of the call)
  const int32_t* list;     // optional: indices into w of the sources to integrate (else identity)
  const int32_t* count;    // optional: device-side number of listed sources (else S)
  int64_t list_off;        // this pass integrates list[list_off .. list_off + S) (gated calls with a small workspace)
  // Limb points by ARRIVAL SLOT, one contiguous record per source (a refinement solve reads and writes whole
  // 16 D-byte columns of it):
  cb200_d2* z;             // [S][NP][D] images of the limb point in arrival slot `slot`
  uint32_t* fw;            // [S][NP]    3 flag bits per image: bit0 real image, bit1 det J > 0, bit2 det J == 0
  double* theta;           // [S][NP]    limb angle by arrival slot
  uint16_t* order;         // [S][NP]    arrival slot of the p-th point in theta order
  // Refinement state (round_select_body / round_solve_body):
  double* rdval;           // [S][NP]    squared width of the interval that starts at this slot (aliases perm / sre)
  uint16_t* rnext;         // [S][NP]    theta-order links
  uint16_t* rlr;           // [S][2][NADD_MAX]  left / right slots of the round's new points
  // Theta-ordered image tracks, structure of arrays with the source index fastest (thread-per-source phases):
  double* sre; double* sim; uint8_t* sflg;  // [NP][D][S], rows = image tracks
  // One-pass uniform-disk path (sweep_body / open_body):
  uint64_t* perm;          // [S][NP]    4 bits per track: which image of limb point p (theta order) the track took
  double* sw_total;        // [S]        signed area of the closed tracks
  uint32_t* sw_closed;     // [S]        mask of closed tracks | mask of tracks with real images that are not closed << 16
  int32_t* open_list;      // [S]        sources (slots of this pass) whose open tracks still have to be stitched
  int32_t* open_count;     //            their number
  cb200_d2* vz; double* vP; double* vQ; uint8_t* vcid; double* vth;   // vth: optional theta per vertex   // [VMAX][S] limb-darkening vertex lists
  int32_t* vcount;         // [S]
  int32_t* ncont;          // [S] number of contours emitted
  cb200_d2* cz0; double* cpar; int32_t* cstart;   // [CMAX(+1)][S] per contour: centroid, parity, first vertex
  const double* glx; const double* glw;          // Gauss-Legendre nodes/weights, n1 then n2
  const double* jit;       // [D][nadd][2] the reference's warm-start jitters (jitter_table_body), same for every source
  double* mag;             // result, indexed like w (through list)
  double* grad;            // optional (uniform disk): [NGRAD][cfg.ngrad_stride] d mag / d(a, e1, e2, Re r3, Im r3, Re w, Im w, rho)
};

__device__ __forceinline__ int64_t nsrc(const ExtCfg& c, const ExtBuf& b) {
  if (!b.count) return c.S;
  const int64_t left = (int64_t)*b.count - b.list_off;
  return left < 0 ? 0 : (left < c.S ? left : c.S);
}
__device__ __forceinline__ int64_t src_index(const ExtBuf& b, int64_t s) {
  return b.list ? (int64_t)b.list[b.list_off + s] : s;
}
#define IS(p, s) ((int64_t)(s) * cfg.NP + (p))                          // per-source records by slot / position
#define IZ(slot, j, s) (((int64_t)(s) * cfg.NP + (slot)) * cfg.D + (j))
#define IT(p, i, s) (((int64_t)(p) * cfg.D + (i)) * cfg.S + (s))        // tracks

__device__ __forceinline__ uint32_t flag_bits(bool real_image, double detj) {
  return (real_image ? 1u : 0u) | (detj > 0 ? 2u : 0u) | (detj == 0 ? 4u : 0u);
}

__device__ __forceinline__ cd source_centre(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L, int64_t s) {
  const cb200_d2 v = b.w[src_index(b, s)];
  return mk(v.x + L.x_cm, v.y);
}
__device__ __forceinline__ cd limb_point(cd w0, double rho, double theta) {
  double sn, cs;
  sincos(theta, &sn, &cs);
  return mk(fma(rho, cs, w0.re), fma(rho, sn, w0.im));
}
__device__ __forceinline__ double theta_init(int k, int N0) {
  // linspace(-pi, pi, N0-1, endpoint=False) ++ [pi - 1e-8], extended_source.py:101-102
  const double pi = 3.14159265358979323846;
  return k < N0 - 1 ? fma((double)k, (2.0 * pi) / (double)(N0 - 1), -pi) : pi - 1e-8;
}

// single lens: the two images analytically (point_source.py:1665-1672)
__device__ __forceinline__ void single_lens_images(cd w, cd (&z)[2]) {
  const double sq = sqrt(1.0 + 4.0 / norm2(w));
  z[0] = (0.5 * (1.0 + sq)) * w;
  z[1] = (0.5 * (1.0 - sq)) * w;
}

// Solve the lens polynomial at w (warm start when `warm`: roots already in the shared planes) and
// write the images and their flag word to arrival slot `slot`.
template <int NL, bool COMP, int NT>
__device__ __forceinline__ uint32_t solve_and_store(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L,
                                                EASmem<NL * NL + 1, COMP, NT>& sm, int tid, bool active,
                                                cd w, bool warm, int slot, int64_t s) {
  constexpr int D = NL * NL + 1;
  cd p[D + 1];
  lens_poly<NL>(L, w, p);
  ea_normalise<D>(p);
  // Coinciding warm-start values make an Aberth term 0 * inf = NaN, and a NaN root passes the stopping
  // test (|h| > eps b is false): the image would silently vanish.  Such a solve is redone from the default
  // initial estimates, as the point-source walks do (ps_walk.cuh).  One copy of the solver in the code (the
  // kernels that walk a limb are bound by instruction fetch, not by arithmetic): the redo is a second trip
  // through the same loop, taken on a warp vote.
  bool todo = active;
#pragma unroll 1
  for (int attempt = 0; attempt < 2; ++attempt) {
    ea_solve_thread<D, COMP, NT, CB200_EXT_STRAIGHT != 0>(p, sm, tid, todo, warm && attempt == 0, EA_INIT_REFERENCE, cfg.itmax);
    if (!warm || attempt == 1) break;
    double chk = 0.0;
    if (active) {
#pragma unroll 1
      for (int j = 0; j < D; ++j) chk += fabs(sm.zre[j][tid]) + fabs(sm.zim[j][tid]);
    }
    todo = active && !(chk < 1e300);
#ifndef CB200_HOSTSIM
    if (!__any_sync(0xffffffffu, todo)) break;
#else
    if (!todo) break;
#endif
  }
  if (!active) return 0u;
  cb200_d2* col = b.z + IZ(slot, 0, s);
  uint32_t fw = 0;
#pragma unroll 2
  for (int j = 0; j < D; ++j) {
    const cd z = mk(sm.zre[j][tid], sm.zim[j][tid]);
    bool real_image;
    double detj;
    image_eval<NL>(L, z, w, real_image, detj);
    col[j] = make_cb200_d2(z.re, z.im);
    fw |= flag_bits(real_image, detj) << (3 * j);
  }
  b.fw[IS(slot, s)] = fw;
  return fw;
}

__device__ __forceinline__ void store_single(const ExtCfg& cfg, const ExtBuf& b, cd w, int slot, int64_t s) {
  cd z[2];
  single_lens_images(w, z);
  uint32_t fw = 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double det = 1.0 - 1.0 / (norm2(z[j]) * norm2(z[j]));  // 1 - 1/|zbar^2|^2, point_source.py:1562
    b.z[IZ(slot, j, s)] = make_cb200_d2(z[j].re, z[j].im);
    fw |= (1u | (det > 0 ? 2u : 0u) | (det == 0 ? 4u : 0u)) << (3 * j);
  }
  b.fw[IS(slot, s)] = fw;
}

// ---------------------------------------------------------------------------------------------
template <int NL, int NT>
__device__ void limb_walk_body(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L, EASmem<NL * NL + 1, false, NT>& sm, int tid, int64_t s) {
  const bool active = s < nsrc(cfg, b);
  const cd w0 = active ? source_centre(cfg, b, L, s) : mk(0.3, 0.2);
  for (int k = 0; k < cfg.N0; ++k) {
    const double th = theta_init(k, cfg.N0);
    const cd w = limb_point(w0, cfg.rho, th);
    // roots_compensated is not forwarded to the sequential walk (extended_source.py:104-106)
    solve_and_store<NL, false, NT>(cfg, b, L, sm, tid, active, w, k > 0, k, s);
    if (active) b.theta[IS(k, s)] = th;
  }
}

// Small batches, lane-per-root: a warp walks G = 32 / D sources, lane g*D + r follows image track r of
// source g.  Limb point 0 is the same cold Gauss-Seidel solve as in limb_walk_body (on the group's first
// lane: same initial estimates, same root order); every later point is warm-started in registers and
// solved by ea_solve_group (the same Gauss-Seidel sweep with the D evaluations done in parallel).
#ifndef CB200_HOSTSIM
template <int NL, int NT>
__device__ void limb_walk_group_body(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L,
                                     EASmem<NL * NL + 1, false, NT>& sm, int tid, int64_t warp_id) {
  constexpr int D = NL * NL + 1, G = 32 / D;
  const int lane = tid & 31, grp = lane / D, r = lane - grp * D;
  const int64_t s = warp_id * G + grp;
  const bool valid = grp < G && s < nsrc(cfg, b);
  const cd w0 = valid ? source_centre(cfg, b, L, s) : mk(0.3, 0.2);
  {
    cd p[D + 1];
    lens_poly<NL>(L, limb_point(w0, cfg.rho, theta_init(0, cfg.N0)), p);
    ea_normalise<D>(p);
    ea_solve_thread<D, false, NT, CB200_EXT_STRAIGHT != 0>(p, sm, tid, valid && r == 0, false, EA_INIT_REFERENCE, cfg.itmax);
  }
  __syncwarp();
  cd z = mk(0.05 + 0.1 * lane, 0.07 * lane);   // idle lanes: distinct dummies
  if (valid) z = mk(sm.zre[r][tid - r], sm.zim[r][tid - r]);
  for (int k = 0; k < cfg.N0; ++k) {
    const double th = theta_init(k, cfg.N0);
    const cd w = limb_point(w0, cfg.rho, th);
    if (k > 0) {
      cd p[D + 1];
      lens_poly<NL>(L, w, p);
      ea_normalise<D>(p);
      ea_solve_group<D>(p, z, lane, valid, cfg.itmax);
    }
    uint32_t mine = 0;
    if (valid) {
      bool real_image;
      double detj;
      image_eval<NL>(L, z, w, real_image, detj);
      b.z[IZ(k, r, s)] = make_cb200_d2(z.re, z.im);
      mine = flag_bits(real_image, detj) << (3 * r);
    }
    // the group's flag word: every lane collects the D contributions of its group
    uint32_t fw = 0;
#pragma unroll
    for (int q = 0; q < D; ++q) fw |= __shfl_sync(0xffffffffu, mine, (grp < G ? grp * D : 0) + q);
    if (valid && r == 0) {
      b.fw[IS(k, s)] = fw;
      b.theta[IS(k, s)] = th;
    }
  }
}
#endif

__device__ void limb_walk_single_body(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L, int64_t s) {
  if (s >= nsrc(cfg, b)) return;
  const cd w0 = source_centre(cfg, b, L, s);
  for (int k = 0; k < cfg.N0; ++k) {
    const double th = theta_init(k, cfg.N0);
    store_single(cfg, b, limb_point(w0, cfg.rho, th), k, s);
    b.theta[IS(k, s)] = th;
  }
}

// squared width of the interval between two stored limb points: the largest image displacement over
// the tracks where at least one end is a real image (extended_source.py:118-125)
template <int D>
__device__ __forceinline__ double interval_width2(const ExtCfg& cfg, const ExtBuf& b, int sa, int sb, int64_t s) {
  const cb200_d2* ca = b.z + IZ(sa, 0, s);
  const cb200_d2* cb_ = b.z + IZ(sb, 0, s);
  const uint32_t f = b.fw[IS(sa, s)] | b.fw[IS(sb, s)];
  double dmax = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const cb200_d2 za = ca[j], zb = cb_[j];
    const double dx = zb.x - za.x, dy = zb.y - za.y;
    const double d2 = ((f >> (3 * j)) & 1u) ? dx * dx + dy * dy : 0.0;
    dmax = fmax(dmax, d2);
  }
  return dmax;
}

// the same for the limb point just solved (images in the lane's shared root planes, flag word fw) against the
// stored limb points in slots sa and sb: both widths in one pass, the 2 D image loads issued together
template <int D, bool COMP, int NT>
__device__ __forceinline__ void interval_widths2_planes(const ExtCfg& cfg, const ExtBuf& b, const EASmem<D, COMP, NT>& sm,
                                                        int tid, uint32_t fw, int sa, int sb, int64_t s, double& wa, double& wb) {
  const cb200_d2* ca = b.z + IZ(sa, 0, s);
  const cb200_d2* cb_ = b.z + IZ(sb, 0, s);
  const uint32_t fa = fw | b.fw[IS(sa, s)], fb = fw | b.fw[IS(sb, s)];
  double ma = 0.0, mb = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const cb200_d2 za = ca[j], zb = cb_[j];
    const double x = sm.zre[j][tid], y = sm.zim[j][tid];
    const double ax = x - za.x, ay = y - za.y, bx = zb.x - x, by = zb.y - y;
    ma = fmax(ma, ((fa >> (3 * j)) & 1u) ? ax * ax + ay * ay : 0.0);
    mb = fmax(mb, ((fb >> (3 * j)) & 1u) ? bx * bx + by * by : 0.0);
  }
  wa = ma; wb = mb;
}

// ---------------------------------------------------------------------------------------------
// Refinement (extended_source.py:109-133), NITER rounds: rank the intervals of the current theta order by
// the largest image displacement across them, take the nadd widest -- ties to the higher index, like
// argsort(...)[::-1] -- put a new limb point at each midpoint (warm-started from the interval's left end)
// and insert the new arrival slots into the theta order.
//
// A source's refinement state is small: the theta order (2 NP bytes of links) and one width per interval
// (8 NP bytes, `dval`, indexed by the slot that STARTS the interval; a new point changes exactly two of them).
#ifdef CB200_HOSTSIM
constexpr int EXT_WARP = 1;      // the host logic tests run a "warp" of one lane
#else
constexpr int EXT_WARP = 32;
#endif

#ifdef CB200_HOSTSIM
__device__ __forceinline__ unsigned cb200_warp_max(unsigned v) { return v; }
#else
__device__ __forceinline__ unsigned cb200_warp_max(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
#endif

// Selection.  The state is indexed by ARRIVAL SLOT and is independent of the theta order: dval[slot] is the
// squared width of the interval that starts at `slot`, next[slot] the slot that follows it in theta order (a
// singly linked list; the last point, slot N0 - 1 at theta = pi - 1e-8, starts no interval).  The reference
// ranks the intervals by width with argsort(...)[::-1], i.e. descending, equal widths by DESCENDING position
// -- and position order is theta order -- so the scan runs over the slots as they lie in memory and breaks
// ties by theta.  A new point is linked in between its interval's ends: no splice, nothing is shifted.
//
// All lanes of a warp select for ONE source (round_select_body).  Each lane holds its share of the widths (slots lane, lane + 32,
// ...) in registers for the round; pass r = every lane's best, a warp arg-max, and the winner drops the entry it contributed.  Widths are
// non-negative doubles, whose order is the order of their bit patterns: the arg-max is two 32-bit
// redux.sync (high words, then low words among the lanes that tie on the high word) and a ballot.  Equal
// widths (rare) are ranked by theta.  Beyond 32 * SEL_REG limb points: n passes over shared memory, a butterfly
// arg-max per pass (this is also what a one-lane host "warp" runs).
constexpr int SEL_REG = 8;         // widths per lane held in registers: cur <= 32 * SEL_REG, else the generic passes
#ifndef CB200_HOSTSIM
// The same ranking with each lane's widths SORTED once per round (a compare-exchange network on the order-preserving
// 64-bit keys: bit pattern + 1, 0 = no entry): a pass then only compares the lanes' heads -- two redux.sync and a
// ballot -- and the winner shifts its registers, instead of every lane rescanning its R widths for every pick
// (~45 warp instructions per pick against ~120).  Equal widths inside one lane (the theta tie-break of the reference
// would have to order them) are detected by the network; the caller then takes the scanning form, which handles them.
// Equal widths in different lanes are ranked by theta as before.  Returns false (nothing written) on an in-lane tie.
template <int R>
__device__ __forceinline__ bool refine_select_sorted(int cur, int n, int N0, const double* dval, const double* theta,
                                                     uint16_t* left, int lane) {
  unsigned kh[R], kl[R];
  int ix[R];
  bool tie = false;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = lane + k * 32;
    const bool ok = i < cur && i != N0 - 1;
    const double d = ok ? dval[i] : 0.0;
    unsigned long long key = ok ? (unsigned long long)__double_as_longlong(d) + 1ull : 0ull;
    kh[k] = (unsigned)(key >> 32); kl[k] = (unsigned)key; ix[k] = ok ? i : 0;
  }
  auto cex = [&](int a, int c) {      // descending
    const bool lt = kh[a] < kh[c] || (kh[a] == kh[c] && kl[a] < kl[c]);
    const unsigned th = kh[a], tl = kl[a]; const int ti = ix[a];
    kh[a] = lt ? kh[c] : th; kl[a] = lt ? kl[c] : tl; ix[a] = lt ? ix[c] : ti;
    kh[c] = lt ? th : kh[c]; kl[c] = lt ? tl : kl[c]; ix[c] = lt ? ti : ix[c];
  };
  if (R == 6) {
    cex(0, 5); cex(1, 3); cex(2, 4); cex(1, 2); cex(3, 4); cex(0, 3); cex(2, 5); cex(0, 1); cex(2, 3); cex(4, 5); cex(1, 2); cex(3, 4);
  } else {
    static_assert(R == 6 || R == 8, "networks for 6 and 8 keys");
    cex(0, 1); cex(2, 3); cex(4, 5); cex(6, 7); cex(0, 2); cex(1, 3); cex(4, 6); cex(5, 7); cex(1, 2); cex(5, 6);
    cex(0, 4); cex(3, 7); cex(1, 5); cex(2, 6); cex(1, 4); cex(3, 6); cex(2, 4); cex(3, 5); cex(3, 4);
  }
  // equal keys end up next to each other
#pragma unroll
  for (int k = 0; k + 1 < R; ++k) tie = tie || (kh[k] == kh[k + 1] && kl[k] == kl[k + 1] && (kh[k] | kl[k]) != 0u);
  if (__any_sync(0xffffffffu, tie)) return false;
  for (int r = 0; r < n; ++r) {
    const unsigned hi = kh[0];
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lo = hi == mh ? kl[0] : 0u;
    const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
    bool win = hi == mh && lo == ml;
    unsigned winners = __ballot_sync(0xffffffffu, win);
    if (winners & (winners - 1u)) {          // several lanes hold this width: the one latest in theta
      const double th = win ? theta[ix[0]] : -1e300;
      double mth = th;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mth = fmax(mth, __shfl_xor_sync(0xffffffffu, mth, off));
      win = win && th == mth;
      winners = __ballot_sync(0xffffffffu, win);
    }
    const int wl = __ffs(winners) - 1;
    const int pick = __shfl_sync(0xffffffffu, ix[0], wl);
    if (lane == 0) left[r] = (uint16_t)pick;
    if (lane == wl) {
#pragma unroll
      for (int k = 0; k + 1 < R; ++k) { kh[k] = kh[k + 1]; kl[k] = kl[k + 1]; ix[k] = ix[k + 1]; }
      kh[R - 1] = 0u; kl[R - 1] = 0u;
    }
  }
  return true;
}
#endif
__device__ __forceinline__ void refine_select_warp(int cur, int n, int N0, uint16_t* next, const double* dval, double* theta,
                                                   uint16_t* left, uint16_t* right, int lane) {
#ifndef CB200_HOSTSIM
  bool done = false;
  if (cur <= 32 * 6) done = refine_select_sorted<6>(cur, n, N0, dval, theta, left, lane);
  else if (cur <= 32 * 8) done = refine_select_sorted<8>(cur, n, N0, dval, theta, left, lane);
  if (done) {
  } else
#endif
  if (cur <= EXT_WARP * SEL_REG) {
    double dv[SEL_REG];
#pragma unroll
    for (int k = 0; k < SEL_REG; ++k) {
      const int i = lane + k * EXT_WARP;
      dv[k] = (i < cur && i != N0 - 1) ? dval[i] : -1.0;
    }
    for (int r = 0; r < n; ++r) {
      double bd = dv[0];
      int bk = 0;
#pragma unroll
      for (int k = 1; k < SEL_REG; ++k) {
        bool better = dv[k] > bd;
        if (dv[k] == bd && bd >= 0.0) better = theta[lane + k * EXT_WARP] > theta[lane + bk * EXT_WARP];
        if (better) { bd = dv[k]; bk = k; }
      }
      const int mine = lane + bk * EXT_WARP;
      // bit patterns of widths >= 0 order like the widths; 0 = "no entry left"
      const unsigned hi = bd >= 0.0 ? (unsigned)__double2hiint(bd) + 1u : 0u;
      const unsigned mh = cb200_warp_max(hi);
      const unsigned lo = hi == mh ? (unsigned)__double2loint(bd) : 0u;
      const unsigned ml = cb200_warp_max(lo);
      bool win = hi == mh && lo == ml;
      unsigned winners = __ballot_sync(0xffffffffu, win);
      if (winners & (winners - 1u)) {          // several lanes hold this width: the one latest in theta
        double th = win ? theta[mine] : -1e300;
        double mth = th;
#pragma unroll
        for (int off = EXT_WARP / 2; off > 0; off >>= 1) mth = fmax(mth, __shfl_xor_sync(0xffffffffu, mth, off));
        win = win && th == mth;
        winners = __ballot_sync(0xffffffffu, win);
      }
      const int wl = __ffs(winners) - 1;
      const int pick = __shfl_sync(0xffffffffu, mine, wl);
      if (lane == 0) left[r] = (uint16_t)pick;
      if (lane == wl) {
#pragma unroll
        for (int k = 0; k < SEL_REG; ++k) if (k == bk) dv[k] = -1.0;
      }
    }
  } else {
    double pd = 1e300;
    int pi = -1;                       // the previous pass's pick (width, slot)
    for (int r = 0; r < n; ++r) {
      double bd = -1.0;
      int bi = 0;
      for (int i = lane; i < cur; i += EXT_WARP) {
        const double d = dval[i];
        bool take = d < pd && d > bd && i != N0 - 1;
        if (((d == pd && i != pi) || d == bd) && i != N0 - 1) {
          const double th = theta[i];
          take = (d < pd || (d == pd && th < theta[pi])) && (d > bd || (d == bd && th > theta[bi]));
        }
        if (take) { bd = d; bi = i; }
      }
#pragma unroll
      for (int off = EXT_WARP / 2; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        // lanes without a candidate hold -1; equal widths (rare) by theta, so that both partners keep the same one
        if (od > bd || (od == bd && od >= 0.0 && oi != bi && theta[oi] > theta[bi])) { bd = od; bi = oi; }
      }
      if (lane == 0) left[r] = (uint16_t)bi;
      pd = bd; pi = bi;
    }
  }
  __syncwarp();
  // new points: rank r -> arrival slot cur + r; theta at the interval midpoint, warm start from the left end
  // (the selected intervals are distinct, so the lanes link their points in independently)
  for (int r = lane; r < n; r += EXT_WARP) {
    const int sl = left[r], sr = next[sl];
    theta[cur + r] = 0.5 * (theta[sl] + theta[sr]);
    right[r] = (uint16_t)sr;
    next[cur + r] = (uint16_t)sr;
    next[sl] = (uint16_t)(cur + r);
  }
  __syncwarp();
}
// the theta order as an array (arrival slot of the p-th point), from the linked list
__device__ __forceinline__ void order_from_links(int NP, const uint16_t* next, uint16_t* order) {
  int slot = 0;   // theta = -pi is the first limb point of the walk
  for (int p = 0; p < NP; ++p) { order[p] = (uint16_t)slot; slot = next[slot]; }
}

// warm start of new point r of a round: the left neighbour's images plus the reference's jitters
// (extended_source.py:83-85), into the lane's shared root planes
template <int D, bool COMP, int NT>
__device__ __forceinline__ void warm_start_from(const ExtCfg& cfg, const ExtBuf& b, EASmem<D, COMP, NT>& sm, int tid,
                                                int lf, int r, int64_t s) {
  const cb200_d2* col = b.z + IZ(lf, 0, s);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const double* jt = b.jit + 2 * (j * cfg.nadd + r);
    const cb200_d2 v = col[j];
    sm.zre[j][tid] = v.x + jt[0];
    sm.zim[j][tid] = v.y + jt[1];
  }
}

// One selection kernel and one solve kernel PER ROUND, state in global memory: a warp per source selects
// (coalesced lane-strided reads of the source's widths and links), then a THREAD per (source, new point) solves --
// consecutive threads are one source's nadd points, so a warp's accesses stay inside three or four records.
// (Round 2 also built the rounds as a loop inside one kernel, a warp owning 32 / nadd sources with widths and
// links in shared memory: 8.7 ms for the C4 batch against 7.0 ms for these twenty launches, and slower or equal
// at every batch size down to 1000 sources -- the selection is a latency-bound phase of ten dependent passes that a
// warp-owned kernel cannot hide behind its own solves; profiles/r02_ext_variants.txt.)
template <int D>
__device__ void round_select_body(const ExtCfg& cfg, const ExtBuf& b, int round, int lane, int64_t s) {
  if (s >= nsrc(cfg, b)) return;   // warp-uniform
  const int NP = cfg.NP, n = cfg.nadd;
  double* dv = b.rdval + IS(0, s);
  uint16_t* nxt = b.rnext + IS(0, s);
  uint16_t* lft = b.rlr + s * 2 * NADD_MAX;
  if (round == 0) {
    for (int i = lane; i < NP; i += EXT_WARP) {
      nxt[i] = (uint16_t)(i + 1 < cfg.N0 ? i + 1 : 0);
      dv[i] = i + 1 < cfg.N0 ? interval_width2<D>(cfg, b, i, i + 1, s) : 0.0;
    }
    __syncwarp();
  }
  refine_select_warp(cfg.N0 + round * n, n, cfg.N0, nxt, dv, b.theta + IS(0, s), lft, lft + NADD_MAX, lane);
  // the links are final after the last selection (solves add no points): the theta order as an array
  if (round == NITER - 1 && lane == 0) order_from_links(NP, nxt, b.order + IS(0, s));
}
template <int NL, bool COMP, int NT>
__device__ void round_solve_body(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L,
                                 EASmem<(NL == 1 ? 2 : NL * NL + 1), COMP, NT>& sm, int tid, int64_t g, int round) {
  constexpr int D = NL == 1 ? 2 : NL * NL + 1;
  const int n = cfg.nadd;
  const int64_t s = g / n;
  const int r = (int)(g - s * n), slot = cfg.N0 + round * n + r;
  const bool active = s < nsrc(cfg, b);
  int lf = 0, rt = 0;
  cd w = mk(0.3, 0.2);
  if (active) {
    const uint16_t* lft = b.rlr + s * 2 * NADD_MAX;
    lf = lft[r]; rt = lft[NADD_MAX + r];
    w = limb_point(source_centre(cfg, b, L, s), cfg.rho, b.theta[IS(slot, s)]);
  }
  double wl = 0.0, wr = 0.0;
  if constexpr (NL == 1) {
    if (active) {
      store_single(cfg, b, w, slot, s);
      wl = interval_width2<D>(cfg, b, lf, slot, s);
      wr = interval_width2<D>(cfg, b, slot, rt, s);
    }
  } else {
    if (active) warm_start_from<D, COMP, NT>(cfg, b, sm, tid, lf, r, s);
    const uint32_t fw = solve_and_store<NL, COMP, NT>(cfg, b, L, sm, tid, active, w, true, slot, s);
    if (active) interval_widths2_planes<D, COMP, NT>(cfg, b, sm, tid, fw, lf, rt, s, wl, wr);
  }
  if (active) {
    b.rdval[IS(lf, s)] = wl;
    b.rdval[IS(slot, s)] = wr;
  }
}

// The (deg, nadd) jitter table is the same for every source and every round: one tiny launch fills it
// (entry t = j * nadd + r) instead of ~20 threefry blocks per refinement solve.
__device__ __forceinline__ void jitter_table_body(int D, int nadd, double* out, int t) {
  if (t >= D * nadd) return;
  const cd v = limb_jitter(t / nadd, t % nadd, D, nadd);
  out[2 * t] = v.re;
  out[2 * t + 1] = v.im;
}
// duplicates are rare: keep the generator out of line so it does not cost the matching loop registers
__device__ __noinline__ double duplicate_jitter_cold(int j, int p, int deg, int npts) {
  return duplicate_jitter(j, p, deg, npts);
}

// ---------------------------------------------------------------------------------------------
// theta order + duplicate guard + greedy nearest-neighbour track matching (utils.py:15-40): for each
// track i in order, the nearest not-yet-taken root of the next limb point (ties: lowest index).
// TrackWalk visits the limb points of one source in theta order; step(p) hands out limb point p's images
// in TRACK order together with the permutation (4 bits per track: which image of the column it took).
template <int D>
struct TrackWalk {
  const ExtCfg& cfg; const ExtBuf& b; const int64_t s;
  double cre[D], cim[D];                  // previous column in track order
  double nzr[D], nzi[D]; uint32_t nfw;    // raw column of the NEXT limb point, fetched while the current one is matched
  int nslot, slot;
  cb200_d2* stg; int sstride;             // this thread's column in shared memory, [image * sstride] (dynamic index `best`)
  __device__ __forceinline__ void fetch(int p) {
    nslot = b.order[IS(p, s)];
    const cb200_d2* col = b.z + IZ(nslot, 0, s);
    nfw = b.fw[IS(nslot, s)];
#pragma unroll
    for (int j = 0; j < D; ++j) { const cb200_d2 v = col[j]; nzr[j] = v.x; nzi[j] = v.y; }
  }
  __device__ __forceinline__ TrackWalk(const ExtCfg& c, const ExtBuf& bb, int64_t ss, cb200_d2* stage, int stride)
      : cfg(c), b(bb), s(ss), stg(stage), sstride(stride) { fetch(0); }
  // WRITE_BACK: a duplicate's offset is also stored into the source's record, so that later random access
  // through the permutation (Tracks, perm mode) sees the value the matching saw
  template <bool WRITE_BACK>
  __device__ __forceinline__ uint64_t step(int p, double (&vr)[D], double (&vi)[D], uint8_t (&vf)[D]) {
    double zr[D], zi[D];
    const uint32_t fw = nfw;
    slot = nslot;
#pragma unroll
    for (int j = 0; j < D; ++j) { zr[j] = nzr[j]; zi[j] = nzi[j]; }
    if (p + 1 < cfg.NP) fetch(p + 1);
    // exact duplicates (inside the column, or an unchanged warm start) get a tiny real offset.  They are rare:
    // the full test (2 D^2 compares of both parts) runs only when some REAL parts coincide, on a warp vote
    bool maybe = false;
#pragma unroll
    for (int j = 0; j < D; ++j) {
#pragma unroll
      for (int k = 0; k < D; ++k) {
        if (k < j) maybe = maybe || zr[k] == zr[j];
        if (p > 0) maybe = maybe || cre[k] == zr[j];
      }
    }
#ifndef CB200_HOSTSIM
    maybe = __any_sync(__activemask(), maybe);
#endif
    if (maybe) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      bool dup = false;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        if (k < j && zr[k] == zr[j] && zi[k] == zi[j]) dup = true;
        if (p > 0 && cre[k] == zr[j] && cim[k] == zi[j]) dup = true;
      }
      if (dup) {
        zr[j] += duplicate_jitter_cold(j, p, D, cfg.NP);   // extended_source.py:144-148
        if (WRITE_BACK) b.z[IZ(slot, j, s)].x = zr[j];
      }
    }
    }
    if (D > 5) {
#pragma unroll
      for (int j = 0; j < D; ++j) stg[j * sstride] = make_cb200_d2(zr[j], zi[j]);
    }
    unsigned used = 0;
    uint64_t perm = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      int best = i;
      if (p > 0) {
        double bd = 1e300;
        best = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double dx = zr[k] - cre[i], dy = zi[k] - cim[i];
          const double d2 = dx * dx + dy * dy;
          if (!((used >> k) & 1u) && d2 < bd) { bd = d2; best = k; }
        }
        if (bd == 1e300) {   // NaNs: take the first free slot
#pragma unroll
          for (int k = D - 1; k >= 0; --k) if (!((used >> k) & 1u)) best = k;
        }
      }
      used |= 1u << best;
      perm |= (uint64_t)best << (4 * i);
      // the chosen image by a shared-memory read at a dynamic index (a D-way register select costs 4 D + D
      // instructions per track: a quarter of this kernel)
      if (D > 5) {
        const cb200_d2 v = stg[best * sstride];
        vr[i] = v.x; vi[i] = v.y;
      } else {   // a handful of images: the register select is as cheap (measured: binary lens 6.63 vs 6.71 ms)
        double r = 0, m = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) if (k == best) { r = zr[k]; m = zi[k]; }
        vr[i] = r; vi[i] = m;
      }
      vf[i] = (uint8_t)((fw >> (3 * best)) & 7u);
    }
    return perm;
  }
  __device__ __forceinline__ void advance(const double (&vr)[D], const double (&vi)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) { cre[i] = vr[i]; cim[i] = vi[i]; }
  }
};

// the theta-ordered tracks as arrays (consumed by contours_body; limb-darkened, tangent and export calls)
template <int D>
__device__ void tracks_body(const ExtCfg& cfg, const ExtBuf& b, int64_t s, cb200_d2* stage, int sstride) {
  if (s >= nsrc(cfg, b)) return;
  TrackWalk<D> W(cfg, b, s, stage, sstride);
  for (int p = 0; p < cfg.NP; ++p) {
    double vr[D], vi[D];
    uint8_t vf[D];
    W.template step<false>(p, vr, vi, vf);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      b.sre[IT(p, i, s)] = vr[i];
      b.sim[IT(p, i, s)] = vi[i];
      b.sflg[IT(p, i, s)] = vf[i];
    }
    W.advance(vr, vi);
  }
}

// Uniform disk, large batches: ONE pass over the source's record does the track matching AND integrates
// every closed track while its vertices go by (the trapezoid sum of integrate.py:23-27 only ever needs
// the previous vertex, which is the matching's carry).  Nothing but the 8-byte permutation per limb point
// is written.  A source whose limb crosses a caustic -- some track holds real images but is not closed --
// is appended to `open_list`; open_body splits, stitches and integrates its open tracks by random access
// through the permutation and adds them to the closed-track sum saved here.  Sums are formed in the order
// contours_body forms them (per track in theta order, closed tracks in track order, then the stitched
// contours), so the two paths agree bit for bit.
template <int D>
__device__ void sweep_body(const ExtCfg& cfg, const ExtBuf& b, int64_t s, cb200_d2* stage, int sstride) {
  if (s >= nsrc(cfg, b)) return;
  constexpr unsigned FULL = (1u << D) - 1u;
  const int NP = cfg.NP;
  TrackWalk<D> W(cfg, b, s, stage, sstride);
  double sum[D];
  unsigned all_real = FULL, any_real = 0, f0 = 0;
  uint64_t* perm = b.perm + IS(0, s);
  {
    double vr[D], vi[D];
    uint8_t vf[D];
    perm[0] = W.template step<true>(0, vr, vi, vf);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      sum[i] = 0.0;
      f0 |= (unsigned)vf[i] << (3 * i);
      if (vf[i] & 1) any_real |= 1u << i; else all_real &= ~(1u << i);
    }
    W.advance(vr, vi);
  }
  for (int p = 1; p < NP; ++p) {
    double vr[D], vi[D];
    uint8_t vf[D];
    perm[p] = W.template step<true>(p, vr, vi, vf);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      sum[i] += 0.5 * (W.cre[i] * vi[i] - vr[i] * W.cim[i]);
      if (vf[i] & 1) any_real |= 1u << i; else all_real &= ~(1u << i);
    }
    W.advance(vr, vi);
  }
  unsigned closed = 0;
  double total = 0.0;
  // the first limb point again (track i = image i of that column; a duplicate's offset was written back)
  const cb200_d2* col0 = b.z + IZ(b.order[IS(0, s)], 0, s);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const cb200_d2 first = col0[i];
    const double fre_i = first.x, fim_i = first.y;
    const double dx = fre_i - W.cre[i], dy = fim_i - W.cim[i];
    const bool cl = cfg.nl == 1 || (((all_real >> i) & 1u) && dx * dx + dy * dy < 1e-10);
    if (cl) {
      closed |= 1u << i;
      const unsigned f = (f0 >> (3 * i)) & 7u;
      const double par = (f & 4) ? 0.0 : ((f & 2) ? 1.0 : -1.0);
      const double S = sum[i] + 0.5 * (W.cre[i] * fim_i - fre_i * W.cim[i]);
      total += par * S;
    }
  }
  if (any_real & ~closed) {
    b.sw_total[s] = total;
    b.sw_closed[s] = closed | ((any_real & ~closed) << 16);   // high half: the tracks the open pass has to look at
    b.open_list[cb200_atomic_inc(b.open_count)] = (int32_t)s;
  } else {
    b.mag[src_index(b, s)] = fabs(total) * (1.0 / (3.14159265358979323846 * cfg.rho * cfg.rho));
  }
}

// ---------------------------------------------------------------------------------------------
// Contours.  A segment is an index range [lo, hi) of one track; a stitched contour is a chain of
// (segment, reversed) pieces.  Points are read from the theta-ordered arrays on demand.
struct Seg { int16_t track, lo, hi; int8_t par; double len; };

// Optional shared-memory copy of one source's tracks ([p][track] planes), used for small batches:
// the stitching logic is a chain of dependent reads, ~20x faster from shared memory than from L2.
// Element (track, p) lives at [p * sp + track * st]: sp = D, st = 1 for the [p][track] planes; the compact open pass
// stages only the tracks it needs, track-major (sp = 1, st = NP), and numbers them 0..W-1.
struct TrackStage { const double* re; const double* im; const uint8_t* f; const double* th; int sp, st; };

struct Tracks {
  const ExtCfg& cfg; const ExtBuf& b; int64_t s; const TrackStage* st;
  const uint64_t* perm;    // non-null: no track arrays, read the source's record through order and permutation
  __device__ __forceinline__ cd pt(int track, int p) const {
    if (st) return mk(st->re[p * st->sp + track * st->st], st->im[p * st->sp + track * st->st]);
    if (perm) {
      const cb200_d2 v = b.z[IZ(b.order[IS(p, s)], (int)((perm[p] >> (4 * track)) & 15u), s)];
      return mk(v.x, v.y);
    }
    return mk(b.sre[(((int64_t)p * cfg.D + track) * cfg.S + s)], b.sim[(((int64_t)p * cfg.D + track) * cfg.S + s)]);
  }
  __device__ __forceinline__ double th(int p) const {
    if (!b.vth) return 0.0;
    if (st) return st->th[p];
    return b.theta[(int64_t)s * cfg.NP + b.order[(int64_t)s * cfg.NP + p]];
  }
  __device__ __forceinline__ uint8_t fl(int track, int p) const {
    if (st) return st->f[p * st->sp + track * st->st];
    if (perm) return (uint8_t)((b.fw[IS(b.order[IS(p, s)], s)] >> (3 * (int)((perm[p] >> (4 * track)) & 15u))) & 7u);
    return b.sflg[(((int64_t)p * cfg.D + track) * cfg.S + s)];
  }
};

struct Chain {
  // pieces live in a deque so that H-T / H-H connections can prepend (extended_source.py:445-492).  The storage
  // (2 x 64 bytes) is the caller's: a local array for the thread-per-source kernels, shared memory where one lane of a
  // warp stitches (a lane's local array is spread over 128-byte lines it alone uses, which thrash L1)
  int8_t* seg; int8_t* rev; int head, tail;   // pieces [head, tail)
  int npts;
  __device__ void init(int8_t* mem, int sg, int n) { seg = mem; rev = mem + 64; head = 32; tail = 33; seg[32] = (int8_t)sg; rev[32] = 0; npts = n; }
};
// a candidate segment seen as a chain of one piece (never materialised)
struct OneSeg { int sg; int npts; };

__device__ __forceinline__ int seg_n(const Seg& g) { return g.hi - g.lo; }
__device__ __forceinline__ cd seg_pt(const Tracks& T, const Seg& g, int k, bool reversed) {
  return T.pt(g.track, reversed ? g.hi - 1 - k : g.lo + k);
}
// k-th point of the chain counted from its head (from_tail: from its tail); 0 beyond the ends, like
// the reference's zero padding
__device__ cd chain_pt(const Tracks& T, const Seg* segs, const Chain& c, int k, bool from_tail) {
  if (k < 0 || k >= c.npts) return mk(0, 0);
  if (!from_tail) {
    for (int q = c.head; q < c.tail; ++q) {
      const Seg& g = segs[c.seg[q]];
      const int n = seg_n(g);
      if (k < n) return seg_pt(T, g, k, c.rev[q]);
      k -= n;
    }
  } else {
    for (int q = c.tail - 1; q >= c.head; --q) {
      const Seg& g = segs[c.seg[q]];
      const int n = seg_n(g);
      if (k < n) return seg_pt(T, g, n - 1 - k, c.rev[q]);
      k -= n;
    }
  }
  return mk(0, 0);
}
__device__ __forceinline__ cd chain_pt(const Tracks& T, const Seg* segs, const OneSeg& c, int k, bool from_tail) {
  if (k < 0 || k >= c.npts) return mk(0, 0);
  return seg_pt(T, segs[c.sg], from_tail ? c.npts - 1 - k : k, false);
}
__device__ __forceinline__ double chain_parity(const Seg* segs, const Chain& c) {
  const double p = segs[c.seg[c.head]].par;
  return c.rev[c.head] ? -p : p;
}
__device__ __forceinline__ double chain_parity(const Seg* segs, const OneSeg& c) { return segs[c.sg].par; }
// the two points that define the direction at an end, connection point last; a near-duplicate end
// vertex (< 1e-5) is skipped (extended_source.py:368-390)
template <class C>
__device__ void end_line(const Tracks& T, const Seg* segs, const C& c, bool tail, cd& a, cd& bb) {
  const int t = c.npts - 1;
  const cd p0 = chain_pt(T, segs, c, 0, tail), p1 = chain_pt(T, segs, c, 1, tail);
  if (norm2(p1 - p0) > 1e-10 || t <= 1) { a = p1; bb = p0; }
  else { a = chain_pt(T, segs, c, 2, tail); bb = p1; }
}
__device__ bool connect_ok(const Tracks& T, const Seg* segs, const Chain& c1, const OneSeg& c2, int ctype) {
  const bool same = chain_parity(segs, c1) * chain_parity(segs, c2) > 0.0;
  if ((ctype < 2) != same) return false;
  cd a1, b1, a2, b2;
  end_line(T, segs, c1, ctype == 0 || ctype == 3, a1, b1);
  end_line(T, segs, c2, ctype == 1 || ctype == 3, a2, b2);
  const double dist2 = norm2(b1 - b2);
  if (dist2 < 1e-10) return true;                       // dist < min_dist
  if (!(dist2 < 1e-2)) return false;                    // dist < max_dist = 0.1
  const cd v1 = b1 - a1, v2 = b2 - a2;
  const double cosang = (v1.re * v2.re + v1.im * v2.im) * rsqrt(norm2(v1) * norm2(v2));
  // 180 - deg(acos(c)) < 60  <=>  acos(c) > 120 deg  <=>  c < -1/2
  if (!(cosang < -0.5)) return false;
  return dist2 < norm2(a1 - a2);                        // ends closer than the points before them
}

struct GreenAcc {
  // uniform: 1/2 sum (x_i y_{i+1} - x_{i+1} y_i), the trapezoid rule of integrate.py:23-27
  double sum; cd first, prev; bool any;
  __device__ void start() { sum = 0.0; any = false; }
  __device__ void add(cd z) {
    if (any) sum += 0.5 * (prev.re * z.im - z.re * prev.im);
    else { first = z; any = true; }
    prev = z;
  }
  __device__ double close() { if (any) add(first); return sum; }
};

// Tangent of the uniform-disk area with respect to the low-level parameters, accumulated while the
// contour is walked (the rule is the reference's implicit-function JVP,
// ehrlich_aberth_primitive.py:290-324, written on the lens equation: every contour vertex is a zero of
// F(z) = z - sum_j eps_j / (conj z - conj r_j) - w,  w = wc + rho e^{i theta}, so
//   dz/dt = (-F_t + g conj(F_t)) / (1 - |g|^2),   g = dF/d(conj z) = sum_j eps_j / (conj z - conj r_j)^2,
// with the sampling theta, the masks and the contour topology constants -- exactly what jax.grad sees).
// Parameters t, in the order of the C ABI: a, e1, e2, Re r3, Im r3, Re wc, Im wc, rho.
// The area 1/2 sum_i (x_i y_{i+1} - x_{i+1} y_i) of a closed polygon has the tangent
//   1/2 sum_i [ dx_i (y_{i+1} - y_{i-1}) - dy_i (x_{i+1} - x_{i-1}) ]   (cyclic),
// so a vertex's tangent meets only the POSITIONS of its neighbours: no tangent is carried along the walk.
constexpr int NGRAD = 8;
template <int NL>
struct GreenTangent {
  const LensConst& L;
  double inv_rho; cd wc;    // 1 / source radius, source centre (lens frame)
  double d[NGRAD];          // tangent of the current contour's area
  cd z0, z1, zp, zc;        // first two vertices; previous and current vertex
  int cnt;
  __device__ GreenTangent(const LensConst& L_, double rho_, cd wc_) : L(L_), inv_rho(1.0 / rho_), wc(wc_) {}
  __device__ void start() {
#pragma unroll
    for (int k = 0; k < NGRAD; ++k) d[k] = 0.0;
    cnt = 0;
  }
  // vertex z with neighbours zprev, znext.  The limb direction e^{i theta} that d w / d rho needs is read off the
  // vertex itself: z solves lens_eq(z) = wc + rho e^{i theta}, and lens_eq(z) = z - sum_j eps_j u_j comes for two
  // FMAs per lens from the u_j the step forms anyway (instead of two dependent, uncoalesced loads of the limb
  // angle through the theta order and a sincos per vertex; it differs from the sampled direction by the root's
  // residual / rho, ~1e-13)
  __device__ void vertex(cd z, cd zprev, cd znext) {
    const cd zb = conj(z);
    cd u[3], g = mk(0, 0), wz = z;
    constexpr int M = NL == 1 ? 1 : NL;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      u[j] = crecip(NL == 1 ? zb : zb - conj(L.r[j]));
      g = g + (NL == 1 ? 1.0 : L.eps[j]) * (u[j] * u[j]);
      wz = wz - (NL == 1 ? 1.0 : L.eps[j]) * u[j];
    }
    const double den = 1.0 / (1.0 - norm2(g));
    const double wy = 0.5 * (znext.im - zprev.im), wx = 0.5 * (znext.re - zprev.re);
    auto acc = [&](int k, cd Ft) {
      const cd dz = den * (g * conj(Ft) - Ft);
      d[k] += dz.re * wy - dz.im * wx;
    };
    const double cs = inv_rho * (wz.re - wc.re), sn = inv_rho * (wz.im - wc.im);
    acc(5, mk(-1.0, 0.0));
    acc(6, mk(0.0, -1.0));
    acc(7, mk(-cs, -sn));
    if (NL >= 2) {
      const cd q0 = L.eps[0] * (u[0] * u[0]), q1 = L.eps[1] * (u[1] * u[1]);
      acc(0, q1 - q0);                                        // r_1 = a, r_2 = -a
      if (NL == 2) acc(1, u[1] - u[0]);                       // eps = (e1, 1 - e1)
      else {
        const cd q2 = L.eps[2] * (u[2] * u[2]);
        acc(1, u[2] - u[0]);                                  // eps = (e1, e2, 1 - e1 - e2)
        acc(2, u[2] - u[1]);
        acc(3, -q2);                                          // r_3 = r3: F_t = -eps_3 u_3^2 conj(dr3/dt)
        acc(4, mk(0.0, 1.0) * q2);
      }
    }
  }
  __device__ void add(cd z) {
    if (cnt == 0) z0 = z;
    else if (cnt == 1) z1 = z;
    else vertex(zc, zp, z);
    zp = cnt == 0 ? z : zc;
    zc = z;
    ++cnt;
  }
  // closes the polygon (the reference appends the first point again, :724-725) and adds parity * tangent to out
  __device__ void close(double par, double (&out)[NGRAD]) {
    if (cnt >= 2) {
      vertex(zc, zp, z0);             // last vertex: neighbours (previous, first)
      vertex(z0, zc, z1);             // first vertex: neighbours (last, second)
    }
#pragma unroll
    for (int k = 0; k < NGRAD; ++k) out[k] += par * d[k];
  }
};

// limb-darkening vertex emitter: appends the vertices of one closed contour (closing point included)
struct LdEmit {
  const ExtCfg& cfg; const ExtBuf& b; int64_t s; int nv, nc; cd first, csum; int cnt0; bool any; double first_th;
  __device__ void start() { any = false; csum = mk(0, 0); cnt0 = nv; }
  __device__ void put(cd z, double th) {
    if (nv < cfg.VMAX) {
      b.vz[(int64_t)nv * cfg.S + s] = make_cb200_d2(z.re, z.im);
      b.vcid[(int64_t)nv * cfg.S + s] = (uint8_t)nc;
      if (b.vth) b.vth[(int64_t)nv * cfg.S + s] = th;
    }
    ++nv; csum = csum + z;
  }
  __device__ void add(cd z, double th) { if (!any) { first = z; first_th = th; any = true; } put(z, th); }
  __device__ void close(double parity) {
    if (!any) return;
    put(first, first_th);
    if (nc < cfg.CMAX) {
      const double inv = 1.0 / (double)(nv - cnt0);      // centroid incl. the closing point (integrate.py:112)
      if (b.cz0) b.cz0[(int64_t)nc * cfg.S + s] = make_cb200_d2(csum.re * inv, csum.im * inv);
      b.cpar[(int64_t)nc * cfg.S + s] = parity;
      b.cstart[(int64_t)nc * cfg.S + s] = cnt0;
    }
    ++nc;
  }
};

// Split the open tracks of one source into segments: runs of real images of one parity without jumps > 0.1
// (extended_source.py:156-250).  track_parts scans ONE track and returns its (start, end) pairs in
// increasing order (the first MAXPARTS starts and ends are kept, :202-203); part_segment turns a pair into
// a segment, or returns false for the reference's "empty" marker and for runs of fewer than two points (:232).
__device__ int track_parts(const ExtCfg& cfg, const Tracks& T, int i, int16_t* lo, int16_t* hi) {
  const int NP = cfg.NP;
  int np_ = 0, nend = 0;
  bool prev_real = false; double prev_par = 0.0; cd prev_z = mk(0, 0);
#pragma unroll 4
  for (int p = 0; p <= NP; ++p) {
    bool real = false; double par = 0.0; cd z = mk(0, 0);
    {
      // unconditional loads (clamped index) so that unrolled iterations overlap their latency
      const int pc = p < NP ? p : NP - 1;
      const uint8_t f = T.fl(i, pc);
      const cd zz = T.pt(i, pc);
      real = (p < NP) && (f & 1);
      if (real) { par = (f & 4) ? 0.0 : ((f & 2) ? 1.0 : -1.0); z = zz; }
    }
    bool start = false, end = false;
    if (p == 0) start = real;
    else if (p == NP) end = prev_real;
    else {
      const double dm = (real ? 1.0 : 0.0) - (prev_real ? 1.0 : 0.0);
      const bool change = norm2(z - prev_z) > 0.01 || par != prev_par || dm != 0.0;
      start = change && dm >= 0.0;
      end = change && dm <= 0.0;
    }
    if (end && nend < MAXPARTS) hi[nend++] = (int16_t)p;
    if (start && np_ < MAXPARTS) lo[np_++] = (int16_t)p;
    prev_real = real; prev_par = par; prev_z = z;
  }
  return np_ < nend ? np_ : nend;   // k-th start pairs with k-th end
}
__device__ bool part_segment(const Tracks& T, int i, int lo, int hi, Seg& g) {
  if (hi - lo < 2 || (lo == 0 && hi == 0)) return false;
  g.track = (int16_t)i; g.lo = (int16_t)lo; g.hi = (int16_t)hi;
  const uint8_t f = T.fl(i, lo);
  g.par = (f & 4) ? 0 : ((f & 2) ? 1 : -1);
  double len = 0.0; cd q0 = T.pt(i, lo);
  for (int p = lo + 1; p < hi; ++p) { const cd q1 = T.pt(i, p); len += sqrt(norm2(q1 - q0)); q0 = q1; }
  g.len = len;
  return true;
}
// The reference keeps, of all parts in (track, start) order REVERSED, the first 3(nl^2+1) that have at least
// two points: walk the tracks backwards and each track's parts backwards.  Returns the number of segments
// written to parts[MAXSEG].
template <int D>
__device__ int build_parts_serial(const ExtCfg& cfg, const Tracks& T, unsigned closed, Seg* parts) {
  int nparts_total = 0;
  const int nseg_max = 3 * (cfg.nl * cfg.nl + 1);
  for (int i = D - 1; i >= 0 && nparts_total < nseg_max; --i) {
    if ((closed >> i) & 1u) continue;
    int16_t lo[MAXPARTS], hi[MAXPARTS];
    const int npair = track_parts(cfg, T, i, lo, hi);
    for (int k = npair - 1; k >= 0 && nparts_total < nseg_max; --k) {
      Seg g;
      if (part_segment(T, i, lo[k], hi[k], g)) parts[nparts_total++] = g;
    }
  }
  return nparts_total;
}
// One warp per source, the tracks staged in shared memory: lane i splits track i (the scans are independent),
// lane 0 then collects the segments in the reference's order.  tp: [D][MAXPARTS] segments, tn: [D] their counts
// (valid[k] in Seg::track < 0 marks a dropped pair).
template <int D>
__device__ int build_parts_warp(const ExtCfg& cfg, const Tracks& T, unsigned closed, Seg* tp, int* tn, Seg* parts, int lane) {
  if (lane < D) {
    int npair = 0;
    if (!((closed >> lane) & 1u)) {
      int16_t lo[MAXPARTS], hi[MAXPARTS];
      npair = track_parts(cfg, T, lane, lo, hi);
      for (int k = 0; k < npair; ++k) {
        Seg g;
        if (!part_segment(T, lane, lo[k], hi[k], g)) g.track = -1;
        tp[lane * MAXPARTS + k] = g;
      }
    }
    tn[lane] = npair;
  }
#ifndef CB200_HOSTSIM
  __syncwarp();
#endif
  int nparts_total = 0;
  if (lane == 0) {
    const int nseg_max = 3 * (cfg.nl * cfg.nl + 1);
    for (int i = D - 1; i >= 0 && nparts_total < nseg_max; --i)
      for (int k = tn[i] - 1; k >= 0 && nparts_total < nseg_max; --k)
        if (tp[i * MAXPARTS + k].track >= 0) parts[nparts_total++] = tp[i * MAXPARTS + k];
  }
  return nparts_total;
}

// `resume`: the closed tracks were integrated by sweep_body (their mask and sum are in sw_closed / sw_total);
// only the open tracks are left, read through the permutation.
template <int D, bool GRAD = false>
__device__ void contours_body(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L, int64_t s,
                              const TrackStage* stage = nullptr, bool resume = false,
                              const Seg* pre_parts = nullptr, int pre_nparts = 0, int8_t* chain_mem = nullptr) {
  if (s >= nsrc(cfg, b)) return;
  constexpr int NLG = D == 2 ? 1 : (D == 5 ? 2 : 3);
  GreenTangent<NLG> GT(L, cfg.rho, GRAD ? source_centre(cfg, b, L, s) : mk(0, 0));
  double dtot[NGRAD];
  if (GRAD) {
#pragma unroll
    for (int k = 0; k < NGRAD; ++k) dtot[k] = 0.0;
  }
  const Tracks T{cfg, b, s, stage, resume ? b.perm + IS(0, s) : nullptr};
  const int NP = cfg.NP;
  const double norm = 1.0 / (3.14159265358979323846 * cfg.rho * cfg.rho);
  const int64_t out_idx = src_index(b, s);
  double total = 0.0;
  const bool emit = cfg.ld || cfg.emit;
  LdEmit E{cfg, b, s, 0, 0};
  GreenAcc G;

  // closed tracks: every image real and the track returns to its start (extended_source.py:290)
  unsigned closed = 0;
  if (resume) { closed = b.sw_closed[s] & ((1u << D) - 1u); total = b.sw_total[s]; }
  for (int i = 0; i < D && !resume; ++i) {
    bool all_real = true;
#pragma unroll 8
    for (int p = 0; p < NP; ++p) all_real = all_real && (T.fl(i, p) & 1);
    const bool cl = cfg.nl == 1 || (all_real && norm2(T.pt(i, 0) - T.pt(i, NP - 1)) < 1e-10);
    if (!cl) continue;
    closed |= 1u << i;
    const uint8_t f0 = T.fl(i, 0);
    const double par = (f0 & 4) ? 0.0 : ((f0 & 2) ? 1.0 : -1.0);
    if (emit) {
      E.start();
      for (int p = 0; p < NP; ++p) E.add(T.pt(i, p), T.th(p));
      E.close(par);
    }
    if (!cfg.ld) {
      G.start();
      if (!GRAD) {
#pragma unroll 8
        for (int p = 0; p < NP; ++p) G.add(T.pt(i, p));
      } else {
        // area and tangent in one pass, the next vertex in flight while this one's implicit-function step runs
        GT.start();
        cd zn = T.pt(i, 0);
        for (int p = 0; p < NP; ++p) {
          const cd z = zn;
          if (p + 1 < NP) zn = T.pt(i, p + 1);
          G.add(z);
          GT.add(z);
        }
        GT.close(par, dtot);
      }
      total += par * G.close();
    }
  }

  if (closed != (1u << D) - 1u) {
    // ---- split the open tracks into runs of real images of one parity without jumps > 0.1
    const int nseg_max = 3 * (cfg.nl * cfg.nl + 1);
    Seg parts_local[MAXSEG];
    const Seg* parts = pre_parts;
    int nparts_total = pre_nparts;
    if (!pre_parts) {
      nparts_total = build_parts_serial<D>(cfg, T, closed, parts_local);
      parts = parts_local;
    }
    // ---- stitch (extended_source.py:495-667): three rounds; the active chain starts from the
    // shortest remaining segment and grows by the closest admissible connection
    static_assert(MAXSEG <= 32, "the live-segment set is one 32-bit mask");
    unsigned alive = nparts_total >= 32 ? 0xffffffffu : (1u << nparts_total) - 1u;
    int8_t chain_local[128];
    int8_t* const cmem = chain_mem ? chain_mem : chain_local;
    int pool_size = nseg_max;                 // fixed-shape pool incl. empty slots, shrinks by one per round
    int max_in = 20;
    for (int round = 0; round < 3; ++round, max_in -= 2) {
      int a = -1;
      for (int k = 0; k < nparts_total; ++k)
        if (((alive >> k) & 1u) && parts[k].len != 0.0 && (a < 0 || parts[k].len < parts[a].len)) a = k;
      --pool_size;
      if (a < 0) break;                       // only empty segments left: the remaining contours are empty
      alive &= ~(1u << a);
      Chain act; act.init(cmem, a, seg_n(parts[a]));
      for (int step = 0; step < max_in; ++step) {
        const int nreal = __popc(alive);
        if (nreal == 0) break;
        const int nempty = pool_size - nreal;
        const cd ah = chain_pt(T, parts, act, 0, false), at = chain_pt(T, parts, act, 0, true);
        const double dh0 = norm2(ah), dt0 = norm2(at);   // distance^2 to an empty segment's (0, 0) ends
        // the four closest (connection type, segment) pairs in increasing distance; empty slots
        // take up places in that ranking exactly as the reference's zero rows do
        double lastd = -1.0; int lastc = -1, lastk = -1;
        bool merged = false, exhausted = false;
        for (int rank = 0; rank < 4 && !merged && !exhausted; ) {
          double bd = 1e300; int bc = -1, bk = -1;
          for (int c = 0; c < 4; ++c)
            for (int k = 0; k < nparts_total; ++k) {
              if (!((alive >> k) & 1u)) continue;
              const cd mine = (c == 0 || c == 3) ? at : ah;
              const Seg& g = parts[k];
              const cd theirs = (c == 0 || c == 2) ? T.pt(g.track, g.lo) : T.pt(g.track, g.hi - 1);
              const double d2 = norm2(mine - theirs);
              const bool after_last = d2 > lastd || (d2 == lastd && (c > lastc || (c == lastc && k > lastk)));
              if (after_last && d2 < bd) { bd = d2; bc = c; bk = k; }
            }
          if (bc < 0) { exhausted = true; break; }
          // empty slots closer than this candidate: T-H and T-T at |tail|, H-T and H-H at |head|
          int ahead = 0;
          if (nempty > 0) ahead = (dt0 < bd ? 2 * nempty : 0) + (dh0 < bd ? 2 * nempty : 0);
          // (rank counts real candidates already examined)
          if (rank + ahead >= 4) { exhausted = true; break; }
          const OneSeg other{bk, seg_n(parts[bk])};
          if (connect_ok(T, parts, act, other, bc)) {
            const int n2 = seg_n(parts[bk]);
            if (bc == 0) { act.seg[act.tail] = (int8_t)bk; act.rev[act.tail] = 0; ++act.tail; }
            else if (bc == 1) { --act.head; act.seg[act.head] = (int8_t)bk; act.rev[act.head] = 0; }
            else if (bc == 2) { --act.head; act.seg[act.head] = (int8_t)bk; act.rev[act.head] = 1; }
            else { act.seg[act.tail] = (int8_t)bk; act.rev[act.tail] = 1; ++act.tail; }
            act.npts += n2;
            alive &= ~(1u << bk);
            merged = true;
          }
          lastd = bd; lastc = bc; lastk = bk; ++rank;
        }
        if (!merged) break;                    // nothing can change in the remaining scan steps
      }
      // close the contour (:724-725) and integrate
      const double par = chain_parity(parts, act);
      if (emit) E.start();
      G.start();
      if (GRAD) GT.start();
      for (int q = act.head; q < act.tail; ++q) {
        const Seg& g = parts[act.seg[q]];
        const int n = seg_n(g);
        for (int k = 0; k < n; ++k) {
          const int pidx = act.rev[q] ? g.hi - 1 - k : g.lo + k;
          const cd z = T.pt(g.track, pidx);
          if (emit) E.add(z, T.th(pidx));
          G.add(z);
          if (GRAD) GT.add(z);
        }
      }
      if (emit) E.close(par);
      total += par * G.close();
      if (GRAD) GT.close(par, dtot);
    }
  }
  if (emit) {
    b.vcount[s] = E.nv < cfg.VMAX ? E.nv : cfg.VMAX;
    b.cstart[(int64_t)(E.nc < cfg.CMAX ? E.nc : cfg.CMAX) * cfg.S + s] = E.nv < cfg.VMAX ? E.nv : cfg.VMAX;
    b.ncont[s] = E.nc < cfg.CMAX ? E.nc : cfg.CMAX;
  }
  if (!cfg.ld && b.mag) b.mag[out_idx] = fabs(total) * norm;
  if (GRAD) {
    // mag = |A| / (pi rho^2): d mag = sign(A) dA / (pi rho^2), and -2 mag / rho more for rho itself
    const double sg = total < 0.0 ? -norm : norm;
    const int64_t n = cfg.ngrad_stride;
#pragma unroll
    for (int k = 0; k < NGRAD; ++k)
      b.grad[(int64_t)k * n + out_idx] = sg * dtot[k] - (k == 7 ? 2.0 * fabs(total) * norm / cfg.rho : 0.0);
  }
}

// ---------------------------------------------------------------------------------------------
// Limb darkening (integrate.py:29-121).  I(r) = 3/(3-u1) (u1 B(r) + 1 - 2 u1), B = 1 + sqrt(1-r^2)
// inside the disk, 1 - sqrt(1 - 1/r^2) outside.
template <int NL>
__device__ __forceinline__ double brightness(const LensConst& L, cd z, cd w0, double inv_rho2, double u1) {
  // lens equation with one-Newton reciprocals (2^-44): this is a quadrature integrand
  cd s1 = mk(0, 0);
  const cd zb = conj(z);
#pragma unroll
  for (int j = 0; j < (NL == 1 ? 1 : NL); ++j) {
    const cd d = NL == 1 ? zb : zb - conj(L.r[j]);
    const double inv = (NL == 1 ? 1.0 : L.eps[j]) * rcp_fast1(norm2(d));
    s1 = mk(fma(d.re, inv, s1.re), fma(-d.im, inv, s1.im));
  }
  const double r2 = norm2((z - s1) - w0) * inv_rho2;
  // one branch-free square root for both sides of the limb (integrate.py:37-43); sqrt_fast returns 0
  // for a non-positive argument, which is the reference's clamp
  const bool inside = r2 <= 1.0;
  const double q = sqrt_fast(inside ? 1.0 - r2 : 1.0 - rcp_fast1(r2));
  const double B = inside ? 1.0 + q : 1.0 - q;
  return 3.0 / (3.0 - u1) * (u1 * B + 1.0 - 2.0 * u1);
}
// two Gauss-Legendre panels [a, split] (n1 nodes) and [split, b] (n2 nodes), integrate.py:56-75.
// `along_y`: integrate I(x + i t) dt, else I(t + i y) dt.
template <int NL>
__device__ double two_panel(const ExtCfg& cfg, const ExtBuf& bf, const LensConst& L, double a, double b,
                            double fixed, bool along_y, cd w0) {
  const double rho = cfg.rho, inv_rho2 = 1.0 / (rho * rho);
  const double ad = fabs(b - a);
  double split = b > a ? b - 2.0 * rho : b + 2.0 * rho;
  if (0.5 * ad <= 2.0 * rho) split = a + 0.5 * ad;   // NB lies outside [b, a] when b < a (App. C-10)
  double tot = 0.0;
  for (int panel = 0; panel < 2; ++panel) {
    const double lo = panel == 0 ? a : split, hi = panel == 0 ? split : b;
    int n = panel == 0 ? cfg.n1 : cfg.n2, off = panel == 0 ? 0 : cfg.n1;
    if (cfg.ld_adapt && panel == 0) {
      // SURVEY 8 f4 (opt-in, default off so that parity with integrate.py:47-121 holds): the panel next to
      // the vertex ends ON the limb, where the brightness has its square-root edge, and keeps its n2 nodes;
      // a far panel that stays >= 2 rho away from the vertex and is only a few source radii long is
      // integrated by the half-order rule.
      const double len = fabs(split - a);
      const int nh = cfg.n1 / 2 > 2 ? cfg.n1 / 2 : 2, nq = cfg.n1 / 4 > 2 ? cfg.n1 / 4 : 2;
      const bool far = 0.5 * ad > 2.0 * rho;        // the far panel stays >= 2 rho away from the vertex
      if (far && len <= CB200_ADAPT_Q * rho) { n = nq; off = cfg.n1 + cfg.n2 + nh; }
      else if (far && len <= CB200_ADAPT_H * rho) { n = nh; off = cfg.n1 + cfg.n2; }
    }
    const double hw = 0.5 * (hi - lo), mid = 0.5 * (hi + lo);
    double acc = 0.0;
    for (int k = 0; k < n; ++k) {
      const double t = fma(hw, bf.glx[off + k], mid);
      const cd z = along_y ? mk(fixed, t) : mk(t, fixed);
      acc += (hw * brightness<NL>(L, z, w0, inv_rho2, cfg.u1)) * bf.glw[off + k];
    }
    tot += acc;
  }
  return tot;
}

template <int NL>
__device__ void ld_pq_item(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L, int v, int64_t s);
template <int NL>
__device__ void ld_pq_body(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L, int64_t g) {
  const int v = (int)(g / cfg.S);
  ld_pq_item<NL>(cfg, b, L, v, g - (int64_t)v * cfg.S);
}
template <int NL>
__device__ void ld_pq_item(const ExtCfg& cfg, const ExtBuf& b, const LensConst& L, int v, int64_t s) {
  if (v >= cfg.VMAX || s >= nsrc(cfg, b)) return;
  if (v >= b.vcount[s]) return;
  const cd w0 = source_centre(cfg, b, L, s);
  const cb200_d2 zz = b.vz[(int64_t)v * cfg.S + s];
  const int c = b.vcid[(int64_t)v * cfg.S + s];
  const cb200_d2 z0 = b.cz0[(int64_t)c * cfg.S + s];
  // P = -1/2 int_{y0}^{y} I(x + i y') dy',  Q = +1/2 int_{x0}^{x} I(x' + i y) dx'
  b.vP[(int64_t)v * cfg.S + s] = -0.5 * two_panel<NL>(cfg, b, L, z0.y, zz.y, zz.x, true, w0);
  b.vQ[(int64_t)v * cfg.S + s] = 0.5 * two_panel<NL>(cfg, b, L, z0.x, zz.x, zz.y, false, w0);
}

__device__ void ld_sum_body(const ExtCfg& cfg, const ExtBuf& b, int64_t s) {
  if (s >= nsrc(cfg, b)) return;
  const int nc = b.ncont[s];
  double total = 0.0;
  for (int c = 0; c < nc; ++c) {
    const int v0 = b.cstart[(int64_t)c * cfg.S + s], v1 = b.cstart[(int64_t)(c + 1) * cfg.S + s];
    double acc = 0.0;
    if (v1 > v0) {
      cb200_d2 za = b.vz[(int64_t)v0 * cfg.S + s];
      double Pa = b.vP[(int64_t)v0 * cfg.S + s], Qa = b.vQ[(int64_t)v0 * cfg.S + s];
#pragma unroll 8
      for (int v = v0 + 1; v < v1; ++v) {
        const cb200_d2 zb = b.vz[(int64_t)v * cfg.S + s];
        const double Pb = b.vP[(int64_t)v * cfg.S + s], Qb = b.vQ[(int64_t)v * cfg.S + s];
        acc += 0.5 * (Pa + Pb) * (zb.x - za.x) + 0.5 * (Qa + Qb) * (zb.y - za.y);
        za = zb; Pa = Pb; Qa = Qb;
      }
    }
    total += acc * b.cpar[(int64_t)c * cfg.S + s];
  }
  const int64_t out_idx = src_index(b, s);
  b.mag[out_idx] = fabs(total) / (3.14159265358979323846 * cfg.rho * cfg.rho);
}

// ---------------------------------------------------------------------------------------------
// `mag` gate for the binary lens (lightcurve.py:202-225): point-source images, hexadecapole value
// and the validity tests; points that fail are appended to the compact list for full integration.
__device__ __forceinline__ cd cpow3(cd a) { return a * a * a; }

template <bool COMP, int NT>
__device__ void gate_body(const cb200_d2* __restrict__ w_in, double* __restrict__ mag, uint8_t* __restrict__ test_out,
                          int32_t* __restrict__ list, int32_t* __restrict__ count, int64_t n, const LensConst& L,
                          double rho, double q, int itmax, EASmem<5, COMP, NT>& sm, int tid, int64_t idx) {
  constexpr int D = 5;
  const bool active = idx < n;
  cd w = mk(0.3, 0.2);
  if (active) { const cb200_d2 v = w_in[idx]; w = mk(v.x + L.x_cm, v.y); }
  cd p[D + 1];
  lens_poly<2>(L, w, p);
  ea_normalise<D>(p);
  ea_solve_thread<D, COMP, NT>(p, sm, tid, active, false, EA_INIT_BINI, itmax, true);   // order-independent sums
  if (!active) return;
  const double a = L.r[0].re, e1 = L.eps[0], e2 = L.eps[1];
  const double rr = rho + 1e-3;  // rho + rho_min
  double mu = 0.0, dmu = 0.0, mu_cusp = 0.0, ffalse = 0.0;
  int nfalse = 0;
#pragma unroll 1
  for (int j = 0; j < D; ++j) {
    const cd z = mk(sm.zre[j][tid], sm.zim[j][tid]);
    bool real_image; double detj;
    image_eval<2>(L, z, w, real_image, detj);
    // derivatives of f(z) = -e1/(z-a) - e2/(z+a), lightcurve.py:29-33
    auto fp = [&](cd x) { const cd u = crecip(x - mk(a, 0)), v = crecip(x + mk(a, 0)); return e1 * (u * u) + e2 * (v * v); };
    auto fpp = [&](cd x) { const cd u = crecip(mk(a, 0) - x), v = crecip(mk(a, 0) + x); return 2.0 * (e1 * cpow3(u) - e2 * cpow3(v)); };
    const cd zb = conj(z);
    const cd fz = -(e1 * crecip(z - mk(a, 0))) - e2 * crecip(z + mk(a, 0));
    const cd zhat = conj(w) - fz;
    const cd fp_z = fp(z), fpp_z = fpp(z), fp_zb = fp(zb), fp_zh = fp(zhat), fpp_zb = fpp(zb);
    const double J = 1.0 - sqrt(norm2(fp_z * fp_zb));
    if (real_image) {
      double m0, dq, dh;
      hexadecapole_terms<2>(L, z, rho, 0.0, m0, dq, dh);   // u1 is not forwarded
      mu += fabs(m0 + dq + dh);
      dmu += fabs(dq) + fabs(dh);
      const cd t = 3.0 * (cpow3(fp_zb) * (fpp_z * fpp_z));
      const double J5 = J * J * J * J * J;
      mu_cusp += fabs(6.0 * t.im / J5 * (rr * rr));
    } else {
      const double Jh = 1.0 - sqrt(norm2(fp_z * fp_zh));
      const cd den = (Jh * fpp_zb) * fp_z - (Jh * fpp_z) * (fp_zb * fp_zh);   // conj(Jhat) = Jhat (real)
      ffalse += fabs(J * Jh * Jh) * rsqrt(norm2(den));
      ++nfalse;
    }
  }
  bool ok = (0.02 * mu_cusp + dmu < 1e-2) && (nfalse == 0 || 0.5 * ffalse > 4.0 * rr);
  if (q < 0.01) {
    // planetary-caustic test, lightcurve.py:78-85
    const double s_ = 2.0 * a, qq = e1 / (1.0 - e1);
    const double wpc = -1.0 / s_, dpc = 3.0 * sqrt(qq) / s_;
    const double dx = wpc - w.re, dy = -w.im;
    ok = ok && (dx * dx + dy * dy > 2.0 * (rho * rho + dpc * dpc));
  }
  mag[idx] = mu;
  if (test_out) test_out[idx] = ok ? 1 : 0;
  if (!ok) list[cb200_atomic_inc(count)] = (int32_t)idx;
}


}  // namespace cb200
