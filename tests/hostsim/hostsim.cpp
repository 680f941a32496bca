// TEST INFRASTRUCTURE ONLY -- see cuda_shim.h.  Host build of the device solver / lens headers.
#include "cuda_shim.h"
#include "../../caustics_b200/csrc/ea_core.cuh"
#include "../../caustics_b200/csrc/lens_core.cuh"
#include "../../caustics_b200/csrc/ps_walk.cuh"
using namespace cb200;

template <int DEG, bool COMP>
static void solve_all(const double* coeffs, const double* ri, double* roots, int32_t* sweeps, int64_t size,
                      int itmax, int custom_init, int init_mode) {
  static EASmem<DEG, COMP, 1> sm;
  for (int64_t n = 0; n < size; ++n) {
    cd p[DEG + 1];
    for (int k = 0; k <= DEG; ++k) p[k] = mk(coeffs[2 * (n * (DEG + 1) + k)], coeffs[2 * (n * (DEG + 1) + k) + 1]);
    ea_normalise<DEG>(p);
    if (custom_init)
      for (int j = 0; j < DEG; ++j) { sm.zre[j][0] = ri[2 * (n * DEG + j)]; sm.zim[j][0] = ri[2 * (n * DEG + j) + 1]; }
    EAResult r = ea_solve_thread<DEG, COMP, 1>(p, sm, 0, true, custom_init != 0, init_mode, itmax);
    for (int j = 0; j < DEG; ++j) { roots[2 * (n * DEG + j)] = sm.zre[j][0]; roots[2 * (n * DEG + j) + 1] = sm.zim[j][0]; }
    if (sweeps) sweeps[n] = r.converged ? r.sweeps : -r.sweeps;
  }
}

template <int NL, bool COMP>
static void mag_all(const double* w_in, double* mag, double* coeffs_out, int64_t n, LensConst L, int itmax, int init_mode) {
  constexpr int DEG = NL * NL + 1;
  static EASmem<DEG, COMP, 1> sm;
  for (int64_t i = 0; i < n; ++i) {
    cd w = mk(w_in[2 * i] + L.x_cm, w_in[2 * i + 1]);
    cd p[DEG + 1];
    lens_poly<NL>(L, w, p);
    if (coeffs_out) for (int k = 0; k <= DEG; ++k) { coeffs_out[2 * (i * (DEG + 1) + k)] = p[k].re; coeffs_out[2 * (i * (DEG + 1) + k) + 1] = p[k].im; }
    ea_normalise<DEG>(p);
    ea_solve_thread<DEG, COMP, 1>(p, sm, 0, true, false, init_mode, itmax);
    double mu = 0;
    for (int j = 0; j < DEG; ++j) {
      bool real; double det;
      image_eval<NL>(L, mk(sm.zre[j][0], sm.zim[j][0]), w, real, det);
      if (real) mu += 1.0 / fabs(det);
    }
    mag[i] = mu;
  }
}

extern "C" {
int hostsim_ea_solve(const double* coeffs, const double* ri, double* roots, int32_t* sweeps, int64_t size, int deg,
                     int itmax, int comp, int custom_init, int init_mode) {
#define C(D) case D: if (comp) solve_all<D, true>(coeffs, ri, roots, sweeps, size, itmax, custom_init, init_mode); \
                     else solve_all<D, false>(coeffs, ri, roots, sweeps, size, itmax, custom_init, init_mode); return 0;
  switch (deg) { C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(13) C(16) }
#undef C
  return 2;
}
// L: nlenses, eps[3], r[3] (re,im), H[4], G[3], x_cm packed by the caller as doubles
int hostsim_mag_ps(const double* w, double* mag, double* coeffs_out, int64_t n, int nlenses, const double* eps,
                   const double* r, const double* H, const double* G, double x_cm, int itmax, int comp, int init_mode) {
  LensConst L; memset(&L, 0, sizeof(L));
  L.nlenses = nlenses; L.x_cm = x_cm;
  for (int i = 0; i < 3; ++i) { L.eps[i] = eps[i]; L.r[i] = mk(r[2 * i], r[2 * i + 1]); }
  for (int i = 0; i < 4; ++i) L.H[i] = mk(H[2 * i], H[2 * i + 1]);
  for (int i = 0; i < 3; ++i) L.G[i] = mk(G[2 * i], G[2 * i + 1]);
  if (nlenses == 2) { if (comp) mag_all<2, true>(w, mag, coeffs_out, n, L, itmax, init_mode); else mag_all<2, false>(w, mag, coeffs_out, n, L, itmax, init_mode); }
  else if (nlenses == 3) { if (comp) mag_all<3, true>(w, mag, coeffs_out, n, L, itmax, init_mode); else mag_all<3, false>(w, mag, coeffs_out, n, L, itmax, init_mode); }
  else return 1;
  return 0;
}
}

// the map walk of ps_walk.cuh, one column at a time (the kernel's CTA/lanes only decide which thread
// owns which column segment)
template <int NL, bool COMP>
static void grid_walk_all(double x0, double y0, double dx, double dy, int64_t nx, int64_t row_begin, int64_t row_end,
                          double* mag, const LensConst& L, int itmax, int run, int extrap) {
  constexpr int DEG = NL * NL + 1;
  static EASmem<DEG, COMP, 1> sm;
  double pre[DEG], pim[DEG];
  const int64_t nrows = row_end - row_begin;
  for (int64_t row0 = 0; row0 < nrows; row0 += run)
    for (int64_t ix = 0; ix < nx; ++ix) {
      const int nrun = (int)std::min<int64_t>(run, nrows - row0);
      WalkColumn src;
      src.wx = fma((double)ix, dx, x0) + L.x_cm; src.y0 = y0; src.dy = dy; src.row_abs0 = row_begin + row0;
      ps_walk_body<NL, COMP, 1>(src, nrun, nrun, mag + row0 * nx + ix, nx, L, itmax, extrap != 0, sm, pre, pim, 0);
    }
}
template <int NL, bool COMP>
static void path_walk_all(const double* w, double* mag, int64_t n, const LensConst& L, int itmax, int run, int extrap) {
  constexpr int DEG = NL * NL + 1;
  static EASmem<DEG, COMP, 1> sm;
  double pre[DEG], pim[DEG];
  for (int64_t first = 0; first < n; first += run) {
    WalkPath src; src.w = w + 2 * first; src.x_cm = L.x_cm;
    ps_walk_body<NL, COMP, 1>(src, run, (int)std::min<int64_t>(run, n - first), mag + first, 1, L, itmax, extrap != 0, sm, pre, pim, 0);
  }
}
extern "C" int hostsim_path_walk(const double* w, double* mag, int64_t n, int nlenses, const double* eps, const double* r,
                                 const double* H, const double* G, double x_cm, int itmax, int comp, int run, int extrap) {
  LensConst L; memset(&L, 0, sizeof(L));
  L.nlenses = nlenses; L.x_cm = x_cm;
  for (int i = 0; i < 3; ++i) { L.eps[i] = eps[i]; L.r[i] = mk(r[2 * i], r[2 * i + 1]); }
  for (int i = 0; i < 4; ++i) L.H[i] = mk(H[2 * i], H[2 * i + 1]);
  for (int i = 0; i < 3; ++i) L.G[i] = mk(G[2 * i], G[2 * i + 1]);
  if (nlenses == 2) { if (comp) path_walk_all<2, true>(w, mag, n, L, itmax, run, extrap); else path_walk_all<2, false>(w, mag, n, L, itmax, run, extrap); }
  else if (nlenses == 3) { if (comp) path_walk_all<3, true>(w, mag, n, L, itmax, run, extrap); else path_walk_all<3, false>(w, mag, n, L, itmax, run, extrap); }
  else return 1;
  return 0;
}
extern "C" int hostsim_grid_walk(double x0, double y0, double dx, double dy, int64_t nx, int64_t row_begin, int64_t row_end,
                                 double* mag, int nlenses, const double* eps, const double* r, const double* H,
                                 const double* G, double x_cm, int itmax, int comp, int run, int extrap) {
  LensConst L; memset(&L, 0, sizeof(L));
  L.nlenses = nlenses; L.x_cm = x_cm;
  for (int i = 0; i < 3; ++i) { L.eps[i] = eps[i]; L.r[i] = mk(r[2 * i], r[2 * i + 1]); }
  for (int i = 0; i < 4; ++i) L.H[i] = mk(H[2 * i], H[2 * i + 1]);
  for (int i = 0; i < 3; ++i) L.G[i] = mk(G[2 * i], G[2 * i + 1]);
  if (nlenses == 2) { if (comp) grid_walk_all<2, true>(x0, y0, dx, dy, nx, row_begin, row_end, mag, L, itmax, run, extrap); else grid_walk_all<2, false>(x0, y0, dx, dy, nx, row_begin, row_end, mag, L, itmax, run, extrap); }
  else if (nlenses == 3) { if (comp) grid_walk_all<3, true>(x0, y0, dx, dy, nx, row_begin, row_end, mag, L, itmax, run, extrap); else grid_walk_all<3, false>(x0, y0, dx, dy, nx, row_begin, row_end, mag, L, itmax, run, extrap); }
  else return 1;
  return 0;
}

// ---- kernel family 3 (extended source): the phase bodies driven sequentially over the sources ----
#include "../../caustics_b200/csrc/extended_host.h"
#include <vector>

static void fill_lens(LensConst& L, int nlenses, const double* eps, const double* r, const double* H,
                      const double* G, double x_cm) {
  memset(&L, 0, sizeof(L));
  L.nlenses = nlenses; L.x_cm = x_cm;
  for (int i = 0; i < 3; ++i) { L.eps[i] = eps[i]; L.r[i] = mk(r[2 * i], r[2 * i + 1]); }
  for (int i = 0; i < 4; ++i) L.H[i] = mk(H[2 * i], H[2 * i + 1]);
  for (int i = 0; i < 3; ++i) L.G[i] = mk(G[2 * i], G[2 * i + 1]);
}

template <int NL>
static void ext_pipeline(const ExtCfg& cfg, ExtBuf b, const LensConst& L, int64_t ns) {
  constexpr int D = NL == 1 ? 2 : NL * NL + 1;
  constexpr int NLS = NL == 1 ? 2 : NL;   // solver instantiation (unused for the single lens)
  static EASmem<NLS * NLS + 1, false, 1> sm0;
  static EASmem<NLS * NLS + 1, true, 1> sm1;
  for (int64_t s = 0; s < ns; ++s) {
    if (NL == 1) limb_walk_single_body(cfg, b, L, s); else limb_walk_body<NLS, 1>(cfg, b, L, sm0, 0, s);
  }
  {
    // the refinement: selection for every source, then every (source, new point), round by round
    static EASmem<D, false, 1> rf0;
    static EASmem<D, true, 1> rf1;
    for (int round = 0; round < NITER; ++round) {
      for (int64_t s = 0; s < ns; ++s) round_select_body<D>(cfg, b, round, 0, s);
      for (int64_t g = 0; g < ns * cfg.nadd; ++g) {
        if (cfg.comp && NL != 1) round_solve_body<NL, true, 1>(cfg, b, L, rf1, 0, g, round);
        else round_solve_body<NL, false, 1>(cfg, b, L, rf0, 0, g, round);
      }
    }
  }
  if (!cfg.ld && !cfg.tracks) {
    // plain uniform disk: one pass, then the listed caustic-crossing sources
    *b.open_count = 0;
    cb200_d2 stage[D];
    for (int64_t s = 0; s < ns; ++s) sweep_body<D>(cfg, b, s, stage, 1);
    for (int32_t g = 0; g < *b.open_count; ++g) contours_body<D, false>(cfg, b, L, b.open_list[g], nullptr, true);
    return;
  }
  cb200_d2 stage2[D];
  for (int64_t s = 0; s < ns; ++s) tracks_body<D>(cfg, b, s, stage2, 1);
  for (int64_t s = 0; s < ns; ++s) {
    if (b.grad) contours_body<D, true>(cfg, b, L, s); else contours_body<D>(cfg, b, L, s);
  }
  if (cfg.ld) {
    for (int64_t g = 0; g < (int64_t)cfg.VMAX * cfg.S; ++g) ld_pq_body<NL>(cfg, b, L, g);
    for (int64_t s = 0; s < ns; ++s) ld_sum_body(cfg, b, s);
  }
}


// optional: where the next hostsim_mag_extended call writes d mag / d(a, e1, e2, Re r3, Im r3, Re w, Im w, rho), (8, n)
static double* g_hostsim_grad = nullptr;
extern "C" void hostsim_set_grad(double* out) { g_hostsim_grad = out; }
// 1: uniform-disk calls also take the track-array path (tracks_body + contours_body) instead of sweep_body / open pass
static int g_hostsim_force_tracks = 0;
extern "C" void hostsim_force_tracks(int on) { g_hostsim_force_tracks = on; }

extern "C" int hostsim_mag_extended(const double* w, double* mag, uint8_t* test_out, int64_t n, double rho, int nlenses,
                                    const double* eps, const double* r, const double* H, const double* G, double x_cm,
                                    double q, int gate, int npts_limb, int ld, double u1, int npts_ld, int itmax, int comp,
                                    double* tracks_out /* optional (NP, D, 2) of source 0 */, uint8_t* flags_out) {
  ExtCfg cfg;
  int rc = make_cfg(n, rho, nlenses, npts_limb, ld, u1, npts_ld, itmax, comp, &cfg);
  if (rc) return rc;
  cfg.tracks = (tracks_out || g_hostsim_grad || g_hostsim_force_tracks) ? 1 : 0;
  Layout lay = make_layout(cfg);
  std::vector<char> ws(lay.total + 256, 0);
  LensConst L;
  fill_lens(L, nlenses, eps, r, H, G, x_cm);
  if (nlenses == 1) { L.eps[0] = 1.0; L.x_cm = 0.0; }
  ExtBuf b = bind(cfg, lay, ws.data());
  b.w = (const cb200_d2*)w;
  b.mag = mag;
  b.grad = ld ? nullptr : g_hostsim_grad;
  cfg.ngrad_stride = n;
  for (int t = 0; t < cfg.D * cfg.nadd; ++t) jitter_table_body(cfg.D, cfg.nadd, (double*)(ws.data() + lay.jit), t);
  int32_t* list = (int32_t*)(ws.data() + lay.list);
  int32_t* count = (int32_t*)(ws.data() + lay.count);
  if (cfg.ld) {
    fill_gl_tables(cfg.n1, cfg.n2, (double*)(ws.data() + lay.gl));
  }
  int64_t ns = n;
  if (gate && nlenses == 2) {
    *count = 0;
    static EASmem<5, false, 1> g0;
    static EASmem<5, true, 1> g1;
    for (int64_t i = 0; i < n; ++i) {
      if (comp) gate_body<true, 1>((const cb200_d2*)w, mag, test_out, list, count, n, L, rho, q, itmax, g1, 0, i);
      else gate_body<false, 1>((const cb200_d2*)w, mag, test_out, list, count, n, L, rho, q, itmax, g0, 0, i);
    }
    b.list = list; b.count = count; ns = *count;
  }
  switch (nlenses) {
    case 1: ext_pipeline<1>(cfg, b, L, ns); break;
    case 2: ext_pipeline<2>(cfg, b, L, ns); break;
    default: ext_pipeline<3>(cfg, b, L, ns); break;
  }
  if (tracks_out)
    for (int p = 0; p < cfg.NP; ++p)
      for (int j = 0; j < cfg.D; ++j) {
        tracks_out[2 * (p * cfg.D + j)] = b.sre[((int64_t)p * cfg.D + j) * cfg.S];
        tracks_out[2 * (p * cfg.D + j) + 1] = b.sim[((int64_t)p * cfg.D + j) * cfg.S];
        if (flags_out) flags_out[p * cfg.D + j] = b.sflg[((int64_t)p * cfg.D + j) * cfg.S];
      }
  return 0;
}

#ifdef CB200_HOSTSIM_COUNT
extern "C" void hostsim_counters(long long* out, int reset) {
  out[0] = cb200::g_ea_evals; out[1] = cb200::g_ea_updates;
  if (reset) { cb200::g_ea_evals = 0; cb200::g_ea_updates = 0; }
}
#endif

// the device code's jitter stream (csrc/jax_prng.cuh): limb (deg, n) complex, dup (deg, npts) real
extern "C" void hostsim_jitters(int deg, int n, int npts, double* limb, double* dup) {
  for (int j = 0; j < deg; ++j) {
    for (int r = 0; r < n; ++r) {
      const cd v = limb_jitter(j, r, deg, n);
      limb[2 * (j * n + r)] = v.re;
      limb[2 * (j * n + r) + 1] = v.im;
    }
    for (int p = 0; p < npts; ++p) dup[j * npts + p] = duplicate_jitter(j, p, deg, npts);
  }
}
