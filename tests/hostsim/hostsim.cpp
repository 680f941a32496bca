// TEST INFRASTRUCTURE ONLY -- see cuda_shim.h.  Host build of the device solver / lens headers.
#include "cuda_shim.h"
#include "../../caustics_b200/csrc/ea_core.cuh"
#include "../../caustics_b200/csrc/lens_core.cuh"
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
  switch (deg) { C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) }
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
