// Point-source magnification as warm-started walks (opt-in: CAUSTICS_FLAG_GRID_WALK for maps,
// CAUSTICS_FLAG_PATH_WALK for trajectories).
//
// A map pixel's lens polynomial differs from its neighbour's by O(dy), so its roots do too.  The
// reference exploits exactly this along the limb of an extended source and along a 1-D path
// (`custom_init`, extended_source.py:118-125, point_source.py:1711-1759); a regular map offers it
// everywhere.  One thread owns a column segment of `nrun` consecutive rows (lanes run along x, so the
// stores stay coalesced): row 0 of the segment is the usual cold solve (complex Bini estimates), every
// later row starts Ehrlich-Aberth from the linear extrapolation 2 z_k - z_{k-1} of the previous two
// rows' roots.  With the map's step (3e-4 for config C5) the extrapolated guess is O(dy^2) ~ 1e-7
// from the root, Aberth's cubic step lands below the stopping tolerance, and the second evaluation
// only confirms: 2 sweeps instead of the cold start's ~7 (measured on the device code compiled for
// the host, tests/test_hostsim.py::test_grid_walk).  The stopping test is the solver's own
// (|p(z)| <= 2^-53 * sum |p_i||z|^i, ehrlich_aberth.h:109), so every returned root meets the same
// backward-error bound as a cold solve; the image filter and the Jacobian sum are unchanged.
// The same walk serves a 1-D array of source positions whose consecutive elements are neighbours (a
// trajectory; CAUSTICS_FLAG_PATH_WALK): one thread owns a run of consecutive elements.
// A warm solve that fails (no convergence within 64 sweeps, or a non-finite root because two
// extrapolated guesses coincided) is redone cold, which also resets the extrapolation history.
#pragma once
#include "ea_core.cuh"
#include "lens_core.cuh"

namespace cb200 {

constexpr int PS_WALK_WARM_ITMAX = 64;

// pre/pim: this thread's previous-position roots, element j at [j * NT] (only touched when extrap).
// wsrc(k) = source position k of this thread's walk (x_cm already added); mag_out: the thread's first
// output, consecutive positions `out_stride` apart.  `nsteps` is the warp-uniform loop bound, `nrun`
// (<= nsteps) the positions this thread really owns.
template <int NL, bool COMP, int NT, class WSRC>
__device__ __forceinline__ void ps_walk_body(const WSRC& wsrc, int nsteps, int nrun, double* mag_out,
                                             int64_t out_stride, const LensConst& L, int itmax, bool extrap,
                                             EASmem<NL * NL + 1, COMP, NT>& sm, double* pre, double* pim, int tid) {
  constexpr int DEG = NL * NL + 1;
  double* zre = &sm.zre[0][tid];
  double* zim = &sm.zim[0][tid];
  for (int k = 0; k < nsteps; ++k) {
    const bool active = k < nrun;
    const cd w = wsrc(active ? k : 0);
    cd p[DEG + 1];
    lens_poly<NL>(L, w, p);
    ea_normalise<DEG>(p);
    bool redo = active;
    if (k > 0) {
      if (extrap && active) {
#pragma unroll
        for (int j = 0; j < DEG; ++j) {
          const double cr = zre[j * NT], ci = zim[j * NT];
          zre[j * NT] = fma(2.0, cr, -pre[j * NT]);
          zim[j * NT] = fma(2.0, ci, -pim[j * NT]);
          pre[j * NT] = cr;
          pim[j * NT] = ci;
        }
      }
      const EAResult r = ea_solve_thread<DEG, COMP, NT, false>(p, sm, tid, active, true, EA_INIT_BINI,
                                                               itmax < PS_WALK_WARM_ITMAX ? itmax : PS_WALK_WARM_ITMAX);
      double chk = 0.0;
#pragma unroll
      for (int j = 0; j < DEG; ++j) chk += fabs(zre[j * NT]) + fabs(zim[j * NT]);
      redo = active && !(r.converged && chk < 1e300);
    }
    if (__any_sync(0xffffffffu, redo)) {
      ea_solve_thread<DEG, COMP, NT, true, false>(p, sm, tid, redo, false, EA_INIT_BINI, itmax, true);
      if (extrap && redo) {
#pragma unroll
        for (int j = 0; j < DEG; ++j) { pre[j * NT] = zre[j * NT]; pim[j * NT] = zim[j * NT]; }
      }
    }
    if (active) {
      double mu = 0.0;
#pragma unroll
      for (int j = 0; j < DEG; ++j) {
        bool real_image;
        double detj;
        image_eval<NL>(L, mk(zre[j * NT], zim[j * NT]), w, real_image, detj);
        if (real_image) mu += 1.0 / fabs(detj);  // point_source.py:1829
      }
      mag_out[(int64_t)k * out_stride] = mu;
    }
  }
}

// map column: position k = (wx, y0 + (row_abs0 + k) dy)
struct WalkColumn {
  double wx, y0, dy;
  int64_t row_abs0;
  __device__ __forceinline__ cd operator()(int k) const { return mk(wx, fma((double)(row_abs0 + k), dy, y0)); }
};
// 1-D path segment: position k = w[k] + x_cm
struct WalkPath {
  const double* w;   // (re, im) pairs
  double x_cm;
  __device__ __forceinline__ cd operator()(int k) const { return mk(w[2 * k] + x_cm, w[2 * k + 1]); }
};

}  // namespace cb200
