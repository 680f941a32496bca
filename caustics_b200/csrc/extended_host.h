// Host-side configuration / workspace layout of kernel family 3, shared by extended.cu (the CUDA
// driver) and the host logic tests (tests/hostsim).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/caustics_b200.h"
#include "extended_core.cuh"

namespace cb200 {

inline void leggauss(int n, double* x, double* w) {
  // Newton iteration on P_n; nodes ascending like numpy.polynomial.legendre.leggauss
  for (int i = 0; i < n; ++i) {
    double t = cos(3.14159265358979323846 * (i + 0.75) / (n + 0.5));
    double dp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = t;
      for (int k = 2; k <= n; ++k) { const double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = p2; }
      if (n == 1) { p0 = 1.0; p1 = t; }
      dp = n * (t * p1 - p0) / (t * t - 1.0);
      const double dt = p1 / dp;
      t -= dt;
      if (fabs(dt) < 1e-16) break;
    }
    double p0 = 1.0, p1 = t;
    for (int k = 2; k <= n; ++k) { const double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = p2; }
    dp = n * (t * p1 - p0) / (t * t - 1.0);
    x[n - 1 - i] = t;
    w[n - 1 - i] = 2.0 / ((1.0 - t * t) * dp * dp);
  }
}

// Gauss-Legendre tables of a limb-darkened call, nodes then weights, each laid out as
// [n1 | n2 | nh = max(2, n1/2) | nq = max(2, n1/4)]: the two panels of integrate.py:56-75 and the two reduced
// orders of the opt-in adaptive far panel (ExtCfg::ld_adapt).  Returns the number of nodes (= weights).
inline int gl_table_nodes(int n1, int n2) {
  const int nh = n1 / 2 > 2 ? n1 / 2 : 2, nq = n1 / 4 > 2 ? n1 / 4 : 2;
  return n1 + n2 + nh + nq;
}
inline int fill_gl_tables(int n1, int n2, double* tab) {
  const int nh = n1 / 2 > 2 ? n1 / 2 : 2, nq = n1 / 4 > 2 ? n1 / 4 : 2;
  const int nn = n1 + n2 + nh + nq;
  leggauss(n1, tab, tab + nn);
  leggauss(n2, tab + n1, tab + nn + n1);
  leggauss(nh, tab + n1 + n2, tab + nn + n1 + n2);
  leggauss(nq, tab + n1 + n2 + nh, tab + nn + n1 + n2 + nh);
  return nn;
}

struct Layout {
  size_t theta, z, fw, order, rnext, rlr, sre, sim, sflg, perm, sw_total, sw_closed, open_list, open_count, vz, vP, vQ, vcid, vcount, ncont, cz0, cpar, cstart, gl, jit, list, count, total;
};
inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// `npoints`: capacity of the gate's compact list (all points of the call); the per-source arrays are
// sized by c.S, the number of sources integrated at a time.
inline Layout make_layout(const ExtCfg& c, int64_t npoints = -1) {
  Layout l; size_t o = 0;
  const size_t S = (size_t)c.S, NP = c.NP, D = c.D;
  const size_t NL = npoints > (int64_t)S ? (size_t)npoints : S;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  l.theta = take(NP * S * 8);
  l.z = take(NP * D * S * 16); l.fw = take(NP * S * 4);
  l.order = take(NP * S * 2);
  l.rnext = take(NP * S * 2); l.rlr = take((size_t)2 * NADD_MAX * S * 2);
  // the theta-ordered track arrays are needed by limb-darkened, tangent and export calls only; the plain
  // uniform-disk call integrates in one pass (sweep_body) and keeps 8 bytes per limb point instead
  if (c.ld || c.tracks) {
    l.sre = take(NP * D * S * 8); l.sim = take(NP * D * S * 8); l.sflg = take(NP * D * S);
    l.perm = l.sw_total = l.sw_closed = l.open_list = l.open_count = 0;
  } else {
    l.sre = l.sim = l.sflg = 0;
    l.perm = take(NP * S * 8); l.sw_total = take(S * 8); l.sw_closed = take(S * 4); l.open_list = take(S * 4); l.open_count = take(256);
  }
  if (c.ld) {
    l.vz = take((size_t)c.VMAX * S * 16); l.vP = take((size_t)c.VMAX * S * 8); l.vQ = take((size_t)c.VMAX * S * 8);
    l.vcid = take((size_t)c.VMAX * S); l.vcount = take(S * 4); l.ncont = take(S * 4);
    l.cz0 = take((size_t)c.CMAX * S * 16); l.cpar = take((size_t)c.CMAX * S * 8); l.cstart = take((size_t)(c.CMAX + 1) * S * 4);
    l.gl = take((size_t)gl_table_nodes(c.n1, c.n2) * 16);
  } else { l.vz = l.vP = l.vQ = l.vcid = l.vcount = l.ncont = l.cz0 = l.cpar = l.cstart = l.gl = 0; }
  l.jit = take((size_t)NADD_MAX * 10 * 16);
  l.list = take(NL * 4); l.count = take(256);
  l.total = o;
  return l;
}

inline int make_cfg(int64_t S, double rho, int nlenses, int npts_limb, int limb_darkening, double u1, int npts_ld,
             int itmax, int compensated, ExtCfg* out) {
  if (S < 0 || !(rho > 0.0) || nlenses < 1 || nlenses > 3 || itmax < 0) return CAUSTICS_ERR_BAD_ARG;
  ExtCfg c; memset(&c, 0, sizeof(c));
  c.nl = nlenses; c.D = nlenses == 1 ? 2 : nlenses * nlenses + 1;
  c.N0 = (int)(0.5 * npts_limb);
  c.nadd = (int)((int)(0.5 * npts_limb) / NITER);
  c.NP = c.N0 + NITER * c.nadd;
  if (c.N0 < 4 || c.nadd < 1 || c.nadd > NADD_MAX || c.NP > 4000) return CAUSTICS_ERR_BAD_ARG;
  c.rho = rho; c.itmax = itmax; c.comp = compensated ? 1 : 0;
  c.ld = (limb_darkening & 3) ? 1 : 0; c.u1 = u1;
  c.ld_adapt = (limb_darkening & 2) ? 1 : 0;
  c.n1 = npts_ld / 2; c.n2 = npts_ld - c.n1;
  if (c.ld && (c.n1 < 1 || npts_ld > 2048)) return CAUSTICS_ERR_BAD_ARG;
  c.CMAX = c.D + 3;
  c.VMAX = c.D * c.NP + c.CMAX;
  c.S = S > 0 ? S : 1;
  *out = c;
  return CAUSTICS_OK;
}

inline ExtBuf bind(const ExtCfg& c, const Layout& l, void* ws) {
  char* base = (char*)ws;
  ExtBuf b; memset(&b, 0, sizeof(b));
  b.theta = (double*)(base + l.theta);
  b.z = (cb200_d2*)(base + l.z); b.fw = (uint32_t*)(base + l.fw);
  b.order = (uint16_t*)(base + l.order);
  b.rnext = (uint16_t*)(base + l.rnext); b.rlr = (uint16_t*)(base + l.rlr);
  if (c.ld || c.tracks) {
    b.sre = (double*)(base + l.sre); b.sim = (double*)(base + l.sim); b.sflg = (uint8_t*)(base + l.sflg);
    b.rdval = b.sre;          // the widths are dead before k_tracks writes the tracks
  } else {
    b.perm = (uint64_t*)(base + l.perm); b.rdval = (double*)(base + l.perm);   // dead before k_sweep writes perm
    b.sw_total = (double*)(base + l.sw_total); b.sw_closed = (uint32_t*)(base + l.sw_closed);
    b.open_list = (int32_t*)(base + l.open_list); b.open_count = (int32_t*)(base + l.open_count);
  }
  if (c.ld) {
    b.vz = (cb200_d2*)(base + l.vz); b.vP = (double*)(base + l.vP); b.vQ = (double*)(base + l.vQ);
    b.vcid = (uint8_t*)(base + l.vcid); b.vcount = (int32_t*)(base + l.vcount); b.ncont = (int32_t*)(base + l.ncont);
    b.cz0 = (cb200_d2*)(base + l.cz0); b.cpar = (double*)(base + l.cpar); b.cstart = (int32_t*)(base + l.cstart);
    b.glx = (const double*)(base + l.gl); b.glw = b.glx + gl_table_nodes(c.n1, c.n2);
  }
  b.jit = (const double*)(base + l.jit);
  return b;
}

}  // namespace cb200
