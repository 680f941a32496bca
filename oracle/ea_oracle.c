/*
 * ea_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * Plain-C99 CPU restatement of the reference's Ehrlich-Aberth root solver, written from the
 * algorithm, not copied: every function cites the reference lines whose behaviour it follows
 * (paths relative to /root/reference).  It is pinned against the reference itself
 * (oracle/_ref, the unmodified lib/ehrlich_aberth/cpu_ops.cc) by tests/test_oracle.py and against
 * the committed golden vectors in tests/golden/.
 *
 * Arithmetic conventions that matter for agreeing with the reference to the last few ulps:
 *   - complex division is the scaled form of thrust (lib/extern/thrust-1.15.0/thrust/detail/
 *     complex/arithmetic.h:119-142), |z| is hypot(), complex mul is the naive 4-mul form;
 *   - compiled with -ffp-contract=off: the reference is built for baseline x86-64 (no FMA
 *     contraction); the only fused operations are the explicit fma() calls;
 *   - z^2 in the reversed correction is exp(2 log z) as thrust::pow does
 *     (ehrlich_aberth.h:56-57, thrust/detail/complex/cpow.h:36-44);
 *   - the initial estimates are purely real, r*sin(.), because of the comma expression at
 *     init_est.h:95.  Replicated here so sweep counts match the reference.
 *
 * Build: make -C oracle libea_oracle.so
 */
#include <complex.h>
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

typedef struct { double re, im; } cplx;

static const double EA_EPS = DBL_EPSILON / 2; /* horner.h:21 */

static inline cplx c_make(double re, double im) { cplx r = {re, im}; return r; }
static inline cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
static inline cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
static inline cplx c_mul(cplx a, cplx b) {
  return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
static inline cplx c_scale(double s, cplx a) { return c_make(s * a.re, s * a.im); }
static inline double c_abs(cplx a) { return hypot(a.re, a.im); }
/* thrust scaled division, arithmetic.h:119-142 */
static inline cplx c_div(cplx x, cplx y) {
  double s = fabs(y.re) + fabs(y.im);
  double oos = 1.0 / s;
  double ars = x.re * oos, ais = x.im * oos, brs = y.re * oos, bis = y.im * oos;
  s = (brs * brs) + (bis * bis);
  oos = 1.0 / s;
  return c_make(((ars * brs) + (ais * bis)) * oos, ((ais * brs) - (ars * bis)) * oos);
}
static inline cplx c_recip(cplx y) { return c_div(c_make(1.0, 0.0), y); }
/* thrust::pow(z, 2) == exp(log(z) * 2), cpow.h:36-44 */
static inline cplx c_sq_pow(cplx z) {
  double complex l = clog(z.re + z.im * I);
  double complex e = cexp(l * 2.0);
  return c_make(creal(e), cimag(e));
}

/* ---- error-free transformations (horner.h:44-84) ---------------------------------------- */
static inline void two_sum(double a, double b, double *s, double *e) {
  double x = a + b; /* Knuth, 6 flops, horner.h:44-49 */
  double t = x - a;
  *e = (a - (x - t)) + (b - t);
  *s = x;
}
static inline void two_prod(double a, double b, double *p, double *e) {
  double x = a * b; /* fma-based, horner.h:51-56 */
  *e = fma(a, b, -x);
  *p = x;
}
/* sort 4 values by decreasing magnitude (selection sort, horner.h:164-188: first maximum wins) */
static void sort4_desc_abs(double *p) {
  for (int i = 0; i < 3; ++i) {
    double mx = fabs(p[i]);
    int ind = i;
    for (int j = i + 1; j < 4; ++j) {
      double t = fabs(p[j]);
      if (t > mx) { mx = t; ind = j; }
    }
    if (ind != i) { double t = p[i]; p[i] = p[ind]; p[ind] = t; }
  }
}
/* Priest doubly-compensated summation of 4 terms, horner.h:190-208 */
static double priest_sum4(double *p) {
  sort4_desc_abs(p);
  double s = p[0], c = 0;
  for (int i = 1; i < 4; ++i) {
    double y = c + p[i];
    double u = p[i] - (y - c);
    double t = y + s;
    double v = y - (t - s);
    double z = u + v;
    s = t + z;
    c = z - (s - t);
  }
  return s;
}
/* complex product with its three error terms + complex sum error (horner.h:58-84) reduced to the
 * one thing the callers use: result and priest-summed total error, plus the |.|-sum for eb. */
typedef struct { cplx res; cplx e1, e2, e3; } cprod_eft;
static inline cprod_eft two_prod_cplx(cplx a, cplx b) {
  double p0, e0, p1, e1, p2, e2, p3, e3, s4, e4, s5, e5;
  two_prod(a.re, b.re, &p0, &e0);
  two_prod(a.im, b.im, &p1, &e1);
  two_prod(a.re, b.im, &p2, &e2);
  two_prod(a.im, b.re, &p3, &e3);
  two_sum(p0, -p1, &s4, &e4);
  two_sum(p2, p3, &s5, &e5);
  cprod_eft r;
  r.res = c_make(s4, s5);
  r.e1 = c_make(e0, e2);
  r.e2 = c_make(-e1, e3);
  r.e3 = c_make(e4, e5);
  return r;
}
/* one compensated multiply-add  out = acc*x + addend, returning the 4 error terms */
static inline cplx comp_step(cplx acc, cplx x, cplx addend, cplx err[4]) {
  cprod_eft pr = two_prod_cplx(acc, x);
  double sr, er, si, ei;
  two_sum(pr.res.re, addend.re, &sr, &er);
  two_sum(pr.res.im, addend.im, &si, &ei);
  err[0] = pr.e1; err[1] = pr.e2; err[2] = pr.e3; err[3] = c_make(er, ei);
  return c_make(sr, si);
}
static inline cplx priest_cplx(const cplx err[4]) {
  double r[4] = {err[0].re, err[1].re, err[2].re, err[3].re};
  double i[4] = {err[0].im, err[1].im, err[2].im, err[3].im};
  return c_make(priest_sum4(r), priest_sum4(i));
}

/* gamma_const, horner.h:30-35 */
static double gamma_const(unsigned n) {
  double s = 1.41421356237309504880;
  double g = (2 * n * EA_EPS) * s;
  return g / ((1 - DBL_EPSILON) - g);
}

/* Real Horner of alpha at x; reversed=1 walks the coefficients from index 0 (horner.h:219-239) */
static double horner_real(const double *a, double x, int deg, int reversed) {
  double h;
  if (reversed) {
    h = a[0];
    for (int i = 1; i <= deg; ++i) h = fma(h, x, a[i]);
  } else {
    h = a[deg];
    for (int i = deg - 1; i >= 0; --i) h = fma(h, x, a[i]);
  }
  return h;
}
/* complex Horner, value and derivative (horner.h:241-267) */
static void horner_cplx(const cplx *p, cplx x, int deg, int reversed, cplx *h, cplx *hd) {
  cplx v = reversed ? p[0] : p[deg];
  cplx d = c_make(0, 0);
  for (int k = 1; k <= deg; ++k) {
    cplx c = reversed ? p[k] : p[deg - k];
    d = c_add(c_mul(d, x), v);
    v = c_add(c_mul(v, x), c);
  }
  *h = v; *hd = d;
}
/* compensated complex Horner with running error bound (horner.h:281-385) */
static void horner_comp_cplx(const cplx *p, cplx x, int deg, int reversed, cplx *h, cplx *hd,
                             double *eb) {
  cplx v = reversed ? p[0] : p[deg];
  cplx d = c_make(0, 0), e = c_make(0, 0), ed = c_make(0, 0);
  double b = 0;
  cplx err[4];
  for (int k = 1; k <= deg; ++k) {
    cplx c = reversed ? p[k] : p[deg - k];
    /* derivative recurrence: hd = hd*x + h, ed = ed*x + e + sum(err) */
    d = comp_step(d, x, v, err);
    ed = c_add(c_add(c_mul(ed, x), e), priest_cplx(err));
    /* value recurrence: h = h*x + c, e = e*x + sum(err) */
    v = comp_step(v, x, c, err);
    e = c_add(c_mul(e, x), priest_cplx(err));
    double ap[4] = {c_abs(err[0]), c_abs(err[1]), c_abs(err[2]), c_abs(err[3])};
    b = b * c_abs(x) + priest_sum4(ap);
  }
  *h = c_add(v, e);
  *hd = c_add(d, ed);
  *eb = b;
}

/* Aberth sum over the other roots, ehrlich_aberth.h:31-40 */
static cplx aberth_sum(const cplx *z, int deg, int j) {
  cplx s = c_make(0, 0);
  for (int i = 0; i < deg; ++i)
    if (i != j) s = c_add(s, c_recip(c_sub(z[j], z[i])));
  return s;
}
/* correction (ehrlich_aberth.h:27-41) and reversed correction (:43-58) */
static cplx correction(const cplx *z, cplx h, cplx hd, int deg, int j) {
  cplx s = aberth_sum(z, deg, j);
  return c_div(h, c_sub(hd, c_mul(h, s)));
}
static cplx rcorrection(const cplx *z, cplx h, cplx hd, int deg, int j) {
  cplx s = aberth_sum(z, deg, j);
  cplx z2h = c_mul(c_sq_pow(z[j]), h);
  /* the reference evaluates pow(z,2)*h*corr left to right, i.e. (pow*h)*corr */
  cplx den = c_sub(c_sub(c_mul(c_scale((double)deg, z[j]), h), hd), c_mul(z2h, s));
  return c_div(z2h, den);
}

/* Bini initial estimates from the upper convex hull of (i, log alpha_i), init_est.h:30-102 */
typedef struct { int x; double y; } hpt;
static double ccw(const hpt *a, const hpt *b, const hpt *c) {
  return (b->x - a->x) * (c->y - a->y) - (b->y - a->y) * (c->x - a->x);
}
static void init_est(const double *alpha, int deg, cplx *roots) {
  hpt pts[64], hull[64];
  const double pi2 = 6.28318530717958647693, sigma = 0.7;
  for (int i = 0; i <= deg; ++i) {
    pts[i].x = i;
    pts[i].y = alpha[i] > 0 ? log(alpha[i]) : -1E+30;
  }
  int k = 0;
  for (int i = deg; i >= 0; --i) {
    while (k >= 2 && ccw(&hull[k - 2], &hull[k - 1], &pts[i]) <= 0) --k;
    hull[k++] = pts[i];
  }
  int hs = k, pos = 0;
  double th = pi2 / deg;
  for (int i = hs - 2; i >= 0; --i) {
    int nz = hull[i].x - hull[i + 1].x;
    double a1 = pow(alpha[hull[i + 1].x], 1.0 / nz);
    double a2 = pow(alpha[hull[i].x], 1.0 / nz);
    double r = a1 / a2, ang = pi2 / nz;
    for (int j = 0; j < nz; ++j) /* comma-operator quirk of init_est.h:95: real part only */
      roots[pos + j] = c_make(r * sin(ang * j + th * i + sigma), 0.0);
    pos += nz;
  }
}

/* One polynomial.  Returns the number of sweeps used; *nupd / *ncupd count plain and compensated
 * root updates (for the flop accounting in DESIGN.md).  ehrlich_aberth.h:60-150 and :153-292 */
static int ea_solve_one(const cplx *poly, const cplx *init, cplx *roots, int deg, int itmax,
                        int compensated, int custom_init, long *nupd, long *ncupd, int *converged) {
  double alpha[64];
  unsigned char c1[64], c2[64];
  for (int i = 0; i <= deg; ++i) alpha[i] = c_abs(poly[i]);
  for (int i = 0; i < deg; ++i) { c1[i] = 0; c2[i] = 0; }
  if (!custom_init) init_est(alpha, deg, roots);
  else for (int i = 0; i < deg; ++i) roots[i] = init[i];
  for (int i = 0; i <= deg; ++i) alpha[i] = alpha[i] * fma(3.8284271247461900976, i, 1);
  const double g = gamma_const(4 * deg + 2);
  int it, done = 0;
  for (it = 0; it < itmax; ++it) {
    for (int j = 0; j < deg; ++j) {
      if (!c1[j]) {
        double az = c_abs(roots[j]);
        int rev = az > 1;
        cplx x = rev ? c_recip(roots[j]) : roots[j];
        double b = horner_real(alpha, rev ? 1. / az : az, deg, rev);
        cplx h, hd;
        horner_cplx(poly, x, deg, rev, &h, &hd);
        if (c_abs(h) > EA_EPS * b) {
          cplx corr = rev ? rcorrection(roots, h, hd, deg, j) : correction(roots, h, hd, deg, j);
          roots[j] = c_sub(roots[j], corr);
          if (nupd) ++*nupd;
        } else c1[j] = 1;
      } else if (compensated && !c2[j]) {
        double az = c_abs(roots[j]);
        int rev = az > 1;
        cplx x = rev ? c_recip(roots[j]) : roots[j];
        cplx h, hd; double b;
        horner_comp_cplx(poly, x, deg, rev, &h, &hd, &b);
        double ah = c_abs(h);
        double bound = EA_EPS * ah + (g * b + 2 * pow(EA_EPS, 2) * ah);
        if (ah > 4 * bound) {
          cplx corr = rev ? rcorrection(roots, h, hd, deg, j) : correction(roots, h, hd, deg, j);
          /* relative test on the reversed branch, ABSOLUTE on the standard one (:238 vs :260) */
          double thr = rev ? 4 * EA_EPS * az : 4 * EA_EPS;
          if (c_abs(corr) > thr) { roots[j] = c_sub(roots[j], corr); if (ncupd) ++*ncupd; }
          else c2[j] = 1;
        } else c2[j] = 1;
      }
    }
    done = 1;
    for (int j = 0; j < deg; ++j) done = done && (compensated ? c2[j] : c1[j]);
    if (done) { ++it; break; }
  }
  if (converged) *converged = done;
  return it;
}

/* Batched entry: coeffs (size, deg+1) complex128 low->high, roots_init (size, deg) or NULL,
 * roots (size, deg).  sweeps (optional, size ints) receives the sweep count per polynomial;
 * stats (optional, 3 longs): plain updates, compensated updates, #polys not converged. */
int ea_oracle_solve(const double *coeffs, const double *roots_init, double *roots, int64_t size,
                    int deg, int itmax, int compensated, int custom_init, int *sweeps, long *stats) {
  if (deg < 1 || deg > 62) return 1;
  long nupd = 0, ncupd = 0, nfail = 0;
  for (int64_t n = 0; n < size; ++n) {
    int conv = 0;
    int it = ea_solve_one((const cplx *)coeffs + n * (deg + 1),
                          custom_init ? (const cplx *)roots_init + n * deg : NULL,
                          (cplx *)roots + n * deg, deg, itmax, compensated, custom_init, &nupd,
                          &ncupd, &conv);
    if (sweeps) sweeps[n] = it;
    if (!conv) ++nfail;
  }
  if (stats) { stats[0] = nupd; stats[1] = ncupd; stats[2] = nfail; }
  return 0;
}
