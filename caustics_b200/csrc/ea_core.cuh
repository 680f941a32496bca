// Ehrlich-Aberth polynomial root solver, one thread per polynomial (device code for sm_100a).
//
// Behaviour follows the reference solver (/root/reference/lib/ehrlich_aberth/ehrlich_aberth.h:60-292,
// horner.h, init_est.h): the same Gauss-Seidel sweep over the roots in index order, the same
// reversed-polynomial evaluation for |z| > 1, the same stopping tests and -- when `compensated` --
// the same second polishing phase built on error-free transformations with a running error bound.
// What is B200-specific is the execution shape:
//   * coefficients live in registers (Horner loops fully unrolled, static register indices); the
//     roots and the |coefficients| live in shared memory in [index][thread] planes so that the rolled
//     loop over roots can index them dynamically without bank conflicts and the kernel fits 5 CTAs
//     of 128 threads per SM at 96 registers;
//   * the standard and the reversed evaluation are ONE instruction stream: the branch of the
//     reference (|z| > 1) becomes per-lane selects of the evaluation point and the coefficient
//     order (select-free variants when a warp vote finds the lanes agree), so a warp never executes
//     both paths;
//   * a cold-start root step is one straight-line basic block (prologue, Horner, Aberth sum,
//     correction, predicated store); warm starts use a branchy step that skips the Aberth sum when an
//     evaluation only confirms convergence;
//   * convergence is a per-lane bit mask; a (sweep, root) step is skipped when a warp vote says no
//     lane needs it, and the sweep loop ends on a warp vote (the reference's per-polynomial loop
//     exit, lifted to the warp).  Each lane still stops updating a root exactly when the reference
//     would, so in the default mode the iteration path of every polynomial is the reference's
//     (same sweep counts, same root order); results do not depend on the batch layout;
//   * coefficients are normalised by a power of two on load (exact, roots unchanged) so that |.|^2
//     comparisons can replace hypot() everywhere; reciprocals / square roots are branch-free
//     (hardware seed + Newton).  Supported range: the moduli of the coefficients, of the roots and of p'(root)
//     may span ~1e+-150 around the largest coefficient (squares must stay inside the double range; the
//     stopping test itself falls back to unsquared moduli, ea_exceeds).  The reference's hypot() and scaled
//     complex division go further (z^10 - 1e-200); lens polynomials span < 1e10;
//   * compensated kernels run the plain sweeps first and the polishing sweeps afterwards, with the
//     per-step error terms summed plainly (see priest_sum4) -- same polished roots as the reference.
// The FP64 pipe is the bound; a DFMA with three distinct register operands issues at 69 % of the
// constant-operand rate on B200 and the kernel sits at that ceiling.
#pragma once
#include "cplx.cuh"

namespace cb200 {

constexpr double EA_EPS = 1.1102230246251565e-16;  // 2^-53, horner.h:21

// gamma_const(n), horner.h:30-35
__host__ __device__ constexpr double ea_gamma(int n) {
  return ((2.0 * n * 1.1102230246251565e-16) * 1.41421356237309504880) /
         ((1.0 - 2.220446049250313e-16) - (2.0 * n * 1.1102230246251565e-16) * 1.41421356237309504880);
}

enum : int { EA_INIT_REFERENCE = 0, EA_INIT_BINI = 1 };

// Reciprocal used inside the Aberth sum  S = sum_i 1/(z_j - z_i).  S only steers the iteration: the
// fixed point of z <- z - h/(hd - h S) is h(z) = 0 whatever S is, and the stopping test looks at
// h alone, so S needs far less than double accuracy: one Newton step on the 20-bit hardware seed
// (~1e-12 relative) reproduces the reference's path (same sweep counts, same root order).  With the
// seed alone (~1e-6 relative per term, `fast`) the converged roots are still identical to 2e-15 and
// the kernel is 7 % faster, but ~5 % of random polynomials take one sweep more or fewer and a root
// pair occasionally swaps labels -- so `fast` is used only where the caller has already given up the
// reference's root order (CAUSTICS_FLAG_INIT_BINI, and the order-independent magnification sums).
#ifndef CB200_STRAIGHT_LINE
#define CB200_STRAIGHT_LINE 1
#endif
#ifndef CB200_ABERTH_MODE
#define CB200_ABERTH_MODE 1
#endif
template <bool FAST>
__device__ __forceinline__ double rcp_aberth(double x) {
#if CB200_ABERTH_MODE >= 1 && !defined(CB200_HOSTSIM)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  if (!FAST) {
    const double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
  }
  return y;
#else
  return rcp_fast(x);
#endif
}
constexpr unsigned EA_COMP_CAP = 12;  // see ea_solve_thread
#if defined(CB200_HOSTSIM) && defined(CB200_HOSTSIM_COUNT)
// work counters of the host-compiled test build (tests/hostsim): polynomial evaluations and root updates
static long long g_ea_evals = 0, g_ea_updates = 0;
#define CB200_COUNT_EVAL(n) (g_ea_evals += (n))
#define CB200_COUNT_UPD(n) (g_ea_updates += (n))
#else
#define CB200_COUNT_EVAL(n) ((void)0)
#define CB200_COUNT_UPD(n) ((void)0)
#endif

// Power-of-two normalisation of the coefficients: p_i *= 2^-e with e = exponent of max |component|.
template <int DEG>
__device__ __forceinline__ void ea_normalise(cd (&p)[DEG + 1]) {
  double m = 0.0;
#pragma unroll
  for (int i = 0; i <= DEG; ++i) m = fmax(m, fmax(fabs(p[i].re), fabs(p[i].im)));
  int hi = __double2hiint(m);
  int ex = (hi >> 20) & 0x7ff;
  if (ex != 0 && ex != 0x7ff) {
    // 2^(1023-ex); for ex = 2046 (max |coefficient| >= 2^1023) that would be the exponent field 0, i.e. +0.0:
    // scale by 2^-1022 instead (the largest coefficient ends up in [2, 4) -- still exact)
    if (ex > 2045) ex = 2045;
    double sc = __hiloint2double((2046 - ex) << 20, 0);
#pragma unroll
    for (int i = 0; i <= DEG; ++i) { p[i].re *= sc; p[i].im *= sc; }
  }
}

// log for the initial estimates (one per coefficient).  libm's log() is ~175 instructions with its
// special-case handling and constant loads -- 5 % of the whole degree-10 kernel -- so the default is the
// classic argument reduction x = 2^k (1 + f), sqrt(1/2) < 1 + f < sqrt(2), s = f / (2 + f),
// log(1 + f) = f - (f^2/2 - s (f^2/2 + R(s^2))) with the degree-7 minimax R of the public fdlibm e_log.c
// (error < 1 ulp: it differs from libm's result in the last bit for ~3 % of arguments, which is also how far
// the reference's host libm and CUDA's libm are apart).  Arguments are the positive, finite |coefficients|;
// a denormal is treated as ~2^-1023, good enough for a starting radius.  CB200_LIBM_LOG=1: libm.
#ifndef CB200_LIBM_LOG
#define CB200_LIBM_LOG 0
#endif
__device__ __forceinline__ double ea_log(double a) {
#if CB200_LIBM_LOG
  return log(a);
#else
  const int hi = __double2hiint(a), lo = __double2loint(a);
  double k = (double)(((hi >> 20) & 0x7ff) - 1023);
  double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);   // mantissa in [1, 2)
  if (m > 1.4142135623730951) { m *= 0.5; k += 1.0; }
  const double f = m - 1.0;
  const double s = f * rcp_fast(2.0 + f);
  const double z = s * s, w = z * z;
  const double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
  const double t2 = z * (6.666666666666735130e-01 + w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 + w * 1.479819860511658591e-01)));
  const double hfsq = 0.5 * f * f;
  return k * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + (t2 + t1)) + k * 1.90821492927058770002e-10)) - f);
#endif
}

// cos / sin of 2 pi / n, n = 0..16 (n = 0 unused)
__device__ const double EA_ROT_COS[17] = {1.0, 1.0, -1.0, -0.4999999999999998, 6.123233995736766e-17, 0.30901699437494745, 0.5000000000000001, 0.6234898018587336, 0.7071067811865476, 0.766044443118978, 0.8090169943749475, 0.8412535328311812, 0.8660254037844387, 0.8854560256532099, 0.9009688679024191, 0.9135454576426009, 0.9238795325112867};
__device__ const double EA_ROT_SIN[17] = {0.0, -2.4492935982947064e-16, 1.2246467991473532e-16, 0.8660254037844387, 1.0, 0.9510565162951535, 0.8660254037844386, 0.7818314824680298, 0.7071067811865475, 0.6427876096865393, 0.5877852522924731, 0.5406408174555976, 0.49999999999999994, 0.4647231720437685, 0.4338837391175581, 0.40673664307580015, 0.3826834323650898};

// cos / sin of 2 pi i / DEG + 0.7 for DEG = 2..16, i = 0..DEG-1, at offset DEG (DEG - 1) / 2 - 1: the starting angle of
// hull edge i (init_est.h:93-96), for the order-independent (EA_INIT_BINI) estimates
__device__ const double EA_START_COS[135] = {0.7648421872844884, -0.7648421872844884, 0.7648421872844884, -0.9403299763573428, 0.1754877890728544, 0.7648421872844884, -0.644217687237691, -0.7648421872844884, 0.644217687237691, 0.7648421872844884, -0.376338195474186, -0.9974319833523373, -0.24010867170377767, 0.8490366632458125, 0.7648421872844884, -0.1754877890728544, -0.9403299763573428, -0.7648421872844884, 0.1754877890728544, 0.9403299763573428, 0.7648421872844884, -0.02679836564196351, -0.7982592026529799, -0.9686145785460706, -0.4095834206573606, 0.45787240696551046, 0.9805409732483756, 0.7648421872844884, 0.08529440196032743, -0.644217687237691, -0.9963557923724988, -0.7648421872844884, -0.08529440196032743, 0.644217687237691, 0.9963557923724988, 0.7648421872844884, 0.171807960134941, -0.5016171209945315, -0.9403299763573428, -0.9390519851789534, -0.4983811337350216, 0.1754877890728544, 0.7672440250440125, 0.9999982547295531, 0.7648421872844884, 0.24010867170377767, -0.376338195474186, -0.8490366632458125, -0.9974319833523373, -0.7648421872844884, -0.24010867170377767, 0.376338195474186, 0.8490366632458125, 0.9974319833523373, 0.7648421872844884, 0.295135815063864, -0.26827409310951694, -0.7465088722547888, -0.987732359038807, -0.9153578008113573, -0.552363608435463, -0.013997873196067564, 0.5288120878788335, 0.903727947459871, 0.9917165691589436, 0.7648421872844884, 0.3402639204555768, -0.1754877890728544, -0.644217687237691, -0.9403299763573428, -0.9844816076932679, -0.7648421872844884, -0.3402639204555768, 0.1754877890728544, 0.644217687237691, 0.9403299763573428, 0.9844816076932679, 0.7648421872844884, 0.377851236305031, -0.09570087931087926, -0.5473290767972883, -0.8735707788198555, -0.9996879430839287, -0.8967886471332002, -0.5884458795990359, -0.14529725259033896, 0.3311372239650917, 0.7317121531462729, 0.9646606461290116, 0.9766170105046313, 0.7648421872844884, 0.4095834206573606, -0.02679836564196351, -0.45787240696551046, -0.7982592026529799, -0.9805409732483756, -0.9686145785460706, -0.7648421872844884, -0.4095834206573606, 0.02679836564196351, 0.45787240696551046, 0.7982592026529799, 0.9805409732483756, 0.9686145785460706, 0.7648421872844884, 0.4366911664900616, 0.033032275794800676, -0.376338195474186, -0.7206363738205124, -0.9403299763573428, -0.9974319833523373, -0.8820689390406132, -0.6141881618240236, -0.24010867170377767, 0.1754877890728544, 0.5607408168622756, 0.8490366632458125, 0.9905263572982096, 0.9607450455242901, 0.7648421872844884, 0.4600906066908837, 0.08529440196032743, -0.3024871022730095, -0.644217687237691, -0.8878719691683112, -0.9963557923724988, -0.9531534781757227, -0.7648421872844884, -0.4600906066908837, -0.08529440196032743, 0.3024871022730095, 0.644217687237691, 0.8878719691683112, 0.9963557923724988, 0.9531534781757227};
__device__ const double EA_START_SIN[135] = {0.644217687237691, -0.644217687237691, 0.644217687237691, 0.3402639204555768, -0.9844816076932679, 0.644217687237691, 0.7648421872844884, -0.644217687237691, -0.7648421872844884, 0.644217687237691, 0.9264823595877222, -0.07162009903527673, -0.9707460150691568, -0.5283339327209797, 0.644217687237691, 0.9844816076932679, 0.3402639204555768, -0.644217687237691, -0.9844816076932679, -0.3402639204555768, 0.644217687237691, 0.9996408593084416, 0.6023140753625378, -0.24856749229941172, -0.9122726684071027, -0.8890179182331535, -0.19631454296900253, 0.644217687237691, 0.9963557923724988, 0.7648421872844884, 0.08529440196032743, -0.644217687237691, -0.9963557923724988, -0.7648421872844884, -0.08529440196032743, 0.644217687237691, 0.9851304608194138, 0.865089743278209, 0.3402639204555768, -0.34377517236046384, -0.8669580413935812, -0.9844816076932679, -0.64135528845895, 0.0018682981153722482, 0.644217687237691, 0.9707460150691568, 0.9264823595877222, 0.5283339327209797, -0.07162009903527673, -0.644217687237691, -0.9707460150691568, -0.9264823595877222, -0.5283339327209797, 0.07162009903527673, 0.644217687237691, 0.9554553106590536, 0.9633426238707941, 0.6653754606572769, 0.15615629000342357, -0.40264139937889587, -0.833603289386597, -0.9999020249734405, -0.8487389326013205, -0.4281072260310128, 0.12844549994302745, 0.644217687237691, 0.9403299763573428, 0.9844816076932679, 0.7648421872844884, 0.3402639204555768, -0.1754877890728544, -0.644217687237691, -0.9403299763573428, -0.9844816076932679, -0.7648421872844884, -0.3402639204555768, 0.1754877890728544, 0.644217687237691, 0.92586632038473, 0.9954101374303581, 0.8369174879832707, 0.4866971279883117, 0.02498032130745356, -0.4424591759394357, -0.8085366081897077, -0.9893880474261346, -0.9435826084157617, -0.6816137652204842, -0.26349542274953835, 0.21498654560924696, 0.644217687237691, 0.9122726684071027, 0.9996408593084416, 0.8890179182331535, 0.6023140753625378, 0.19631454296900253, -0.24856749229941172, -0.644217687237691, -0.9122726684071027, -0.9996408593084416, -0.8890179182331535, -0.6023140753625378, -0.19631454296900253, 0.24856749229941172, 0.644217687237691, 0.8996114856478598, 0.9994542854757371, 0.9264823595877222, 0.6933132168989876, 0.3402639204555768, -0.07162009903527673, -0.4711203527547574, -0.7891596174889001, -0.9707460150691568, -0.9844816076932679, -0.827991386612583, -0.5283339327209797, -0.13732274209882214, 0.2774327981701691, 0.644217687237691, 0.8878719691683112, 0.9963557923724988, 0.9531534781757227, 0.7648421872844884, 0.4600906066908837, 0.08529440196032743, -0.3024871022730095, -0.644217687237691, -0.8878719691683112, -0.9963557923724988, -0.9531534781757227, -0.7648421872844884, -0.4600906066908837, -0.08529440196032743, 0.3024871022730095};
// exp for a STARTING RADIUS (EA_INIT_BINI only; ~1e-7 relative): 2^k by the exponent field, 2^f by the
// single-precision hardware exponential
__device__ __forceinline__ double ea_exp_start(double x) {
  double t = x * 1.4426950408889634;
  t = fmin(fmax(t, -1000.0), 1000.0);
  const double k = rint(t);
  const float e = exp2f((float)(t - k));
  return __hiloint2double(((int)k + 1023) << 20, 0) * (double)e;
}

// Bini initial estimates from the upper convex hull of (i, log|p_i|) -- init_est.h:57-102.
// mode EA_INIT_REFERENCE reproduces the reference's purely real guesses r*sin(.) (the comma
// expression at init_est.h:95), so sweep counts and root order match the reference; EA_INIT_BINI
// uses the intended r*(cos, sin) which converges in 10-30 % fewer updates.
template <int DEG, int NT, class ALPHA>
__device__ __noinline__ void ea_init_est(const ALPHA& al, double* zre, double* zim, int mode) {
  double ly[DEG + 1];
  int hx[DEG + 1];
#pragma unroll
  for (int i = 0; i <= DEG; ++i) { const double a = al.get(i); ly[i] = a > 0 ? ea_log(a) : -1e30; }
  int k = 0;
  for (int i = DEG; i >= 0; --i) {
    while (k >= 2) {
      int x1 = hx[k - 2], x2 = hx[k - 1];
      double ccw = (double)(x2 - x1) * (ly[i] - ly[x1]) - (ly[x2] - ly[x1]) * (double)(i - x1);
      if (ccw <= 0) --k; else break;
    }
    hx[k++] = i;
  }
  int pos = 0;
  for (int i = k - 2; i >= 0; --i) {
    int lo = hx[i + 1], up = hx[i];
    int nz = up - lo;
    // (|p_lo| / |p_up|)^(1/nz), from the logs already at hand
    // angles 2 pi j / nz + 2 pi i / DEG + 0.7: one sincospi per hull edge, then exact-table rotations
    // by 2 pi / nz (instead of one libm sincos per root).  The order-independent estimates take radius and
    // starting angle from a cheap exponential and a table: they only steer the iteration, and libm's exp and
    // sincospi were a third of the instruction-fetch stalls of the cold point-source kernel (profiles/)
    double r, s, c;
    if (mode == EA_INIT_REFERENCE) {
      r = exp((ly[lo] - ly[up]) / nz);
      sincospi(fma((double)i, 2.0 / DEG, 0.22281692032865347), &s, &c);   // 0.7 / pi
    } else {
      r = ea_exp_start((ly[lo] - ly[up]) / nz);
      c = EA_START_COS[DEG * (DEG - 1) / 2 - 1 + i];
      s = EA_START_SIN[DEG * (DEG - 1) / 2 - 1 + i];
    }
    const double rc = EA_ROT_COS[nz], rs = EA_ROT_SIN[nz];
    for (int j = 0; j < nz; ++j) {
      zre[(pos + j) * NT] = (mode == EA_INIT_REFERENCE) ? r * s : r * c;
      zim[(pos + j) * NT] = (mode == EA_INIT_REFERENCE) ? 0.0 : r * s;
      const double cn = fma(c, rc, -s * rs), sn = fma(s, rc, c * rs);
      c = cn; s = sn;
    }
    pos += nz;
  }
}

// sort four values by decreasing magnitude with the tie behaviour of the reference's selection
// sort (horner.h:164-188: the first maximum wins), written with static indices only.
__device__ __forceinline__ void sort4_desc_abs(double (&p)[4]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double mx = fabs(p[i]);
    int ind = i;
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      double t = fabs(p[j]);
      if (t > mx) { mx = t; ind = j; }
    }
    double head = p[i], best = p[i];
#pragma unroll
    for (int j = i + 1; j < 4; ++j)
      if (ind == j) { best = p[j]; p[j] = head; }
    p[i] = best;
  }
}
// Priest's doubly compensated summation of 4 terms, horner.h:190-208.
// CB200_PRIEST = 0 replaces it by the plain sum ((p0 + p1) + (p2 + p3)): the four terms are the
// rounding errors of ONE Horner step, and what consumes their sum is `e = e*x + sum` in ordinary
// double arithmetic, which commits a relative 2^-53 error of its own -- summing the terms more
// accurately than that cannot change h + e.  The sorted, doubly compensated form costs ~35 FP64
// instructions against 3 and is more than half of the reference's compensated Horner step.
#ifndef CB200_PRIEST
#define CB200_PRIEST 0
#endif
__device__ __forceinline__ double priest_sum4(double (&p)[4]) {
#if !CB200_PRIEST
  return __dadd_rn(__dadd_rn(p[0], p[1]), __dadd_rn(p[2], p[3]));
#endif
  sort4_desc_abs(p);
  double s = p[0], c = 0.0;
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    double y = __dadd_rn(c, p[i]);
    double u = __dsub_rn(p[i], __dsub_rn(y, c));
    double t = __dadd_rn(y, s);
    double v = __dsub_rn(y, __dsub_rn(t, s));
    double z = __dadd_rn(u, v);
    s = __dadd_rn(t, z);
    c = __dsub_rn(z, __dsub_rn(s, t));
  }
  return s;
}

// One compensated Horner step  acc <- acc*x + add  (horner.h:58-84,296-328): returns the rounded
// result, the Priest-summed complex error `err` and leaves the four error terms' moduli in `ab`.
__device__ __forceinline__ cd comp_step(cd acc, cd x, cd add, cd& err, double (&ab)[4]) {
  double p0, e0, p1, e1, p2, e2, p3, e3, sr, er, si, ei, vr, fr, vi, fi;
  two_prod(acc.re, x.re, p0, e0);
  two_prod(acc.im, x.im, p1, e1);
  two_prod(acc.re, x.im, p2, e2);
  two_prod(acc.im, x.re, p3, e3);
  two_sum(p0, -p1, sr, er);
  two_sum(p2, p3, si, ei);
  two_sum(sr, add.re, vr, fr);
  two_sum(si, add.im, vi, fi);
  // error terms: (e0, e2), (-e1, e3), (er, ei), (fr, fi)
  double re4[4] = {e0, -e1, er, fr};
  double im4[4] = {e2, e3, ei, fi};
  ab[0] = sqrt_fast(e0 * e0 + e2 * e2);
  ab[1] = sqrt_fast(e1 * e1 + e3 * e3);
  ab[2] = sqrt_fast(er * er + ei * ei);
  ab[3] = sqrt_fast(fr * fr + fi * fi);
  err = mk(priest_sum4(re4), priest_sum4(im4));
  return mk(vr, vi);
}

// |coefficient| holders: registers (static indices) or a shared-memory plane per thread.  The shared
// variant frees 2*(DEG+1) registers per thread, which buys another resident CTA per SM.
#ifndef CB200_ALPHA_SMEM
#define CB200_ALPHA_SMEM 1
#endif
template <int DEG, int NT>
struct AlphaRegs {
  double v[DEG + 1];
  __device__ __forceinline__ double get(int k) const { return v[k]; }
  __device__ __forceinline__ double get2(int k, bool rev) const { return rev ? v[k] : v[DEG - k]; }
  __device__ __forceinline__ void set(int k, double x) { v[k] = x; }
};
template <int DEG, int NT>
struct AlphaSmem {
  double* base;
  __device__ __forceinline__ double get(int k) const { return base[k * NT]; }
  __device__ __forceinline__ double get2(int k, bool rev) const { return base[(rev ? k : DEG - k) * NT]; }
  __device__ __forceinline__ void set(int k, double x) { base[k * NT] = x; }
};

template <int DEG, int NT, bool SMEM> struct AlphaPick { typedef AlphaRegs<DEG, NT> type; };
template <int DEG, int NT> struct AlphaPick<DEG, NT, true> { typedef AlphaSmem<DEG, NT> type; };
template <int DEG, int NT> __device__ __forceinline__ void alpha_bind(AlphaRegs<DEG, NT>&, double*) {}
template <int DEG, int NT> __device__ __forceinline__ void alpha_bind(AlphaSmem<DEG, NT>& a, double* base) { a.base = base; }

// The stopping test |h| > thr (thr = EPS * b, ehrlich_aberth.h:109/:122) on squares, nh = |h|^2 -- except when
// thr^2 would underflow (a polynomial whose coefficients span more than ~150 decades, e.g. z^10 - 1e-200:
// |h| keeps shrinking far below sqrt(DBL_MIN) while the reference's hypot-based test is still unmet); there
// the moduli are compared unsquared through a scaled hypot.  Rare path, inside the lazily evaluated block.
#ifndef CB200_EXCEEDS_INLINE
#define CB200_EXCEEDS_INLINE 0
#endif
#if CB200_EXCEEDS_INLINE
__device__ __forceinline__
#else
static __device__ __noinline__     // out of line: its division and square root cost the headline 3 % when they sit in the step
#endif
bool ea_exceeds_unsquared(double hre, double him, double thr) {
  const double ar = fabs(hre), ai = fabs(him);
  const double m = fmax(ar, ai), q = fmin(ar, ai);
  if (!(m > 0.0)) return false;
  const double r = q / m;
  return m * sqrt(fma(r, r, 1.0)) > thr;
}
__device__ __forceinline__ bool ea_exceeds(cd h, double nh, double thr) {
  if (thr > 1e-150) return nh > __dmul_rn(thr, thr);
  return ea_exceeds_unsquared(h.re, h.im, thr);
}

// Value, derivative and the real bound polynomial in one unrolled Horner pass (horner.h:219-267).
// MODE 0: every lane evaluates p at x (coefficients high->low); MODE 1: every lane evaluates the
// reversed polynomial (coefficients from index 0, reference's rhorner_*); MODE 2: per-lane select.
// MODE 0/1 are taken when a warp vote finds the lanes agree, which saves the 6 selects per step.
template <int DEG, int MODE, class ALPHA, bool WITH_B = true>
__device__ __forceinline__ void horner_plain(const cd (&p)[DEG + 1], const ALPHA& al, cd x,
                                             double ax, bool rev, cd& h, cd& hd, double& b) {
#define CB200_COEF(k) (MODE == 0 ? p[DEG - (k)] : MODE == 1 ? p[k] : csel(rev, p[k], p[DEG - (k)]))
#define CB200_ALPH(k) (MODE == 0 ? al.get(DEG - (k)) : MODE == 1 ? al.get(k) : al.get2(k, rev))
  h = CB200_COEF(0);
  if (WITH_B) b = CB200_ALPH(0);
  hd = h;
  h = cfma(h, x, CB200_COEF(1));
  if (WITH_B) b = fma(b, ax, CB200_ALPH(1));
#pragma unroll
  for (int k = 2; k <= DEG; ++k) {
    hd = cfma(hd, x, h);
    h = cfma(h, x, CB200_COEF(k));
    if (WITH_B) b = fma(b, ax, CB200_ALPH(k));
  }
}
// the real bound polynomial alone (same operations, same order as inside horner_plain: same bits)
template <int DEG, int MODE, class ALPHA>
__device__ __forceinline__ double horner_bound(const ALPHA& al, double ax, bool rev) {
  double b = CB200_ALPH(0);
#pragma unroll
  for (int k = 1; k <= DEG; ++k) b = fma(b, ax, CB200_ALPH(k));
  return b;
#undef CB200_COEF
#undef CB200_ALPH
}

// One plain (non-compensated) root step as ONE straight-line block: prologue, Horner pass, Aberth sum,
// correction and a predicated store, with no data-dependent branch in between.  The Aberth sum does
// not depend on the Horner result, so inside one basic block the scheduler overlaps the
// latency-bound Horner chains with the throughput-bound reciprocals (in the branchy formulation the
// convergence test separated them).  Lanes that do not need the root (already converged) execute the
// arithmetic on their converged value and commit nothing; so does the one evaluation per root that
// only confirms convergence.  MODE as in horner_plain.
// CB200_LAZY_BOUND: the stopping test |h| <= EPS*b needs the bound polynomial b = sum alpha_i |x|^i (and, in
// the standard evaluation, |z| itself) only when it can pass.  Every evaluation point has |x| <= 1 (z for
// |z| <= 1, 1/z otherwise), so b <= A = sum alpha_i; an evaluation with |h| > EPS*A -- all but the last one
// or two of a root -- is "not converged" whatever b is.  The step therefore evaluates b (and |z|) only when
// a warp vote finds a lane inside that margin: same decisions, same roots, bit for bit, ~6 % fewer FP64
// instructions at degree 10.  thrA2 = (EPS * A * (1 + 1e-9))^2.
#ifndef CB200_LAZY_BOUND
#define CB200_LAZY_BOUND 1
#endif
// PASSZ: the caller hands over the root and |z|^2 it already loaded for the variant vote (plain kernels:
// 2 shared loads and 2 FP64 instructions fewer per step, headline 1.648 -> 1.609 ms); the compensated kernels
// reload them here instead, because keeping them live across the vote costs their plain stage registers
// (measured: 4.13 -> 4.49 ms with PASSZ).
template <int DEG, int MODE, int NT, class ALPHA, bool FAST, bool PASSZ>
__device__ __forceinline__ void ea_step_plain(const cd (&p)[DEG + 1], const ALPHA& al,
                                              double* zre, double* zim, int j, bool need, unsigned& c1,
                                              double thrA2, cd z, double az2) {
  if (!PASSZ) {
    z = mk(zre[j * NT], zim[j * NT]);
    az2 = norm2(z);
  }
  const bool rev = MODE == 2 ? az2 > 1.0 : MODE == 1;
  double rs = 0.0, absz = 0.0;
  if (!CB200_LAZY_BOUND || MODE != 0) {
    rs = rsqrt_fast(az2);
    rs = az2 > 0.0 ? rs : 0.0;
    absz = az2 * rs;
  }
  cd x = z;
  double ax = absz;
  if (MODE != 0) {
    const double inv = rs * rs;
    const cd xr = mk(z.re * inv, -z.im * inv);
    x = MODE == 1 ? xr : csel(rev, xr, z);
    ax = MODE == 1 ? rs : (rev ? rs : absz);
  }
  cd h, hd;
  double b = 0.0;
  horner_plain<DEG, MODE, ALPHA, !CB200_LAZY_BOUND>(p, al, x, ax, rev, h, hd, b);
  cd s = mk(0, 0);
#pragma unroll
  for (int i = 0; i < DEG - 1; ++i) {
    const int ii = i + (i >= j ? 1 : 0);  // skip root j without a branch
    const cd a = z - mk(zre[ii * NT], zim[ii * NT]);
    const double inv = rcp_aberth<FAST>(norm2(a));
    s = mk(fma(a.re, inv, s.re), fma(-a.im, inv, s.im));
  }
  cd num = h, den = hd;
  if (MODE != 0) {
    const cd z2 = z * z;
    const cd numr = z2 * h;
    const cd denr = cfma((double)DEG * z, h, -hd);
    num = MODE == 1 ? numr : csel(rev, numr, h);
    den = MODE == 1 ? denr : csel(rev, denr, hd);
  }
  den = cfma(-num, s, den);
  const cd corr = cdiv(num, den);
  const double nh = norm2(h);
  bool big;
  if (CB200_LAZY_BOUND) {
    big = true;
    if (__any_sync(0xffffffffu, need && !(nh > thrA2))) {
      if (MODE == 0) {
        double r0 = rsqrt_fast(az2);
        r0 = az2 > 0.0 ? r0 : 0.0;
        ax = az2 * r0;
      }
      b = horner_bound<DEG, MODE, ALPHA>(al, ax, rev);
      big = ea_exceeds(h, nh, EA_EPS * b);
    }
  } else {
    big = ea_exceeds(h, nh, EA_EPS * b);  // |h| > EPS*b, ehrlich_aberth.h:109/:122
  }
  if (need && big) {
    zre[j * NT] = z.re - corr.re;
    zim[j * NT] = z.im - corr.im;
  }
  if (need && !big) c1 |= (1u << j);
  CB200_COUNT_EVAL(need ? 1 : 0);
  CB200_COUNT_UPD(need && big ? 1 : 0);
}

// Shared-memory planes owned by one CTA of NT threads.
template <int DEG, bool COMP, int NT>
struct EASmem {
  double zre[DEG][NT];
  double zim[DEG][NT];
  // (plain kernels only: the compensated ones already use 43 KB for the coefficient planes)
  double al[(CB200_ALPHA_SMEM && !COMP) ? DEG + 1 : 1][(CB200_ALPHA_SMEM && !COMP) ? NT : 1];
  // compensated phase only: coefficient planes for the rolled compensated Horner loop
  double cre[COMP ? DEG + 1 : 1][COMP ? NT : 1];
  double cim[COMP ? DEG + 1 : 1][COMP ? NT : 1];
};

struct EAResult {
  int sweeps;      // sweeps this polynomial used (== itmax when not converged)
  bool converged;  // all roots satisfied the stopping test
};

// Solve one polynomial per thread.  p: coefficients low->high (already normalised) in registers.
// Roots are read from / left in sm.zre/zim[.][tid] (the caller stores custom initial roots there
// when custom_init).  All 32 lanes of a warp must call this together; `active` = false lanes idle.
// STRAIGHT selects the straight-line plain step (best for cold starts, where ~90 % of the evaluations
// are followed by an update) or the branchy one that skips the Aberth sum when no lane updates (best
// for warm starts, where every root's second evaluation only confirms convergence).
template <int DEG, bool COMP, int NT, bool STRAIGHT = (CB200_STRAIGHT_LINE != 0), bool PASSZ = !COMP>
__device__ __forceinline__ EAResult ea_solve_thread(const cd (&p)[DEG + 1], EASmem<DEG, COMP, NT>& sm,
                                                    int tid, bool active, bool custom_init,
                                                    int init_mode, int itmax, bool fast = false) {
  constexpr unsigned FULL = (1u << DEG) - 1u;
  double* zre = &sm.zre[0][tid];
  double* zim = &sm.zim[0][tid];

  typedef typename AlphaPick<DEG, NT, (CB200_ALPHA_SMEM && !COMP)>::type ALPHA;
  ALPHA al;
  alpha_bind(al, &sm.al[0][tid]);
#pragma unroll
  for (int i = 0; i <= DEG; ++i) al.set(i, fast ? sqrt_fast(norm2(p[i])) : cabs_fast(p[i]));  // ehrlich_aberth.h:77-82 (fast: branch-free square root, <= 1 ulp)
  if (active && !custom_init) {
    // A polynomial whose coefficients are all exactly real keeps the reference's real-axis guesses
    // on the real axis in exact arithmetic; the reference only escapes through rounding noise in
    // thrust::pow (exp(2 log z) of a negative real), taking ~30 sweeps.  Use the complex Bini
    // guesses for those polynomials instead (e.g. every binary-lens source on the lens axis).
    double imsum = 0.0;
#pragma unroll
    for (int i = 0; i <= DEG; ++i) imsum += fabs(p[i].im);
    ea_init_est<DEG, NT, ALPHA>(al, zre, zim, imsum == 0.0 ? (int)EA_INIT_BINI : init_mode);
  }
#pragma unroll
  for (int i = 0; i <= DEG; ++i) al.set(i, al.get(i) * fma(3.8284271247461900976, (double)i, 1.0));  // :96-99
  double thrA2 = 0.0;   // (EPS * sum alpha_i)^2 with a 1e-9 margin: see CB200_LAZY_BOUND
  if (CB200_LAZY_BOUND) {
    double asum = 0.0;
#pragma unroll
    for (int i = 0; i <= DEG; ++i) asum += al.get(i);
    asum *= EA_EPS * (1.0 + 1e-9);
    thrA2 = asum * asum;
  }
  if (COMP) {
#pragma unroll
    for (int i = 0; i <= DEG; ++i) { sm.cre[i][tid] = p[i].re; sm.cim[i][tid] = p[i].im; }
  }

  unsigned c1 = active ? 0u : FULL;  // plain-phase convergence bits
  unsigned c2 = c1;                  // compensated-phase bits
  // Compensated updates applied per root, 4 bits each.  The reference's polishing phase can
  // limit-cycle between neighbouring doubles (|corr| stays just above its 4*EPS exit test) until
  // itmax = 2500 sweeps (~1e-4 of lens polynomials); a lane doing that would pin
  // its whole warp for ~10^3 x the normal run time.  A root that has taken EA_COMP_CAP polishing
  // updates (legitimate roots need 1-3) is inside such a cycle -- every iterate of the cycle is
  // within a few ulp of the others -- and is declared converged.
  unsigned long long ncomp = 0ull;
  EAResult res;
  res.sweeps = 0;
  res.converged = !active;
  // Compensated kernels run in two stages: plain sweeps until every lane of the warp has all roots
  // through the plain stopping test, then the polishing sweeps.  (The reference lets a root start
  // polishing while others are still in the plain phase; polishing a root depends on the others only
  // through the Aberth sum, which steers but does not set the fixed point, so the polished roots are
  // the same -- golden vectors: identical to 5e-15 -- while the warp never executes the plain and
  // the compensated evaluation for the same root, and the plain stage runs the straight-line step.)
  bool stage2 = false, spent = false;
  int s1 = 0, it2 = 0;   // own plain sweeps; polishing sweeps since the warp switched
  int it = 0;
  // Compensated kernels: a lane's budget is itmax of its OWN sweeps (plain + polishing); the sweeps it spends
  // waiting for the other lanes of its warp to leave the plain stage are not charged, so whether a polynomial
  // converges within itmax does not depend on its warp neighbours (the warp loop may run up to 2 itmax).
  const int it_end = COMP ? (itmax > 0x3fffffff ? 0x7fffffff : 2 * itmax) : itmax;
  for (; it < it_end; ++it) {
    if (COMP && !stage2 && __all_sync(0xffffffffu, c1 == FULL)) stage2 = true;
    const unsigned done_bits = COMP ? c2 : c1;
    if (__all_sync(0xffffffffu, done_bits == FULL)) break;
    if (COMP && stage2) ++it2;
#pragma unroll 1
    for (int j = 0; j < DEG; ++j) {
      const bool need1 = !((c1 >> j) & 1u);
      const bool need2 = COMP && stage2 && !need1 && !((c2 >> j) & 1u);
      if (!__any_sync(0xffffffffu, need1 || need2)) continue;
      if (STRAIGHT && (!COMP || !stage2)) {
        // warp-uniform choice of the evaluation variant, then one straight-line step
        cd zj = mk(0, 0);
        double az2j = 0.0;
        bool rv;
        if (PASSZ) {
          zj = mk(zre[j * NT], zim[j * NT]);
          az2j = norm2(zj);
          rv = az2j > 1.0;
        } else {
          rv = zre[j * NT] * zre[j * NT] + zim[j * NT] * zim[j * NT] > 1.0;
        }
        const bool all_std = __all_sync(0xffffffffu, !need1 || !rv);
        const bool all_rev = __all_sync(0xffffffffu, !need1 || rv);
        if (fast) {
          if (all_std) ea_step_plain<DEG, 0, NT, ALPHA, true, PASSZ>(p, al, zre, zim, j, need1, c1, thrA2, zj, az2j);
          else if (all_rev) ea_step_plain<DEG, 1, NT, ALPHA, true, PASSZ>(p, al, zre, zim, j, need1, c1, thrA2, zj, az2j);
          else ea_step_plain<DEG, 2, NT, ALPHA, true, PASSZ>(p, al, zre, zim, j, need1, c1, thrA2, zj, az2j);
        } else {
          if (all_std) ea_step_plain<DEG, 0, NT, ALPHA, false, PASSZ>(p, al, zre, zim, j, need1, c1, thrA2, zj, az2j);
          else if (all_rev) ea_step_plain<DEG, 1, NT, ALPHA, false, PASSZ>(p, al, zre, zim, j, need1, c1, thrA2, zj, az2j);
          else ea_step_plain<DEG, 2, NT, ALPHA, false, PASSZ>(p, al, zre, zim, j, need1, c1, thrA2, zj, az2j);
        }
        continue;
      }

      const cd z = mk(zre[j * NT], zim[j * NT]);
      const double az2 = norm2(z);
      const bool rev = az2 > 1.0;  // |z| > 1, ehrlich_aberth.h:106
      // |z| and (reversed lanes) 1/z, 1/|z| from one reciprocal square root; standard lanes need |z| only
      // for the bound polynomial, which is evaluated lazily (CB200_LAZY_BOUND)
      double rs = 0.0, absz = 0.0;
      if (!CB200_LAZY_BOUND || rev || (COMP && need2)) {
        rs = az2 > 0.0 ? rsqrt_fast(az2) : 0.0;
        absz = az2 * rs;
      }
      cd x = z;
      double ax = absz;
      if (rev) {
        const double inv = rs * rs;
        x = mk(z.re * inv, -z.im * inv);
        ax = rs;
      }
      cd h, hd;
      bool upd = false;
      // warp votes are taken by all lanes, outside the per-lane branch
      const bool all_std = __all_sync(0xffffffffu, !need1 || !rev);
      const bool all_rev = __all_sync(0xffffffffu, !need1 || rev);
      CB200_COUNT_EVAL(need1 ? 1 : 0);
      if (need1) {
        double b = 0.0;
        if (all_std) horner_plain<DEG, 0, ALPHA, !CB200_LAZY_BOUND>(p, al, x, ax, rev, h, hd, b);
        else if (all_rev) horner_plain<DEG, 1, ALPHA, !CB200_LAZY_BOUND>(p, al, x, ax, rev, h, hd, b);
        else horner_plain<DEG, 2, ALPHA, !CB200_LAZY_BOUND>(p, al, x, ax, rev, h, hd, b);
        const double nh = norm2(h);
        if (CB200_LAZY_BOUND && nh > thrA2) {
          upd = true;                                   // |h| > EPS * sum alpha_i >= EPS*b
        } else {
          if (CB200_LAZY_BOUND) {
            if (!rev && !(COMP && need2)) {
              const double r0 = az2 > 0.0 ? rsqrt_fast(az2) : 0.0;
              ax = az2 * r0;
            }
            b = horner_bound<DEG, 2, ALPHA>(al, ax, rev);
          }
          if (ea_exceeds(h, nh, EA_EPS * b)) upd = true;  // |h| > EPS*b, :109/:122
          else c1 |= (1u << j);
        }
      }
      if (COMP) {
        if (need2) {
          // compensated Horner with running error bound, horner.h:281-385 (rolled; coefficients
          // come from the shared planes so the loop index can be dynamic)
          cd e = mk(0, 0), ed = mk(0, 0), err;
          double eb = 0.0, ab[4], abd[4];
          h = mk(sm.cre[rev ? 0 : DEG][tid], sm.cim[rev ? 0 : DEG][tid]);
          hd = mk(0, 0);
#pragma unroll 1
          for (int k = 1; k <= DEG; ++k) {
            const int idx = rev ? k : DEG - k;
            const cd c = mk(sm.cre[idx][tid], sm.cim[idx][tid]);
            hd = comp_step(hd, x, h, err, abd);
            ed = (ed * x + e) + err;
            h = comp_step(h, x, c, err, ab);
            e = e * x + err;
            eb = eb * ax + priest_sum4(ab);
          }
          h = h + e;
          hd = hd + ed;
          const double ah = cabs_fast(h);
          const double bound = EA_EPS * ah + (ea_gamma(4 * DEG + 2) * eb + 2.0 * EA_EPS * EA_EPS * ah);
          if (ah > 4.0 * bound) upd = true;  // :235/:257
          else c2 |= (1u << j);
        }
      }
      CB200_COUNT_UPD(upd ? 1 : 0);
      if (__any_sync(0xffffffffu, upd)) {
        if (upd) {
          // Aberth sum over the other roots (:31-40) and the (reversed) correction (:41,:56-57)
          cd s = mk(0, 0);
#pragma unroll
          for (int i = 0; i < DEG - 1; ++i) {
            const int ii = i + (i >= j ? 1 : 0);  // skip root j without a branch
            const cd a = z - mk(zre[ii * NT], zim[ii * NT]);
            const double inv = rcp_aberth<false>(norm2(a));
            s = mk(fma(a.re, inv, s.re), fma(-a.im, inv, s.im));
          }
          cd num = h, den = hd;
          if (rev) {
            const cd z2 = z * z;
            num = z2 * h;
            den = ((double)DEG * z) * h - hd;
          }
          den = den - num * s;
          const cd corr = cdiv(num, den);
          bool apply = true;
          if (COMP) {
            if (need2) {
              // relative test on the reversed branch, absolute on the standard one (:238 vs :260)
              const double t = rev ? 4.0 * EA_EPS * absz : 4.0 * EA_EPS;
              if (!(norm2(corr) > t * t)) { apply = false; c2 |= (1u << j); }
            }
          }
          if (COMP) {
            if (need2 && apply) {
              const unsigned cnt = (unsigned)(ncomp >> (4 * j)) & 15u;
              if (cnt + 1u >= EA_COMP_CAP) c2 |= (1u << j);
              ncomp += 1ull << (4 * j);
            }
          }
          if (apply) {
            zre[j * NT] = z.re - corr.re;
            zim[j * NT] = z.im - corr.im;
          }
        }
      }
    }
    if (active && !res.converged && !spent) {
      if (!COMP) {
        res.sweeps = it + 1;
        if (c1 == FULL) res.converged = true;
      } else {
        // own plain sweeps + polishing sweeps: independent of the other lanes of the warp
        if (s1 == 0 && c1 == FULL) s1 = it + 1;
        res.sweeps = (s1 ? s1 : it + 1) + it2;
        if (c2 == FULL) res.converged = true;
        else if (res.sweeps >= itmax) { spent = true; c1 = FULL; c2 = FULL; }   // budget spent: stop, reported unconverged
      }
    }
  }
  return res;
}

// ---------------------------------------------------------------------------------------------
// Lane-per-root solve for batches too small to fill the machine with one thread per polynomial.
// A warp holds G = 32 / DEG polynomials; lane g*DEG + r owns root r of polynomial g.  One sweep =
//   (A) every lane evaluates the polynomial (value, derivative, bound) at its own root -- the
//       expensive part, in parallel -- and applies the reference's stopping test;
//   (B) the roots that have to move are corrected ONE AFTER THE OTHER in root order: the group's
//       lanes each form their term 1/(z_j - z_i) of the Aberth sum with the roots as they are now
//       (roots < j already moved in this sweep), a shuffle tree adds them, lane j applies the correction.
// That is the reference's Gauss-Seidel sweep (ehrlich_aberth.h:100-143) -- root j is evaluated at the
// value it had when the sweep started, its Aberth sum sees the already-updated roots -- with the
// per-polynomial dependent chain cut from DEG evaluations to one; only the order in which the DEG - 1
// terms of a sum are added differs from ea_solve_thread (rounding level: which root a lane converges to,
// hence the labelling of the image tracks, is the thread solver's).  Plain mode only.
// p: normalised coefficients low->high (identical in all lanes of a group); z: the lane's root (in:
// warm start, out: converged root); valid: lane owns a root.  Returns the sweeps used.
#ifndef CB200_HOSTSIM
template <int DEG>
__device__ __forceinline__ int ea_solve_group(const cd (&p)[DEG + 1], cd& z, int lane, bool valid, int itmax) {
  constexpr unsigned FULLW = 0xffffffffu;
  const int base = (lane / DEG) * DEG, r = lane - base;
  AlphaRegs<DEG, 1> al;
#pragma unroll
  for (int i = 0; i <= DEG; ++i) al.v[i] = cabs_fast(p[i]) * fma(3.8284271247461900976, (double)i, 1.0);
  bool conv = !valid;
  int it = 0;
  for (; it < itmax; ++it) {
    if (__all_sync(FULLW, conv)) break;
    // (A) evaluation at the own root
    const double az2 = norm2(z);
    const bool rev = az2 > 1.0;
    const double rs = az2 > 0.0 ? rsqrt_fast(az2) : 0.0;
    cd x = z;
    double ax = az2 * rs;
    if (rev) {
      const double inv = rs * rs;
      x = mk(z.re * inv, -z.im * inv);
      ax = rs;
    }
    cd h, hd;
    double b;
    horner_plain<DEG, 2, AlphaRegs<DEG, 1> >(p, al, x, ax, rev, h, hd, b);
    const bool upd = !conv && ea_exceeds(h, norm2(h), EA_EPS * b);
    conv = conv || !upd;
    cd num = h, den = hd;
    if (rev) {
      const cd z2 = z * z;
      num = z2 * h;
      den = ((double)DEG * z) * h - hd;
    }
    // (B) corrections in root order
    const unsigned moving = __ballot_sync(FULLW, upd);
#pragma unroll 1
    for (int j = 0; j < DEG; ++j) {
      // does root j move in any of the warp's polynomials?  (lanes j, DEG + j, 2 DEG + j, ...)
      unsigned sel = 0;
#pragma unroll
      for (int g = 0; g < 32 / DEG; ++g) sel |= 1u << (g * DEG + j);
      if (!(moving & sel)) continue;
      const cd zj = mk(__shfl_sync(FULLW, z.re, base + j), __shfl_sync(FULLW, z.im, base + j));
      const cd a = zj - z;
      const double inv = (r != j && base + DEG <= 32) ? rcp_aberth<false>(norm2(a)) : 0.0;
      double tr = a.re * inv, ti = -a.im * inv;
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
        const double ur = __shfl_down_sync(FULLW, tr, off), ui = __shfl_down_sync(FULLW, ti, off);
        if (r + off < DEG) { tr += ur; ti += ui; }
      }
      const cd s = mk(__shfl_sync(FULLW, tr, base), __shfl_sync(FULLW, ti, base));
      if (upd && r == j) z = z - cdiv(num, den - num * s);
    }
  }
  return it;
}
#endif

}  // namespace cb200
