/*
 * caustics_b200 -- C ABI of the B200 (sm_100a) implementation of caustics' hot path.
 *
 * Plain pointers and sizes only.  Every launcher BORROWS its buffers, enqueues on the caller's
 * stream and returns without synchronising; none allocates device memory (the *_host entry points,
 * which own an internal pinned/device workspace, are the exception and say so).  Return value:
 * 0 on success, CAUSTICS_ERR_* for bad arguments, 1000 + cudaError_t for CUDA failures.  Nothing
 * throws across this boundary and there is no CPU fallback: without a CUDA device every compute
 * entry point returns 1000 + cudaErrorNoDevice (or the launch error).
 *
 * Reference interfaces replaced (paths under /root/reference):
 *   caustics_ea_xla            lib/ehrlich_aberth/kernels.h:19-20  gpu_ehrlich_aberth(stream, buffers,
 *                              opaque, opaque_len), registered by gpu_ops.cc:14-19 and bound in
 *                              src/caustics/ehrlich_aberth_primitive.py:22-28,223-242
 *   caustics_ea_descriptor     lib/ehrlich_aberth/kernels.h:11-17 EhrlichAberthDescriptor +
 *                              gpu_ops.cc:24-25 build_ehrlich_aberth_descriptor
 *   caustics_ea_solve[_host]   lib/ehrlich_aberth/cpu_ops.cc:15-81 cpu_ehrlich_aberth (same operands:
 *                              size, deg, itmax, compensated, custom_init, coeffs, roots_init -> roots)
 *   caustics_images_point_source  src/caustics/point_source.py:1655-1709 _images_point_source
 *   caustics_mag_point_source     src/caustics/point_source.py:1762-1830 mag_point_source
 *   caustics_mag_extended_source  src/caustics/extended_source.py:741-904 mag_extended_source
 *   caustics_mag                  src/caustics/lightcurve.py:99-254 mag
 *   caustics_trajectory           src/caustics/trajectory.py:106-158 AnnualParallaxTrajectory.compute
 *   caustics_marginalized_log_likelihood  src/caustics/linalg.py:55-70 (diagonal covariance)
 *
 * complex128 arrays are passed as `const void*` to interleaved (re, im) doubles, C order.
 */
#ifndef CAUSTICS_B200_H_
#define CAUSTICS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAUSTICS_OK 0
#define CAUSTICS_ERR_BAD_ARG 1
#define CAUSTICS_ERR_UNSUPPORTED_DEGREE 2
#define CAUSTICS_ERR_BAD_DESCRIPTOR 3
#define CAUSTICS_ERR_CUDA_BASE 1000

/* `flags` bits.  Default (0): initial estimates exactly as the reference computes them, including
 * the real-axis quirk of init_est.h:95 (same sweep counts and root ORDER as the reference);
 * coefficients stored low->high as the reference primitive expects. */
#define CAUSTICS_FLAG_INIT_BINI 1         /* intended complex Bini estimates: 10-30 % fewer updates, other root order */
#define CAUSTICS_FLAG_COEFFS_HIGH_FIRST 2 /* coeffs rows are high->low (saves poly_roots' flip, primitive.py:73) */
#define CAUSTICS_FLAG_GRID_WALK 4         /* caustics_mag_point_source_grid only: warm-started walks along y (below) */
#define CAUSTICS_FLAG_PATH_WALK 8         /* caustics_mag_point_source only: consecutive elements of w are neighbours
                                             (a trajectory) -> warm-started walks along the array (below) */

/* Low-level lens parameters, exactly the reference's `_params` dict plus the centre-of-mass shift
 * its public functions add to the source positions (point_source.py:1796-1819).
 * nlenses = 1: nothing else used.  2: a, e1.  3: a, e1, e2, r3. */
typedef struct caustics_lens {
  int32_t nlenses;
  int32_t reserved;
  double a;
  double e1;
  double e2;
  double r3_re, r3_im;
  double x_cm;
} caustics_lens;

/* Opaque descriptor for the XLA custom call (the reference's descriptor, kernels.h:11-17, extended
 * by custom_init which the reference's GPU path silently dropped, and by flags). */
typedef struct caustics_ea_descriptor {
  int64_t size;
  int32_t deg;
  int32_t itmax;
  uint8_t compensated;
  uint8_t custom_init;
  uint8_t flags;
  uint8_t reserved;
  int32_t pad;
} caustics_ea_descriptor;

/* Opaque descriptors of the two magnification custom calls (no reference counterpart: the
 * reference lowers them to thousands of XLA ops around "gpu_ehrlich_aberth"; same legacy signature). */
typedef struct caustics_mag_ps_descriptor {
  int64_t n;
  caustics_lens lens;
  int32_t itmax;
  uint8_t compensated;
  uint8_t flags;
  uint8_t reserved[2];
} caustics_mag_ps_descriptor;

typedef struct caustics_mag_ext_descriptor {
  int64_t n;
  caustics_lens lens;
  double rho;
  double u1;
  double q;                 /* user's mass ratio, gate only */
  uint64_t workspace_bytes; /* size of buffers[2] */
  int32_t npts_limb;
  int32_t npts_ld;
  int32_t itmax;
  uint8_t limb_darkening;
  uint8_t compensated;
  uint8_t gate;             /* 1: lightcurve.py mag (hexadecapole where valid), 0: mag_extended_source */
  uint8_t reserved;
} caustics_mag_ext_descriptor;

/* library / device queries (no compute) */
const char* caustics_version(void);
int caustics_device_count(void);
/* degrees for which a kernel is instantiated (2..16): 1 if supported */
int caustics_ea_degree_supported(int deg);
const char* caustics_error_string(int code);

/* ---- kernel 1: Ehrlich-Aberth roots --------------------------------------------------------
 * coeffs (size, deg+1) complex128 LOW->HIGH, roots_init (size, deg) complex128 (may be NULL when
 * custom_init == 0), roots (size, deg) complex128, sweeps (size) int32 or NULL: sweeps used per
 * polynomial, negated when the polynomial did not converge within itmax.  Device pointers. */
int caustics_ea_solve(const void* coeffs, const void* roots_init, void* roots, int32_t* sweeps,
                      int64_t size, int deg, int itmax, int compensated, int custom_init,
                      int flags, void* stream);

/* Implicit-function tangent / cotangent of the roots (ehrlich_aberth_primitive.py:254-324), device
 * pointers, coeffs LOW->HIGH (size, deg+1), roots (size, deg), deg <= 32:
 *   jvp: dcoeffs (size, deg+1) -> droots (size, deg),  dz_j = -(sum_k dp_k z_j^k) / p'(z_j)
 *   vjp: groots (size, deg) -> gcoeffs (size, deg+1),  gp_k = sum_j conj(-z_j^k / p'(z_j)) gz_j */
int caustics_ea_jvp(const void* coeffs, const void* roots, const void* dcoeffs, void* droots, int64_t size,
                    int deg, void* stream);
int caustics_ea_vjp(const void* coeffs, const void* roots, const void* groots, void* gcoeffs, int64_t size,
                    int deg, void* stream);

/* Same contract with HOST pointers: chunked H2D -> kernel -> D2H pipeline over an internal pinned +
 * device workspace on the current device (grows on demand, freed by caustics_release_workspace).
 * Synchronous: returns when `roots` is complete. */
int caustics_ea_solve_host(const void* coeffs, const void* roots_init, void* roots, int32_t* sweeps,
                           int64_t size, int deg, int itmax, int compensated, int custom_init,
                           int flags);
void caustics_release_workspace(void);

/* XLA GPU custom call (legacy api_version 0 signature, kernels.h:19-20).
 * buffers = [coeffs, roots_init, roots]; opaque = caustics_ea_descriptor bytes.  On a bad descriptor
 * nothing is launched and the sticky error is readable with caustics_last_xla_error(). */
void caustics_ea_xla(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* returns the first error recorded since the last call and clears it (process-wide, atomic: XLA runs
 * custom calls on its own thread) */
int caustics_last_xla_error(void);
/* buffers = [w (n) complex128, mag (n) float64]; opaque = caustics_mag_ps_descriptor bytes */
void caustics_mag_ps_xla(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers = [w (n) complex128, mag (n) float64, workspace (descriptor.workspace_bytes, an extra
 * uint8 result XLA allocates as scratch)]; opaque = caustics_mag_ext_descriptor bytes */
void caustics_mag_ext_xla(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* fills *out; returns sizeof(caustics_ea_descriptor) */
size_t caustics_ea_make_descriptor(caustics_ea_descriptor* out, int64_t size, int deg, int itmax,
                                   int compensated, int custom_init, int flags);

/* _images_point_source_sequential (point_source.py:1711-1759): npaths independent 1-D paths of n
 * source positions each; position 0 of a path is solved from the default initial estimates, position
 * k > 0 is warm-started from the images of position k-1, so row j follows one image along the path.
 * w (npaths, n) complex128 -> z (npaths, deg, n) complex128, mask (npaths, deg, n) uint8. */
int caustics_images_point_source_sequential(const void* w, void* z, uint8_t* mask, int64_t npaths, int64_t n,
                                            const caustics_lens* lens, int itmax, int compensated, void* stream);

/* ---- kernel 2: fused point-source images / magnification -----------------------------------
 * w (n) complex128 source positions (the x_cm shift in `lens` is added inside).
 * images: z (deg, n) complex128 with the root axis FIRST like the reference, mask (deg, n) uint8,
 *         z_init (n, deg) complex128 root axis LAST like the reference's z_init (custom_init).
 * mag:    mag (n) float64 = sum over real images of 1/|det J|.  nimages (n) uint8 or NULL.
 *         flags & CAUSTICS_FLAG_PATH_WALK (with nimages == NULL): the caller states that w is a path, i.e.
 *         w[i+1] is close to w[i]; each thread then owns a run of consecutive elements and starts every
 *         solve after its first from the extrapolated roots of the elements before (the reference's
 *         custom_init warm start, point_source.py:1711-1759).  Same stopping test, other iteration path:
 *         agrees with the default to rounding x conditioning; an array that is not a path is still solved
 *         correctly (a failed warm solve is redone cold), only slower.  Batches too small to fill the GPU
 *         with runs of >= 2 elements take the default kernel. */
int caustics_images_point_source(const void* w, const void* z_init, void* z, uint8_t* mask,
                                 int64_t n, const caustics_lens* lens, int itmax, int compensated,
                                 int custom_init, int flags, void* stream);
int caustics_mag_point_source(const void* w, double* mag, uint8_t* nimages, int64_t n,
                              const caustics_lens* lens, int itmax, int compensated, int flags,
                              void* stream);
/* magnification map: w generated on the fly, w[iy*nx + ix] = (x0 + ix*dx) + i (y0 + iy*dy),
 * rows [row_begin, row_end) written to mag[(iy-row_begin)*nx + ix].
 * flags & CAUSTICS_FLAG_GRID_WALK: each thread walks a column segment of up to 32 rows, starting every
 * solve after the first from the extrapolated roots of the rows before -- the reference's custom_init
 * warm start (point_source.py:1711-1759) applied to a regular map; ~3x fewer root updates.  Every
 * root still passes the solver's own stopping test, but the iteration path differs from a cold solve,
 * so a pixel agrees with the default entry to rounding x conditioning (<= 1e-10 relative away from
 * caustics) rather than bit for bit, and depends on where row_begin cuts the map. */
int caustics_mag_point_source_grid(double x0, double y0, double dx, double dy, int64_t nx,
                                   int64_t row_begin, int64_t row_end, double* mag,
                                   const caustics_lens* lens, int itmax, int compensated,
                                   int flags, void* stream);
int caustics_mag_point_source_host(const void* w, double* mag, int64_t n, const caustics_lens* lens,
                                   int itmax, int compensated, int flags);
/* the map entry with a HOST result buffer (rows [row_begin, row_end) -> mag[(iy-row_begin)*nx + ix]):
 * row blocks (multiples of 32 rows, so walked maps are cut on walk boundaries) flow through the
 * internal workspace as kernel -> D2H, the copy of block k overlapping the kernel of block k+1.
 * Synchronous.  `mag` may be pinned, pageable, or a shared-memory segment several ranks write
 * disjoint row blocks of (how a one-process-per-GPU driver assembles the map on the host). */
int caustics_mag_point_source_grid_host(double x0, double y0, double dx, double dy, int64_t nx,
                                        int64_t row_begin, int64_t row_end, double* mag,
                                        const caustics_lens* lens, int itmax, int compensated, int flags);

/* ---- kernel family 3: extended-source magnification by contour integration ------------------
 * w (n) complex128 source-disk centres, rho the source radius, npts_limb / limb_darkening / u1 /
 * npts_ld as in the reference (extended_source.py:741-753).  The phases keep their per-source state
 * in `workspace` (device memory, caustics_ext_workspace_bytes(n, ...) bytes, borrowed for the call).
 * caustics_mag is the light-curve entry (lightcurve.py:99-254): binary lenses use the hexadecapole
 * approximation wherever the reference's validity tests pass (q is the user's mass ratio, tested
 * against 0.01 for the planetary-caustic test) and full integration elsewhere; used_hexadecapole
 * (n) uint8, optional, records the decision.  Limits: npts_limb <= 1280, n * (D*npts_limb) < 2^31.
 * limb_darkening: 0 uniform disk, 1 linear limb darkening with the reference's quadrature
 * (integrate.py:47-121: npts_ld/2 + npts_ld/2 Gauss-Legendre nodes per P and Q integral of every contour
 * vertex), 1 | CAUSTICS_LD_ADAPTIVE the same with the half-order rule (npts_ld/4 nodes) on far panels that
 * are at most 8 rho long and stay 2 rho away from the vertex (the panel that ends on the limb
 * keeps its nodes).  Opt-in: it moves results by up to 7e-5, so the default stays the
 * reference's rule.
 * Streams: every launch is ordered on `stream`.  Un-gated uniform-disk calls (caustics_mag_extended_source,
 * caustics_mag_extended_source_grad) of at least 4 096 sources run as two windows, the second on a library-owned
 * side stream that forks from and joins `stream` by events inside the call: to the caller the call is still one
 * stream-ordered operation (and one capturable sub-graph); results are bit for bit those of a single window. */
#define CAUSTICS_LD_ADAPTIVE 2
/* caustics_ext_workspace_bytes only, OR-ed into limb_darkening = 0: the workspace will serve
 * caustics_mag_extended_source alone (not the tangent / contour-export entry points, which keep the image
 * tracks as arrays): about half the bytes per source. */
#define CAUSTICS_WS_MAG_ONLY 4
size_t caustics_ext_workspace_bytes(int64_t n, int nlenses, int npts_limb, int limb_darkening, int npts_ld);
/* workspace of a GATED call (caustics_mag, caustics_mag_extended_source_list) over n points that
 * integrates at most max_full sources at a time: per-source arrays for max_full sources + the compact
 * list of n points.  caustics_mag accepts any workspace >= caustics_mag_workspace_bytes(n, 1, ...): the
 * points that fail the gate are integrated in windows of as many sources as the workspace holds, the
 * window loop is enqueued without reading the device-side count back (a window past the count exits at
 * once).  A light curve sends a few per cent of its points to the integration, so a workspace for
 * n / 8 sources is ample and 8x smaller than caustics_ext_workspace_bytes(n, ...). */
size_t caustics_mag_workspace_bytes(int64_t n, int64_t max_full, int nlenses, int npts_limb, int limb_darkening,
                                    int npts_ld);
int caustics_mag_extended_source(const void* w, double* mag, int64_t n, double rho, const caustics_lens* lens,
                                 int npts_limb, int limb_darkening, double u1, int npts_ld, int itmax,
                                 int compensated, void* workspace, size_t workspace_bytes, void* stream);
int caustics_mag(const void* w, double* mag, uint8_t* used_hexadecapole, int64_t n, double rho,
                 const caustics_lens* lens, double q, int npts_limb, int limb_darkening, double u1, int npts_ld,
                 int itmax, int compensated, void* workspace, size_t workspace_bytes, void* stream);

/* Uniform-disk magnification AND its tangent in one pass: grad (8, n) float64 =
 * d mag / d(a, e1, e2, Re r3, Im r3, Re w, Im w, rho) per source, low-level lens parameters as in
 * caustics_lens (w is the source centre AFTER the x_cm shift; entries of parameters the lens does not have
 * are 0).  The rule is the reference's implicit-function JVP (ehrlich_aberth_primitive.py:290-324) applied
 * on the lens equation at every contour vertex and pushed through the trapezoid sum while the contour is
 * walked; limb sampling, masks, gate decisions and contour topology are constants, which is what jax.grad
 * differentiates in the reference (tests/test_extended_source.py:294-331).  The chain to the high-level
 * (s, q, q3, r3, psi) parameters stays in Python (caustics_b200/extended_source.py). */
int caustics_mag_extended_source_grad(const void* w, double* mag, double* grad, int64_t n, double rho,
                                      const caustics_lens* lens, int npts_limb, int itmax, int compensated,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* Second half of the two-call form  caustics_mag_gate -> (host reads *count, sizes the workspace) ->
 * here: full contour integration of the points w[list[k]], k < *count (device int32, <= max_count),
 * results to mag[list[k]].  workspace >= caustics_mag_workspace_bytes(max_count, m, ...) for any m >= 1. */
int caustics_mag_extended_source_list(const void* w, double* mag, const int32_t* list, const int32_t* count,
                                      int64_t max_count, double rho, const caustics_lens* lens, int npts_limb,
                                      int limb_darkening, double u1, int npts_ld, int itmax, int compensated,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* The gate alone (binary lens, lightcurve.py:202-225): hexadecapole magnification of every point in
 * mag (n), the validity decision in used_hexadecapole (n, optional), and the compacted indices of
 * the points that FAIL the tests in list (capacity n) with their number in *count (device int32).
 * Lets a multi-GPU driver balance the expensive full integrations before dealing them out. */
int caustics_mag_gate(const void* w, double* mag, uint8_t* used_hexadecapole, int32_t* list, int32_t* count,
                      int64_t n, double rho, const caustics_lens* lens, double q, int itmax, int compensated,
                      void* stream);

/* ---- either side of the magnification in a light-curve likelihood -------------
 * caustics_trajectory: AnnualParallaxTrajectory.compute (src/caustics/trajectory.py:106-158) with the
 * Sun's projected position/velocity tables (t_jpl, s_e, s_n, s_e_dot, s_n_dot: n_jpl doubles each,
 * what trajectory.py:60-104 derives from the JPL ephemeris) passed in; n_jpl = 0: rectilinear motion.
 * t (n) float64 -> w (n) complex128 = u_e + i u_n.  psi/piE are the polar parametrisation
 * (piEE = piE sin psi, piEN = piE cos psi).
 * caustics_marginalized_log_likelihood: marginalized_log_likelihood, diagonal covariance
 * (src/caustics/linalg.py:55-70) for ONE light curve: out (3 doubles, device) = F_s, F_b, log-likelihood.
 * Deterministic (one CTA, fixed summation order).  All pointers are device pointers; nothing
 * synchronises, so trajectory -> caustics_mag -> likelihood is one stream-ordered enqueue. */
int caustics_trajectory(const double* t, void* w, int64_t n, double t0, double tE, double u0, double psi,
                        double piE, const double* t_jpl, const double* s_e, const double* s_n,
                        const double* s_e_dot, const double* s_n_dot, int n_jpl, void* stream);
int caustics_marginalized_log_likelihood(const double* mag, const double* fobs, const double* c_inv, int64_t n,
                                         double* out, void* stream);

/* Contours of the images of the source limb, for differentiating the uniform-disk magnification on
 * the host side (the implicit-function rule stays in Python, north_star): per source, the vertices
 * of every closed contour in integration order, closing vertex included.  vz (VMAX, n) complex128,
 * vtheta (VMAX, n) limb angle of each vertex, vcid (VMAX, n) contour id, vcount (n), cpar (CMAX, n)
 * contour parity, cstart (CMAX + 1, n) first vertex of each contour, ncont (n); source index fastest.
 * mag (n), optional, receives the uniform-disk magnification.  VMAX, CMAX from _capacity. */
int caustics_ext_contour_capacity(int nlenses, int npts_limb, int* vmax, int* cmax);
int caustics_ext_contours(const void* w, double* mag, int64_t n, double rho, const caustics_lens* lens,
                          int npts_limb, int itmax, int compensated, void* workspace, size_t workspace_bytes,
                          void* vz, double* vtheta, uint8_t* vcid, int32_t* vcount, double* cpar,
                          int32_t* cstart, int32_t* ncont, void* stream);

/* ---- track matching (critical / caustic curves, point_source.py:1582-1649) -------------------
 * z (nsets, npts, deg) complex128 -> out, same shape: every row re-ordered so that entry i continues
 * entry i of the previous row (greedy nearest neighbour without reuse, utils.py:15-40). deg <= 16. */
int caustics_match_tracks(const void* z, void* out, int64_t nsets, int npts, int deg, void* stream);

/* ---- multi-GPU: the final result gather without a collective ---------------------
 * Source positions are independent (lightcurve.py:245-254, point_source.py:1762-1830,
 * cpu_ops.cc:45-72), so the only inter-GPU traffic is the gather of the results.  Instead of a
 * collective after the kernels, the destination rank exports one buffer and every other rank passes
 * `opened base + its slice offset` as the result pointer of the launchers above: the kernels' own
 * stores travel over NVLink while they compute.  alloc/export on the destination, open/close on the
 * writers (one process per GPU, CUDA IPC; handle = CAUSTICS_PEER_HANDLE_BYTES opaque bytes moved by
 * any host-side means); a single process driving several GPUs calls caustics_peer_enable(owner
 * device) on each writer device instead and uses the owner's pointer as is.  After the writers have
 * synchronised their streams and a host-side barrier, the destination reads the assembled result. */
#define CAUSTICS_PEER_HANDLE_BYTES 64
int caustics_peer_alloc(void** ptr, size_t bytes);
int caustics_peer_free(void* ptr);
int caustics_peer_export(void* ptr, void* handle);
int caustics_peer_open(const void* handle, void** ptr);
int caustics_peer_close(void* ptr);
int caustics_peer_enable(int peer_device);

/* ---- launch-shape overrides for tests and experiments ---------------------------------------
 * key in {"grid_run", "path_run", "grid_extrap", "ext_variants", "open_wsmall", "host_slots",
 * "host_chunk_log2", "ext_split", "ext_windows"}; value -1 restores the launcher's own rule.
 * ext_variants: phase-variant mask of the extended-source pipeline (csrc/extended.cu);
 * open_wsmall: sources with at most this many marked tracks go to the first open-pass launch (default 2);
 * host_slots (1..8, default 4) / host_chunk_log2 (10..24, default 15): shape of the host-buffer pipeline;
 * ext_windows (1..4) / ext_split (length of the first window, 0 = one window): the windows of an un-gated call.
 * (These replace environment variables: nothing in the library calls getenv.) */
int caustics_set_tuning(const char* key, int value);

/* ---- measurement aid -----------------------------------------------------------------------
 * Launches blocks x 256 threads, each running 8 independent chains of `iters` double-precision
 * FMAs (2 * 8 * 256 * blocks * iters flop).  bench.py times it to get the FP64 roofline peak. */
int caustics_bench_fp64_peak(double* sink, int blocks, int iters, void* stream);
/* same flop count (iters rounded up to a multiple of 3), but every DFMA reads three distinct,
 * changing register pairs: the register-file-limited DFMA rate */
int caustics_bench_fp64_peak3(double* sink, int blocks, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAUSTICS_B200_H_ */
