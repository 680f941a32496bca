// HOST-buffer entry points: the call a reference-side binding makes with host arrays
// (the shape of lib/ehrlich_aberth/cpu_ops.cc:15-81, whose operands are host pointers).
// The batch is cut into chunks that flow through a few slots (stream + buffers) as H2D -> kernel -> D2H, so copies
// in both directions overlap compute when the caller's memory is pinned (pageable memory works,
// the driver then stages the copies itself).  The device workspace is owned here, per device,
// grown on demand and reused across calls; nothing else in the library allocates.
//
// Concurrency: one lock PER DEVICE (a single-process driver of eight GPUs runs eight calls at once,
// SURVEY 8b); two callers on the same device take turns.  Every exit path -- errors included --
// waits for the slot streams first, so no copy into the caller's buffers is in flight on return.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <mutex>

#include "../../include/caustics_b200.h"
#include "nvtx_range.h"
#include "tuning.h"

namespace {

constexpr int NSLOT = 8;          // slots that exist; nslots() of them are used
constexpr int NSLOT_DEFAULT = 4;
inline int nslots() {
  const int v = cb200::tuning_get(cb200::TUNE_HOST_SLOTS);
  return v >= 1 && v <= NSLOT ? v : NSLOT_DEFAULT;
}
constexpr int MAXDEV = 16;

struct Slot {
  cudaStream_t st = nullptr;
  cudaEvent_t in_ready = nullptr, in_free = nullptr;   // input landed (H2D stream) / input consumed (slot stream)
  void* d_in = nullptr;
  void* d_in2 = nullptr;
  void* d_out = nullptr;
  void* d_out2 = nullptr;
  size_t cap_in = 0, cap_in2 = 0, cap_out = 0, cap_out2 = 0;
};
struct Workspace {
  std::mutex mu;
  bool init = false;
  cudaStream_t h2d = nullptr;     // all host-to-device copies of the solver pipeline, back to back
  Slot slot[NSLOT];
};
Workspace g_ws[MAXDEV];

inline int rc_of(cudaError_t e) { return e == cudaSuccess ? CAUSTICS_OK : CAUSTICS_ERR_CUDA_BASE + (int)e; }

#define CK(expr)                      \
  do {                                \
    cudaError_t e__ = (expr);         \
    if (e__ != cudaSuccess) return rc_of(e__); \
  } while (0)

int ensure(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return CAUSTICS_OK;
  if (*p) CK(cudaFree(*p));
  *p = nullptr; *cap = 0;
  CK(cudaMalloc(p, need));
  *cap = need;
  return CAUSTICS_OK;
}

int current_ws(Workspace** out) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAXDEV) return CAUSTICS_ERR_BAD_ARG;
  *out = &g_ws[dev];
  return CAUSTICS_OK;
}

// call with w.mu held
int init_ws(Workspace& w) {
  if (!w.init) {
    for (int i = 0; i < NSLOT; ++i) {
      CK(cudaStreamCreateWithFlags(&w.slot[i].st, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&w.slot[i].in_ready, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&w.slot[i].in_free, cudaEventDisableTiming));
    }
    CK(cudaStreamCreateWithFlags(&w.h2d, cudaStreamNonBlocking));
    w.init = true;
  }
  return CAUSTICS_OK;
}

// wait for everything enqueued on the slot streams; returns the first error seen (or `rc` if set)
int drain(Workspace& w, int rc) {
  if (w.h2d) {
    const cudaError_t e = cudaStreamSynchronize(w.h2d);
    if (!rc && e != cudaSuccess) rc = rc_of(e);
  }
  for (int i = 0; i < NSLOT; ++i) {
    if (!w.slot[i].st) continue;
    const cudaError_t e = cudaStreamSynchronize(w.slot[i].st);
    if (!rc && e != cudaSuccess) rc = rc_of(e);
  }
  return rc;
}

// chunk length: large enough to amortise launch + copy latency, small enough that the pipeline's
// fill (first H2D) and drain (last kernel + D2H) stay a small part of the call (measured on one B200:
// 32 Ki polynomials; 64 Ki .. 256 Ki and ramped schedules are slower).
inline int64_t pick_chunk(int64_t n) {
  const int lg = cb200::tuning_get(cb200::TUNE_HOST_CHUNK_LOG2);
  const int64_t c = (int64_t)1 << (lg >= 10 && lg <= 24 ? lg : 15);
  return c < n ? c : n;
}

int ea_solve_host_locked(Workspace& w, const void* coeffs, const void* roots_init, void* roots, int32_t* sweeps,
                         int64_t size, int deg, int itmax, int compensated, int custom_init, int flags) {
  int rc = init_ws(w);
  if (rc) return rc;
  const int64_t chunk = pick_chunk(size);
  const int ns = nslots();
  const size_t bc = (size_t)(deg + 1) * 16, br = (size_t)deg * 16;
  for (int i = 0; i < ns; ++i) {
    Slot& s = w.slot[i];
    if ((rc = ensure(&s.d_in, &s.cap_in, chunk * bc))) return rc;
    if ((rc = ensure(&s.d_out, &s.cap_out, chunk * br))) return rc;
    if (custom_init && (rc = ensure(&s.d_in2, &s.cap_in2, chunk * br))) return rc;
    if (sweeps && (rc = ensure(&s.d_out2, &s.cap_out2, chunk * 4))) return rc;
  }
  // The host-to-device copies run back to back on their own stream -- a slot's next input does not queue behind the
  // slot's previous device-to-host copy, only behind the kernel that read the buffer (in_free) -- and the slot stream
  // (kernel, then the copy back) waits for its input (in_ready).  4.26 -> 4.03 ms per C2 step with four slots (profiles/r02_e2e_pipeline_probe.txt).
  int k = 0;
  for (int64_t off = 0; off < size; off += chunk, ++k) {
    Slot& s = w.slot[k % ns];
    const int64_t m = (size - off < chunk) ? size - off : chunk;
    if (k >= ns) CK(cudaStreamWaitEvent(w.h2d, s.in_free, 0));
    CK(cudaMemcpyAsync(s.d_in, (const char*)coeffs + off * bc, m * bc, cudaMemcpyHostToDevice, w.h2d));
    if (custom_init)
      CK(cudaMemcpyAsync(s.d_in2, (const char*)roots_init + off * br, m * br, cudaMemcpyHostToDevice, w.h2d));
    CK(cudaEventRecord(s.in_ready, w.h2d));
    CK(cudaStreamWaitEvent(s.st, s.in_ready, 0));
    rc = caustics_ea_solve(s.d_in, custom_init ? s.d_in2 : nullptr, s.d_out, sweeps ? (int32_t*)s.d_out2 : nullptr,
                           m, deg, itmax, compensated, custom_init, flags, s.st);
    if (rc) return rc;
    CK(cudaEventRecord(s.in_free, s.st));
    CK(cudaMemcpyAsync((char*)roots + off * br, s.d_out, m * br, cudaMemcpyDeviceToHost, s.st));
    if (sweeps) CK(cudaMemcpyAsync(sweeps + off, s.d_out2, m * 4, cudaMemcpyDeviceToHost, s.st));
  }
  return CAUSTICS_OK;
}

int mag_ps_host_locked(Workspace& w, const void* wpts, double* mag, int64_t n, const caustics_lens* lens,
                       int itmax, int compensated, int flags) {
  int rc = init_ws(w);
  if (rc) return rc;
  const int64_t chunk = pick_chunk(n);
  const int ns = nslots();
  for (int i = 0; i < ns; ++i) {
    Slot& s = w.slot[i];
    if ((rc = ensure(&s.d_in, &s.cap_in, chunk * 16))) return rc;
    if ((rc = ensure(&s.d_out, &s.cap_out, chunk * 8))) return rc;
  }
  int k = 0;
  for (int64_t off = 0; off < n; off += chunk, ++k) {
    Slot& s = w.slot[k % ns];
    const int64_t m = (n - off < chunk) ? n - off : chunk;
    CK(cudaMemcpyAsync(s.d_in, (const char*)wpts + off * 16, m * 16, cudaMemcpyHostToDevice, s.st));
    rc = caustics_mag_point_source(s.d_in, (double*)s.d_out, nullptr, m, lens, itmax, compensated, flags, s.st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(mag + off, s.d_out, m * 8, cudaMemcpyDeviceToHost, s.st));
  }
  return CAUSTICS_OK;
}

// Row blocks of a magnification map: kernel on block k overlaps the D2H of block k-1.  Blocks are whole
// multiples of 32 rows so that a walked map (CAUSTICS_FLAG_GRID_WALK) is cut on walk boundaries and the
// result does not depend on the chunking.
int mag_grid_host_locked(Workspace& w, double x0, double y0, double dx, double dy, int64_t nx, int64_t row_begin,
                         int64_t row_end, double* mag, const caustics_lens* lens, int itmax, int compensated,
                         int flags) {
  int rc = init_ws(w);
  if (rc) return rc;
  int64_t rows = ((int64_t)1 << 21) / nx;          // ~2 Mi pixels = 16 MB per block
  rows = rows < 32 ? 32 : (rows / 32) * 32;
  const int ns = nslots();
  for (int i = 0; i < ns; ++i)
    if ((rc = ensure(&w.slot[i].d_out, &w.slot[i].cap_out, (size_t)rows * nx * 8))) return rc;
  int k = 0;
  for (int64_t r0 = row_begin; r0 < row_end; r0 += rows, ++k) {
    Slot& s = w.slot[k % ns];
    const int64_t r1 = r0 + rows < row_end ? r0 + rows : row_end;
    rc = caustics_mag_point_source_grid(x0, y0, dx, dy, nx, r0, r1, (double*)s.d_out, lens, itmax, compensated, flags, s.st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(mag + (r0 - row_begin) * nx, s.d_out, (size_t)(r1 - r0) * nx * 8, cudaMemcpyDeviceToHost, s.st));
  }
  return CAUSTICS_OK;
}

}  // namespace

extern "C" {

void caustics_release_workspace(void) {
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }
  for (int d = 0; d < MAXDEV; ++d) {
    Workspace& w = g_ws[d];
    std::lock_guard<std::mutex> lk(w.mu);
    if (!w.init) continue;
    cudaSetDevice(d);
    for (int i = 0; i < NSLOT; ++i) {
      Slot& s = w.slot[i];
      cudaStreamSynchronize(s.st);
      cudaFree(s.d_in); cudaFree(s.d_in2); cudaFree(s.d_out); cudaFree(s.d_out2);
      cudaStreamDestroy(s.st);
      if (s.in_ready) cudaEventDestroy(s.in_ready);
      if (s.in_free) cudaEventDestroy(s.in_free);
      s = Slot();
    }
    if (w.h2d) { cudaStreamSynchronize(w.h2d); cudaStreamDestroy(w.h2d); w.h2d = nullptr; }
    w.init = false;
  }
  cudaSetDevice(cur);
}

int caustics_ea_solve_host(const void* coeffs, const void* roots_init, void* roots, int32_t* sweeps,
                           int64_t size, int deg, int itmax, int compensated, int custom_init,
                           int flags) {
  if (size < 0 || itmax < 0) return CAUSTICS_ERR_BAD_ARG;
  if (!caustics_ea_degree_supported(deg)) return CAUSTICS_ERR_UNSUPPORTED_DEGREE;
  if (size == 0) return CAUSTICS_OK;
  if (!coeffs || !roots || (custom_init && !roots_init)) return CAUSTICS_ERR_BAD_ARG;
  CB200_NVTX("caustics_ea_solve_host");
  Workspace* w;
  int rc = current_ws(&w);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(w->mu);
  rc = ea_solve_host_locked(*w, coeffs, roots_init, roots, sweeps, size, deg, itmax, compensated, custom_init, flags);
  return drain(*w, rc);
}

int caustics_mag_point_source_host(const void* wpts, double* mag, int64_t n, const caustics_lens* lens,
                                   int itmax, int compensated, int flags) {
  if (n < 0 || itmax < 0 || !lens) return CAUSTICS_ERR_BAD_ARG;
  if (n == 0) return CAUSTICS_OK;
  if (!wpts || !mag) return CAUSTICS_ERR_BAD_ARG;
  CB200_NVTX("caustics_mag_point_source_host");
  Workspace* w;
  int rc = current_ws(&w);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(w->mu);
  rc = mag_ps_host_locked(*w, wpts, mag, n, lens, itmax, compensated, flags);
  return drain(*w, rc);
}

int caustics_mag_point_source_grid_host(double x0, double y0, double dx, double dy, int64_t nx, int64_t row_begin,
                                        int64_t row_end, double* mag, const caustics_lens* lens, int itmax,
                                        int compensated, int flags) {
  if (nx <= 0 || row_end < row_begin || itmax < 0 || !lens) return CAUSTICS_ERR_BAD_ARG;
  if (row_end == row_begin) return CAUSTICS_OK;
  if (!mag) return CAUSTICS_ERR_BAD_ARG;
  CB200_NVTX("caustics_mag_point_source_grid_host");
  Workspace* w;
  int rc = current_ws(&w);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(w->mu);
  rc = mag_grid_host_locked(*w, x0, y0, dx, dy, nx, row_begin, row_end, mag, lens, itmax, compensated, flags);
  return drain(*w, rc);
}

}  // extern "C"
