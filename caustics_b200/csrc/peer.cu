// Peer buffers: the final result gather of a multi-GPU run WITHOUT a collective (SURVEY 8e,
// BASELINE config C5: "no collective beyond the final result gather").
//
// The reference gathers nothing itself: jax.pmap / lax.map would return a host-side concatenation of
// the per-point results (lightcurve.py:245-254, point_source.py:1762-1830).  Here every rank's kernel
// stores its slice of the result straight into the destination rank's buffer over NVLink while it
// computes (8 B per evaluation against ~6e3 flop), so the "gather" has no separate phase at all:
//   destination rank : caustics_peer_alloc -> caustics_peer_export -> (handle travels by any means)
//   every other rank : caustics_peer_open  -> pass  base + slice offset  as the `mag` / `roots` pointer
//                      of any launcher in this header -> stream sync -> barrier -> caustics_peer_close
// One process per GPU uses CUDA IPC handles; a single process driving several GPUs calls
// caustics_peer_enable instead and uses the owner's pointer directly (unified addressing).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/caustics_b200.h"

namespace {
inline int rc_of(cudaError_t e) { return e == cudaSuccess ? CAUSTICS_OK : CAUSTICS_ERR_CUDA_BASE + (int)e; }
static_assert(sizeof(cudaIpcMemHandle_t) == CAUSTICS_PEER_HANDLE_BYTES, "IPC handle size is part of the ABI");
}  // namespace

extern "C" {

int caustics_peer_alloc(void** ptr, size_t bytes) {
  if (!ptr || bytes == 0) return CAUSTICS_ERR_BAD_ARG;
  *ptr = nullptr;
  return rc_of(cudaMalloc(ptr, bytes));      // a dedicated allocation: its IPC handle has offset 0
}

int caustics_peer_free(void* ptr) { return ptr ? rc_of(cudaFree(ptr)) : CAUSTICS_OK; }

int caustics_peer_export(void* ptr, void* handle) {
  if (!ptr || !handle) return CAUSTICS_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return rc_of(e);
  memcpy(handle, &h, sizeof(h));
  return CAUSTICS_OK;
}

int caustics_peer_open(const void* handle, void** ptr) {
  if (!handle || !ptr) return CAUSTICS_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  *ptr = nullptr;
  return rc_of(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
}

int caustics_peer_close(void* ptr) { return ptr ? rc_of(cudaIpcCloseMemHandle(ptr)) : CAUSTICS_OK; }

int caustics_peer_enable(int peer_device) {
  int cur = 0, can = 0;
  cudaError_t e = cudaGetDevice(&cur);
  if (e != cudaSuccess) return rc_of(e);
  if (peer_device == cur) return CAUSTICS_OK;
  e = cudaDeviceCanAccessPeer(&can, cur, peer_device);
  if (e != cudaSuccess) return rc_of(e);
  if (!can) return CAUSTICS_ERR_CUDA_BASE + (int)cudaErrorPeerAccessUnsupported;
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return CAUSTICS_OK; }
  return rc_of(e);
}

}  // extern "C"
