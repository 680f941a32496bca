// NVTX ranges around the C-ABI launchers (tracing).  NVTX3 is header-only and a
// no-op unless a profiler (nsys / ncu --nvtx) has injected its library, so the ranges cost a
// function-pointer test per call.  CB200_NVTX("name") opens a range that closes at scope exit.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace cb200 {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace cb200
#define CB200_NVTX(name) ::cb200::NvtxRange cb200_nvtx_range__(name)
