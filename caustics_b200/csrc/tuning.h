// Launch-shape overrides for tests and experiments (caustics_set_tuning in the C ABI).  The launchers
// read these atomics instead of environment variables: no getenv on the launch path, no hidden
// process state.  -1 = unset (the launcher's own rule applies).
#pragma once
#include <atomic>
#include <string.h>

namespace cb200 {
enum { TUNE_GRID_RUN = 0, TUNE_PATH_RUN, TUNE_GRID_EXTRAP, TUNE_EXT_VARIANTS, TUNE_OPEN_WSMALL, TUNE_HOST_SLOTS, TUNE_HOST_CHUNK_LOG2, TUNE_EXT_SPLIT, TUNE_EXT_WINDOWS, TUNE_COUNT };
inline std::atomic<int>* tuning_slots() {
  static std::atomic<int> v[TUNE_COUNT] = {{-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}};
  return v;
}
inline int tuning_get(int k) { return tuning_slots()[k].load(std::memory_order_relaxed); }
inline bool tuning_set(const char* key, int value) {
  static const char* const names[TUNE_COUNT] = {"grid_run", "path_run", "grid_extrap", "ext_variants", "open_wsmall", "host_slots", "host_chunk_log2", "ext_split", "ext_windows"};
  if (!key) return false;
  for (int k = 0; k < TUNE_COUNT; ++k)
    if (strcmp(key, names[k]) == 0) { tuning_slots()[k].store(value, std::memory_order_relaxed); return true; }
  return false;
}
}  // namespace cb200
