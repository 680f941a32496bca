// TEST INFRASTRUCTURE ONLY.  Lets g++ compile the product's device headers (ea_core.cuh,
// lens_core.cuh, ...) for the host with a "warp" of ONE lane, so the algorithmic logic can be
// checked against the oracle in the no-GPU test tier.  It is never built into the product library
// and is not a CPU fallback: nothing under caustics_b200/ can reach it.
#pragma once
#define CB200_HOSTSIM 1
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(x)
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int __double2loint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)u; }
static inline double __hiloint2double(int hi, int lo) {
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
}
static inline bool __all_sync(unsigned, bool p) { return p; }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
template <class T> static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int, int = 32) { return v; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs(x); }
static inline void sincospi(double x, double* s, double* c) { *s = sin(3.14159265358979323846 * x); *c = cos(3.14159265358979323846 * x); }
using std::max;
using std::min;
