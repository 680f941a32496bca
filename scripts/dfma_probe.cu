// DFMA issue-rate probe: how the FP64 pipe rate depends on how many FRESH register pairs an
// instruction reads (operand-reuse cache hits do not touch the register file).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_variants/dfma_probe scripts/dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

// 1 fresh + 2 loop-invariant operands
__global__ void __launch_bounds__(256) k_fresh1(double* sink, int iters, double seed) {
  double m = 1.0000001 + seed * 1e-12, c = 1e-9 * threadIdx.x;
#define D(i) double a##i = seed + i;
  CHAINS8(D)
#undef D
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
#define S(i) a##i = fma(a##i, m, c);
    CHAINS8(S)
#undef S
  }
  double r = 0;
#define R(i) r += a##i;
  CHAINS8(R)
#undef R
  if (r == 123.456) sink[0] = r;
}

// 2 fresh + 1 loop-invariant multiplier
__global__ void __launch_bounds__(256) k_fresh2(double* sink, int iters, double seed) {
  double m = 1.0000001 + seed * 1e-12;
#define D(i) double a##i = seed + i, b##i = 1e-9 * threadIdx.x + i;
  CHAINS8(D)
#undef D
#pragma unroll 2
  for (int i = 0; i < iters; i += 2) {
#define S(i) a##i = fma(m, a##i, b##i);
    CHAINS8(S)
#undef S
#define S(i) b##i = fma(m, b##i, a##i);
    CHAINS8(S)
#undef S
  }
  double r = 0;
#define R(i) r += a##i + b##i;
  CHAINS8(R)
#undef R
  if (r == 123.456) sink[0] = r;
}

// 3 fresh
__global__ void __launch_bounds__(256) k_fresh3(double* sink, int iters, double seed) {
#define D(i) double a##i = seed + i, b##i = 1.0000001 + 1e-7 * i, c##i = 1e-9 * threadIdx.x + 1e-9 * i;
  CHAINS8(D)
#undef D
#pragma unroll 2
  for (int i = 0; i < iters; i += 3) {
#define S(i) a##i = fma(a##i, b##i, c##i);
    CHAINS8(S)
#undef S
#define S(i) b##i = fma(b##i, c##i, a##i);
    CHAINS8(S)
#undef S
#define S(i) c##i = fma(c##i, a##i, b##i);
    CHAINS8(S)
#undef S
  }
  double r = 0;
#define R(i) r += a##i + b##i + c##i;
  CHAINS8(R)
#undef R
  if (r == 123.456) sink[0] = r;
}

// complex Horner shape: 4 independent complex accumulators, shared evaluation point z per thread,
// "coefficients" rotating through 8 register pairs (fresh every step)
__global__ void __launch_bounds__(256) k_horner(double* sink, int iters, double seed) {
  const double zr = 0.70710678 + seed * 1e-12, zi = 0.70710677 - 1e-9 * threadIdx.x;
  double pr[8], pi[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { pr[k] = 1e-3 * (k + 1) + seed; pi[k] = -1e-3 * (k + 2) + 1e-9 * threadIdx.x; }
  double tr[4], ti[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { tr[c] = seed + c; ti[c] = seed - c; }
#pragma unroll 1
  for (int i = 0; i < iters; i += 8 * 4 * 4) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double otr = tr[c], oti = ti[c];
        const double ur = fma(zr, otr, pr[(k + c) & 7]);
        const double ui = fma(zr, oti, pi[(k + c) & 7]);
        tr[c] = fma(-zi, oti, ur);
        ti[c] = fma(zi, otr, ui);
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) r += tr[c] + ti[c];
  if (r == 123.456) sink[0] = r;
}

template <class K>
static double run(K kern, const char* name, int per_iter_scale) {
  double* sink;
  cudaMalloc(&sink, 8);
  const int blocks = 148 * 8, iters = 1 << 15;
  kern<<<blocks, 256>>>(sink, 1024, 0.5);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    kern<<<blocks, 256>>>(sink, iters, 0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double dfma = (double)blocks * 256 * (double)iters * per_iter_scale;
  const double tf = 2.0 * dfma / (best * 1e-3) / 1e12;
  printf("%-10s %8.3f ms  %6.2f TFLOP/s\n", name, best, tf);
  cudaFree(sink);
  return tf;
}

int main() {
  run(k_fresh1, "fresh1", 8);
  run(k_fresh2, "fresh2", 8);
  run(k_fresh3, "fresh3", 8);
  run(k_horner, "horner", 1);
  return 0;
}
