// Microbenchmark: does SHFL share the L1 data-pipe wavefront budget with unaligned row gathers?
// Each warp repeatedly gathers 32 consecutive floats starting at an unaligned offset of a small,
// L1-resident buffer (2 wavefronts per LDG) in four mixes. Build: nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ src, float* __restrict__ out, int iters, int off) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* base = src + warp * 512 + off + lane;   // per-warp 2 KB window, unaligned by `off`
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const float* p = base + (it & 7) * 32;
#pragma unroll
    for (int r = 0; r < 12; ++r) {
      float lo = __ldg(p + r * 40);
      float hi;
      if (MODE == 0) hi = __ldg(p + r * 40 + 1);                          // second unaligned gather
      if (MODE == 1) hi = __shfl_down_sync(0xffffffffu, lo, 1);           // neighbour lane instead
      if (MODE == 2) {                                                    // shuffle + rarely taken fallback load
        hi = __shfl_down_sync(0xffffffffu, lo, 1);
        if (lane == 31) hi = __ldg(p + r * 40 + 1);
      }
      if (MODE == 3) hi = lo * 1.5f;                                      // lower gathers only
      acc = fmaf(lo, 0.5f, acc) + hi;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
float run(const float* src, float* out, int iters, int off) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(src, out, 10, off);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(src, out, iters, off);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float *src, *out;
  cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 2000;
  for (int off = 0; off <= 1; ++off) {
    float t0 = run<0>(src, out, iters, off), t1 = run<1>(src, out, iters, off), t2 = run<2>(src, out, iters, off), t3 = run<3>(src, out, iters, off);
    // cycles per (row) group per SM: 8 CTAs x 8 warps x iters x 12 groups per SM
    const double groups = 64.0 * iters * 12, clk = 1.965e6;  // cycles per ms
    printf("offset %d: LDG+LDG %.3f ms (%.2f clk/group/SM) | LDG+SHFL %.3f (%.2f) | LDG+SHFL+fallback %.3f (%.2f) | LDG only %.3f (%.2f)\n",
           off, t0, t0 * clk / groups, t1, t1 * clk / groups, t2, t2 * clk / groups, t3, t3 * clk / groups);
  }
  return 0;
}
