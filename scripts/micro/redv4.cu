// microbenchmark: three scalar REDs into planar channel volumes vs ONE 128-bit vector RED into a
// channel-packed volume (red.global.v4.f32.add, sm_90+), same warp access pattern as the splat
// kernels (lane = z, 8 corner rows per voxel, smooth pseudo-displacement).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o redv4 redv4.cu && ./redv4
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) splat(float* planar, float4* packed, int X, int Y, int Z) {
  const int k = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y, i = blockIdx.z;
  const int V = X * Y * Z;
  // pseudo displacement: a smooth shift of up to ~2 voxels
  const int dx = (i + (j >> 4)) & 1, dy = (j + (k >> 5)) & 1, dz = (k >> 3) & 1;
  const int x0 = min(i + dx, X - 2), y0 = min(j + dy, Y - 2), z0 = min(k + dz, Z - 2);
  const float w = 0.125f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int idx = ((x0 + (c >> 2)) * Y + (y0 + ((c >> 1) & 1))) * Z + z0 + (c & 1);
    if (MODE == 0) {
      atomicAdd(planar + idx, w);
      atomicAdd(planar + V + idx, w);
      atomicAdd(planar + 2 * (size_t)V + idx, w);
    } else if (MODE == 1) {
      asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(packed + idx), "f"(w), "f"(w), "f"(w), "f"(0.f) : "memory");
    } else {
      asm volatile("red.global.v2.f32.add [%0], {%1, %2};" ::"l"((float2*)packed + idx), "f"(w), "f"(w) : "memory");
    }
  }
}
int main() {
  const int X = 256, Y = 256, Z = 256;
  float* p; float4* q;
  cudaMalloc(&p, sizeof(float) * 3 * X * Y * Z);
  cudaMalloc(&q, sizeof(float4) * X * Y * Z);
  cudaMemset(p, 0, sizeof(float) * 3 * X * Y * Z);
  cudaMemset(q, 0, sizeof(float4) * X * Y * Z);
  dim3 grid(Z / 32, Y / 8, X), block(32, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0);
      if (mode == 0) splat<0><<<grid, block>>>(p, q, X, Y, Z);
      else if (mode == 1) splat<1><<<grid, block>>>(p, q, X, Y, Z);
      else splat<2><<<grid, block>>>(p, q, X, Y, Z);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%s: %.3f ms for %d M voxels x 8 corners x 3 channels (%s)\n", mode == 0 ? "3 x RED.32 planar" : mode == 1 ? "1 x RED.v4 packed" : "1 x RED.v2 (2 of 3 channels)", best, X * Y * Z >> 20, cudaGetErrorString(cudaGetLastError()));
  }
  float h[8]; cudaMemcpy(h, q + (128 * 256 + 128) * 256 + 128, 16, cudaMemcpyDeviceToHost);
  printf("check packed value %.3f %.3f %.3f %.3f\n", h[0], h[1], h[2], h[3]);
  return 0;
}
