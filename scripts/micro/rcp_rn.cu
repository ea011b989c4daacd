// Is MUFU.RCP + one FMA Newton step == __frcp_rn (IEEE round-to-nearest reciprocal) for the range the
// Fourier multiplier uses? Exhaustive over every float in [2^-64, 2^64] (about 1.07e9 values).
//   nvcc -O3 -arch=sm_100a -o /tmp/rcp_rn scripts/micro/rcp_rn.cu && /tmp/rcp_rn
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float rcp_newton(float s) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
  const float e = __fmaf_rn(-s, r, 1.f);
  return __fmaf_rn(r, e, r);
}
__global__ void check(uint32_t lo, uint32_t hi, unsigned long long* bad, uint32_t* first) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= hi; b += stride) {
    const float s = __uint_as_float((uint32_t)b);
    const float a = __frcp_rn(s), c = rcp_newton(s);
    if (__float_as_uint(a) != __float_as_uint(c)) {
      if (atomicAdd(bad, 1ULL) == 0) *first = (uint32_t)b;
    }
  }
}
int main() {
  unsigned long long* bad; uint32_t* first;
  cudaMallocManaged(&bad, 8); cudaMallocManaged(&first, 4);
  *bad = 0; *first = 0;
  const float flo = 5.421010862427522e-20f /* 2^-64 */, fhi = 18446744073709551616.f /* 2^64 */;
  uint32_t lo, hi;
  memcpy(&lo, &flo, 4); memcpy(&hi, &fhi, 4);
  check<<<148 * 8, 256>>>(lo, hi, bad, first);
  cudaDeviceSynchronize();
  float f; memcpy(&f, first, 4);
  printf("floats checked: %llu, mismatches: %llu (first at %g)\n", (unsigned long long)hi - lo + 1, *bad, f);
  return 0;
}
