// gather_common.cuh -- coordinate / corner helpers shared by the fp32 3-D gather and splat kernels
// (gather3.cu forward paths, epdiff_bwd.cu fused backward).
#pragma once
#include "common.cuh"

namespace lgm {

// RN_f32(fi + d*u) for d = dh + dl (double split in two floats), product and sum carried as
// float pairs; equals the double-rounded reference value except on ~2^-21 of inputs (1 ulp).
__device__ __forceinline__ float coord_f32(float fi, float u, float dh, float dl) {
  float ph = __fmul_rn(dh, u);
  float pe = __fmaf_rn(dh, u, -ph);
  float pl = __fmaf_rn(dl, u, pe);
  float sh = __fadd_rn(fi, ph);
  float bb = __fsub_rn(sh, fi);
  float se = __fadd_rn(__fsub_rn(fi, __fsub_rn(sh, bb)), __fsub_rn(ph, bb));
  return __fadd_rn(sh, __fadd_rn(se, pl));
}

struct Ax3 {
  int i0, i1;
  float t;
};

// floor / fraction / clamped corner indices of one coordinate. Coordinates are first clamped to
// +-2^22 voxels (far outside any volume: both corners are the border voxel there, and the result is
// the border value for any weight), which lets floor() be a magic-number add with no slow path.
__device__ __forceinline__ Ax3 axis_fast(float x0, int n) {
  Ax3 a;
  float x = fminf(fmaxf(x0, -4194304.f), 4194304.f);
  float r = __fadd_rn(x, 12582912.f);  // 1.5 * 2^23: rounds x to an integer in the mantissa
  int f = __float_as_int(r) - 0x4B400000;
  float rf = __fsub_rn(r, 12582912.f);
  if (rf > x) {
    rf -= 1.f;
    f -= 1;
  }
  // a NaN / Inf coordinate must give a NaN weight like the reference's t = x - floor(x) (the clamp above
  // would turn it into a border sample): x0 * 0 is 0 for finite x0, NaN otherwise
  a.t = __fmaf_rn(x0, 0.f, x - rf);
  a.i0 = min(max(f, 0), n - 1);
  a.i1 = min(max(f + 1, 0), n - 1);
  return a;
}

// Cheaper twin of axis_fast() for the forward gather kernels. The coordinate is clamped to
// [-1, n - 1/2] (outside of it both corners are the border voxel and the result is the border value for
// any weight, see z_pair()), floor() is ONE round-down add of the magic number, and the clamps of the
// two corner indices shrink to one instruction each because floor is already in [-1, n-1].
__device__ __forceinline__ Ax3 axis_fwd(float x0, int n, float hi) {  // hi = n - 0.5f
  Ax3 a;
  const float x = fminf(fmaxf(x0, -1.f), hi);
  const float r = __fadd_rd(x, 12582912.f);  // floor(x) + 1.5 * 2^23, exact
  const int f = __float_as_int(r) - 0x4B400000;
  a.t = __fmaf_rn(x0, 0.f, x - __fsub_rn(r, 12582912.f));  // NaN / Inf coordinates stay NaN (see axis_fast)
  a.i0 = max(f, 0);
  a.i1 = min(f + 1, n - 1);
  return a;
}

// base + 4*idx as ONE IMAD.WIDE.U32 whose multiplier lives in a register for the whole kernel. With
// a literal 4 and the (uniform) base in a uniform register ptxas has to materialise the constant with
// a MOV in front of every address: 10 % of the gather kernels' issue slots. An empty asm() does not
// hide the constant from ptxas, a load from a mutable __device__ word does (one LDG per thread).
// (indexed by the lane so that the load is not uniform: a uniform-register copy would again be MOVed
// into a vector register at every use)
static __device__ unsigned g_four[32] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
                                         4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};
__device__ __forceinline__ unsigned opaque_four() { return g_four[threadIdx.x & 31]; }
template <typename T>
__device__ __forceinline__ T* at4(T* base, unsigned idx, unsigned four) {
  return (T*)((char*)base + (unsigned long long)idx * four);
}
template <typename T>
__device__ __forceinline__ const T* at4(const T* base, unsigned idx, unsigned four) {
  return (const T*)((const char*)base + (unsigned long long)idx * four);
}

// Nested lerp of the 8 corner values (corner numbering / evaluation order of include/interp.h:115-122):
//   omv*(omu*(omt*v0 + t*v1) + u*(omt*v3 + t*v2)) + v*(omu*(omt*v4 + t*v5) + u*(omt*v7 + t*v6)).
// Every a*x + b*y is spelled fma(b, y, a*x) explicitly, so that all kernels sampling through this
// function (L1 gathers, the shared-memory ring of compose_ring.cu) round identically whatever
// contraction the compiler would have picked in their context.
__device__ __forceinline__ float lerp8(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7,
                                       float t, float u, float v, float omt, float omu, float omv) {
  const float a0 = __fmaf_rn(t, v1, __fmul_rn(omt, v0));
  const float a1 = __fmaf_rn(t, v2, __fmul_rn(omt, v3));
  const float a2 = __fmaf_rn(t, v5, __fmul_rn(omt, v4));
  const float a3 = __fmaf_rn(t, v6, __fmul_rn(omt, v7));
  const float b0 = __fmaf_rn(u, a1, __fmul_rn(omu, a0));
  const float b1 = __fmaf_rn(u, a3, __fmul_rn(omu, a2));
  return __fmaf_rn(v, b1, __fmul_rn(omv, b0));
}

// 8-corner gather + nested lerp (corner numbering / evaluation order of include/interp.h:91-122).
// i00..i11 are the element indices of the four (x,y) corner rows at the lower z corner; the upper z
// corner is always the +1 neighbour (an immediate offset on the same address register), see z_pair().
__device__ __forceinline__ float trilerp(const float* __restrict__ img, unsigned i00, unsigned i01,
                                         unsigned i10, unsigned i11, float t, float u, float v,
                                         float omt, float omu, float omv, unsigned four) {
  const float* p00 = at4(img, i00, four);
  const float* p01 = at4(img, i01, four);
  const float* p10 = at4(img, i10, four);
  const float* p11 = at4(img, i11, four);
  float v0 = __ldg(p00), v4 = __ldg(p00 + 1);
  float v3 = __ldg(p01), v7 = __ldg(p01 + 1);
  float v1 = __ldg(p10), v5 = __ldg(p10 + 1);
  float v2 = __ldg(p11), v6 = __ldg(p11 + 1);
  return lerp8(v0, v1, v2, v3, v4, v5, v6, v7, t, u, v, omt, omu, omv);
}

// Along the contiguous axis the two corners are fetched as (zs, zs+1) with zs <= Z-2 so that the
// pair never leaves the row. Where the reference clamps both corners onto one border voxel
// ((1-v)*B + v*B) the weight is replaced by 0 (lower border) or 1 (upper border): the same value
// up to the rounding of (1-v)*B + v*B, i.e. <= 1 ulp, and only outside the volume.
__device__ __forceinline__ void z_pair(const Ax3& az, int Z, int& zs, float& v) {
  zs = min(az.i0, Z - 2);
  v = az.t;
  if (az.i1 == az.i0) v = __fmaf_rn(az.t, 0.f, (az.i0 == 0) ? 0.f : 1.f);  // keeps a NaN weight NaN
}

// L2 prefetch for the row-walking kernels (grid = (z chunks, y tiles of 8 rows, N * X), blockDim = (32, 8)):
// thread t < NTHR of warp 0 of the CTAs with blockIdx.x == 0 pulls the 8 rows x Z of channel (t % 3) of
// volume SEL (an expression of t choosing among the kernel's 3-channel inputs) that the CTA LGM_PF_ROWS
// (y, z) grid positions AHEAD in launch order will read first. These kernels open with loads whose DRAM
// latency only the other resident CTAs hide; with the rows already in L2 that latency is an L2 hit
// (measured on B200: Ad_star 0.444 -> 0.399 ms at C2, 1.745 -> 1.485 ms at C3; an empty asm in place of
// the prefetch instruction gives 0.444, so it is the prefetch, not a scheduling side effect).
// A macro, not a function: the function form of the same code measured no gain (profiles/r2_notes.md).
#ifndef LGM_PF_ROWS
#define LGM_PF_ROWS 148  /* distance in CTAs; 0 = off */
#endif
#define LGM_PREFETCH_ROWS_AHEAD(NTHR, SEL, X, Y, Z, REV)                                                          \
  if (LGM_PF_ROWS > 0 && blockIdx.x == 0 && threadIdx.y == 0 && threadIdx.x < (NTHR)) {                            \
    const unsigned long long lin_ = (unsigned long long)blockIdx.z * gridDim.y + blockIdx.y + LGM_PF_ROWS;         \
    if (lin_ < (unsigned long long)gridDim.y * gridDim.z) {                                                        \
      const unsigned pby0_ = (unsigned)(lin_ % gridDim.y), pbz0_ = (unsigned)(lin_ / gridDim.y);                   \
      const unsigned pby_ = (REV) ? gridDim.y - 1 - pby0_ : pby0_, pbz_ = (REV) ? gridDim.z - 1 - pbz0_ : pbz0_;   \
      const int pj_ = pby_ * 8, pi_ = pbz_ % (X), pn_ = pbz_ / (X);                                                \
      if (pj_ + 8 <= (Y) && ((Z) & 3) == 0) {                                                                      \
        const float* src_ = (SEL) + ((size_t)pn_ * 3 + threadIdx.x % 3) * ((size_t)(X) * (Y) * (Z)) +              \
                            ((size_t)pi_ * (Y) + pj_) * (Z);                                                       \
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_), "r"((unsigned)(8 * (Z) * 4))        \
                     : "memory");                                                                                  \
      }                                                                                                            \
    }                                                                                                              \
  }

// RN_f32(g * d) for a double d = dh + dl: stands in for the reference's "(float)((double)g * dt)"
// (cuda/interp.cu:230) without fp64 instructions; equal except for rare double-rounding ties (1 ulp).
__device__ __forceinline__ float mul_f32_by_double(float g, float dh, float dl) {
  float ph = __fmul_rn(g, dh);
  float pe = __fmaf_rn(g, dh, -ph);
  return __fadd_rn(ph, __fmaf_rn(g, dl, pe));
}

}  // namespace lgm
