// gather3.cu -- fp32 3-D fast paths of the two gather kernels of an EPDiff step:
//   Ad_star : out_c = sum_d (D_d phi_c + delta_cd) * m_d(x + phi(x))          (adjrep.py:86-97)
//   compose : out_c = ds*u_c + dt*v_c(x + ds*u(x))                            (deform.py:53-55)
//
// Layout of the work: a thread owns 4 consecutive voxels along z (one float4 of every channel),
// a warp one 128-voxel z segment, a CTA 8 neighbouring y rows of one x slab, so that
//   - every direct load / store is a 16-byte vector, the Jacobian's y/x neighbours are float4
//     loads of adjacent rows (L1 hits across the CTA's rows) and its z neighbours come from the
//     thread's own vector plus two scalars;
//   - the 8-corner gathers of a warp land on runs of consecutive addresses for smooth flows
//     (served by L1/L2, the 24 scalar loads per voxel are the floor of this formulation);
//   - all index arithmetic is 32-bit, with no division in the hot loop.
// Sample coordinates reproduce the reference's "form in double, round to float"
// (cuda/interp.cu:68-73) with error-free float transformations instead of fp64/conversion
// instructions; floor() is a magic-number add.
#include "common.cuh"

#ifndef LGM_GATHER_BX
#define LGM_GATHER_BX 1  /* 2,4,8 measured equal on B200: the gathers are L1-data-pipe bound, not L2 bound */
#endif

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
__device__ __forceinline__ Ax3 axis_fast(float x, int n) {
  Ax3 a;
  x = fminf(fmaxf(x, -4194304.f), 4194304.f);
  float r = __fadd_rn(x, 12582912.f);  // 1.5 * 2^23: rounds x to an integer in the mantissa
  int f = __float_as_int(r) - 0x4B400000;
  float rf = __fsub_rn(r, 12582912.f);
  if (rf > x) {
    rf -= 1.f;
    f -= 1;
  }
  a.t = x - rf;
  a.i0 = min(max(f, 0), n - 1);
  a.i1 = min(max(f + 1, 0), n - 1);
  return a;
}

// 8-corner gather + nested lerp (corner numbering / evaluation order of include/interp.h:91-122).
// i00..i11 are the element indices of the four (x,y) corner rows at the lower z corner; the upper z
// corner is always the +1 neighbour (an immediate offset on the same address register), see z_pair().
__device__ __forceinline__ float trilerp(const float* __restrict__ img, unsigned i00, unsigned i01,
                                         unsigned i10, unsigned i11, float t, float u, float v,
                                         float omt, float omu, float omv) {
  const float* p00 = img + i00;
  const float* p01 = img + i01;
  const float* p10 = img + i10;
  const float* p11 = img + i11;
  float v0 = __ldg(p00), v4 = __ldg(p00 + 1);
  float v3 = __ldg(p01), v7 = __ldg(p01 + 1);
  float v1 = __ldg(p10), v5 = __ldg(p10 + 1);
  float v2 = __ldg(p11), v6 = __ldg(p11 + 1);
  return omv * (omu * (omt * v0 + t * v1) + u * (omt * v3 + t * v2)) +
         v * (omu * (omt * v4 + t * v5) + u * (omt * v7 + t * v6));
}

// Along the contiguous axis the two corners are fetched as (zs, zs+1) with zs <= Z-2 so that the
// pair never leaves the row. Where the reference clamps both corners onto one border voxel
// ((1-v)*B + v*B) the weight is replaced by 0 (lower border) or 1 (upper border): the same value
// up to the rounding of (1-v)*B + v*B, i.e. <= 1 ulp, and only outside the volume.
__device__ __forceinline__ void z_pair(const Ax3& az, int Z, int& zs, float& v) {
  zs = min(az.i0, Z - 2);
  v = az.t;
  if (az.i1 == az.i0) v = (az.i0 == 0) ? 0.f : 1.f;
}

// MODE 0: Ad_star (a = phiinv, b = m0); MODE 1: compose (a = u, b = v).
// blockDim = (32, 8): a warp walks one z row (lane = z, NV chunks of 32), a CTA covers 8 y rows.
template <int MODE, int NV, int BX>
__global__ void __launch_bounds__(256)
gather3_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
               int X, int Y, int Z, float dh, float dl, float dsr, float dtr) {
  // blockDim = (32, 8/BX, BX): BX neighbouring x slabs share a CTA so that the upper-x corner rows
  // of one slab are the lower-x rows of the next (L1 reuse instead of a second L2 fetch)
  const int j = blockIdx.y * (8 / BX) + threadIdx.y;
  const int XB = (X + BX - 1) / BX;
  const int i = (blockIdx.z % XB) * BX + threadIdx.z;
  if (j >= Y || i >= X) return;
  const int n = blockIdx.z / XB;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const float* an = a + (size_t)n * 3 * V;
  const float* bn = b + (size_t)n * 3 * V;
  const float* bn1 = bn + V;
  const float* bn2 = bn1 + V;
  float* on = out + (size_t)n * 3 * V;
  // keep the per-subject channel bases as plain 64-bit registers so that every gather address is
  // one IMAD.WIDE.U32 (base + 4*index) instead of a 64-bit add chain
  asm volatile("" : "+l"(bn), "+l"(bn1), "+l"(bn2));
  const int row = i * sx + j * sy;
  const float fi = (float)i, fj = (float)j;
  const int xm = (i > 0) ? -sx : 0, xp = (i < X - 1) ? sx : 0;
  const int ym = (j > 0) ? -sy : 0, yp = (j < Y - 1) ? sy : 0;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = (blockIdx.x * NV + v) * 32 + threadIdx.x;
    if (k >= Z) break;
    const int c0 = row + k;
    const float* a0 = an + c0;
    const float* a1 = a0 + V;
    const float* a2 = a1 + V;
    const float A0 = __ldg(a0), A1 = __ldg(a1), A2 = __ldg(a2);
    float hx, hy, hz;
    const float fk = (float)k;
    if (MODE == 0) {  // dt == 1: the double sum is exact before rounding
      hx = __fadd_rn(fi, A0);
      hy = __fadd_rn(fj, A1);
      hz = __fadd_rn(fk, A2);
    } else {
      hx = coord_f32(fi, A0, dh, dl);
      hy = coord_f32(fj, A1, dh, dl);
      hz = coord_f32(fk, A2, dh, dl);
    }
    const Ax3 ax = axis_fast(hx, X), ay = axis_fast(hy, Y), az = axis_fast(hz, Z);
    int zs;
    float wv;
    z_pair(az, Z, zs, wv);
    const unsigned rx0 = ax.i0 * sx + zs, rx1 = ax.i1 * sx + zs;
    const unsigned ry0 = ay.i0 * sy, ry1 = ay.i1 * sy;
    const unsigned i00 = rx0 + ry0, i01 = rx0 + ry1, i10 = rx1 + ry0, i11 = rx1 + ry1;
    const float omt = 1.f - ax.t, omu = 1.f - ay.t, omv = 1.f - wv;
    const float m0v = trilerp(bn, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv);
    const float m1v = trilerp(bn1, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv);
    const float m2v = trilerp(bn2, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv);
    if (MODE == 1) {
      on[c0] = __fadd_rn(__fmul_rn(dsr, A0), __fmul_rn(dtr, m0v));
      on[c0 + V] = __fadd_rn(__fmul_rn(dsr, A1), __fmul_rn(dtr, m1v));
      on[c0 + 2 * V] = __fadd_rn(__fmul_rn(dsr, A2), __fmul_rn(dtr, m2v));
    } else {
      const int zm = (k > 0) ? -1 : 0, zp = (k < Z - 1) ? 1 : 0;
      const float* ac[3] = {a0, a1, a2};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float g0 = 0.5f * (__ldg(ac[c] + xp) - __ldg(ac[c] + xm));
        float g1 = 0.5f * (__ldg(ac[c] + yp) - __ldg(ac[c] + ym));
        float g2 = 0.5f * (__ldg(ac[c] + zp) - __ldg(ac[c] + zm));
        if (c == 0) g0 += 1.f;
        if (c == 1) g1 += 1.f;
        if (c == 2) g2 += 1.f;
        on[c0 + c * V] = g0 * m0v + g1 * m1v + g2 * m2v;  // diff.cu:118-120
      }
    }
  }
}

static bool fast3_ok(const void* p0, const void* p1, const void* p2, int64_t N, const int64_t* sh) {
  if (((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2) & 15) return false;
  if (sh[0] < 2 || sh[1] < 2 || sh[2] < 2) return false;
  if (sh[0] * sh[1] * sh[2] >= (1LL << 31) / 4) return false;  // 32-bit offsets incl. channel stride
  if (N * sh[0] > 65535 || sh[1] > 65535LL) return false;
  return true;
}

// returns LGM_EUNSUP when the fast path does not apply (caller falls back to the generic kernel)
int Ad_star3_f32(void* out, const void* phi, const void* m, int64_t N, const int64_t* sh, cudaStream_t s) {
  if (!fast3_ok(out, phi, m, N, sh)) return LGM_EUNSUP;
  constexpr int BX = LGM_GATHER_BX;
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8 / BX), (unsigned)(N * cdiv(sh[0], BX))), block(32, 8 / BX, BX);
  gather3_kernel<0, 4, BX><<<grid, block, 0, s>>>((float*)out, (const float*)phi, (const float*)m, (int)sh[0],
                                           (int)sh[1], (int)sh[2], 1.f, 0.f, 0.f, 0.f);
  count_launch("Ad_star", s);
  return finish(s, "lgm_Ad_star_fwd");
}

int compose3_f32(void* out, const void* u, const void* v, int64_t N, const int64_t* sh, double ds, double dt,
                 cudaStream_t s) {
  if (!fast3_ok(out, u, v, N, sh)) return LGM_EUNSUP;
  const float dh = (float)ds, dl = (float)(ds - (double)dh);
  constexpr int BX = LGM_GATHER_BX;
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8 / BX), (unsigned)(N * cdiv(sh[0], BX))), block(32, 8 / BX, BX);
  gather3_kernel<1, 4, BX><<<grid, block, 0, s>>>((float*)out, (const float*)u, (const float*)v, (int)sh[0],
                                           (int)sh[1], (int)sh[2], dh, dl, (float)ds, (float)dt);
  count_launch("compose", s);
  return finish(s, "lgm_compose_fwd");
}

}  // namespace lgm
