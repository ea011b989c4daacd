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

__device__ __forceinline__ Ax3 axis_fast(float x, int n) {
  Ax3 a;
  int f;
  if (fabsf(x) < 4194304.f) {
    float r = __fadd_rn(x, 12582912.f);  // 1.5 * 2^23: rounds x to an integer in the mantissa
    f = __float_as_int(r) - 0x4B400000;
    float rf = __fsub_rn(r, 12582912.f);
    if (rf > x) {
      rf -= 1.f;
      f -= 1;
    }
    a.t = x - rf;
  } else {
    a.t = x - floorf(x);
    f = __float2int_rd(x);
    if (f == 0x7fffffff) f = 0x7ffffffe;
  }
  a.i0 = min(max(f, 0), n - 1);
  a.i1 = min(max(f + 1, 0), n - 1);
  return a;
}

__device__ __forceinline__ float trilerp(const float* __restrict__ img, int o00, int o01, int o10,
                                         int o11, const Ax3& az, float t, float u, float v) {
  // corner numbering / evaluation order of include/interp.h:91-122
  float v0 = __ldg(img + o00 + az.i0), v4 = __ldg(img + o00 + az.i1);
  float v3 = __ldg(img + o01 + az.i0), v7 = __ldg(img + o01 + az.i1);
  float v1 = __ldg(img + o10 + az.i0), v5 = __ldg(img + o10 + az.i1);
  float v2 = __ldg(img + o11 + az.i0), v6 = __ldg(img + o11 + az.i1);
  float omt = 1.f - t, omu = 1.f - u, omv = 1.f - v;
  return omv * (omu * (omt * v0 + t * v1) + u * (omt * v3 + t * v2)) +
         v * (omu * (omt * v4 + t * v5) + u * (omt * v7 + t * v6));
}

__device__ __forceinline__ float f4get(const float4& v, int i) {
  return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}

// MODE 0: Ad_star (a = phiinv, b = m0); MODE 1: compose (a = u, b = v)
template <int MODE>
__global__ void __launch_bounds__(256)
gather3_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
               int X, int Y, int Z, float dh, float dl, float dsr, float dtr) {
  const int k0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (k0 >= Z || j >= Y) return;
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const size_t V = (size_t)X * sx;
  const float* an = a + (size_t)n * 3 * V;
  const float* bn = b + (size_t)n * 3 * V;
  float* on = out + (size_t)n * 3 * V;
  const int c0 = i * sx + j * sy + k0;

  float4 A[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) A[c] = __ldg(reinterpret_cast<const float4*>(an + c * V + c0));

  float val[3][4];
  const float fi = (float)i, fj = (float)j;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    float hx, hy, hz;
    const float fk = (float)(k0 + v);
    if (MODE == 0) {  // dt == 1: the double sum is exact before rounding
      hx = __fadd_rn(fi, f4get(A[0], v));
      hy = __fadd_rn(fj, f4get(A[1], v));
      hz = __fadd_rn(fk, f4get(A[2], v));
    } else {
      hx = coord_f32(fi, f4get(A[0], v), dh, dl);
      hy = coord_f32(fj, f4get(A[1], v), dh, dl);
      hz = coord_f32(fk, f4get(A[2], v), dh, dl);
    }
    const Ax3 ax = axis_fast(hx, X), ay = axis_fast(hy, Y), az = axis_fast(hz, Z);
    const int o00 = ax.i0 * sx + ay.i0 * sy, o01 = ax.i0 * sx + ay.i1 * sy;
    const int o10 = ax.i1 * sx + ay.i0 * sy, o11 = ax.i1 * sx + ay.i1 * sy;
#pragma unroll
    for (int c = 0; c < 3; ++c) val[c][v] = trilerp(bn + c * V, o00, o01, o10, o11, az, ax.t, ay.t, az.t);
  }

  if (MODE == 1) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4 o;
      o.x = __fadd_rn(__fmul_rn(dsr, A[c].x), __fmul_rn(dtr, val[c][0]));
      o.y = __fadd_rn(__fmul_rn(dsr, A[c].y), __fmul_rn(dtr, val[c][1]));
      o.z = __fadd_rn(__fmul_rn(dsr, A[c].z), __fmul_rn(dtr, val[c][2]));
      o.w = __fadd_rn(__fmul_rn(dsr, A[c].w), __fmul_rn(dtr, val[c][3]));
      *reinterpret_cast<float4*>(on + c * V + c0) = o;
    }
  } else {
    const int xm = (i > 0) ? -sx : 0, xp = (i < X - 1) ? sx : 0;
    const int ym = (j > 0) ? -sy : 0, yp = (j < Y - 1) ? sy : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* ac = an + c * V + c0;
      const float4 XM = __ldg(reinterpret_cast<const float4*>(ac + xm));
      const float4 XP = __ldg(reinterpret_cast<const float4*>(ac + xp));
      const float4 YM = __ldg(reinterpret_cast<const float4*>(ac + ym));
      const float4 YP = __ldg(reinterpret_cast<const float4*>(ac + yp));
      const float zl = (k0 > 0) ? __ldg(ac - 1) : A[c].x;
      const float zr = (k0 + 4 < Z) ? __ldg(ac + 4) : A[c].w;
      float gx[4] = {0.5f * (XP.x - XM.x), 0.5f * (XP.y - XM.y), 0.5f * (XP.z - XM.z), 0.5f * (XP.w - XM.w)};
      float gy[4] = {0.5f * (YP.x - YM.x), 0.5f * (YP.y - YM.y), 0.5f * (YP.z - YM.z), 0.5f * (YP.w - YM.w)};
      float gz[4] = {0.5f * (A[c].y - zl), 0.5f * (A[c].z - A[c].x), 0.5f * (A[c].w - A[c].y), 0.5f * (zr - A[c].z)};
      float o[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float g0 = gx[v], g1 = gy[v], g2 = gz[v];
        if (c == 0) g0 += 1.f;
        if (c == 1) g1 += 1.f;
        if (c == 2) g2 += 1.f;
        o[v] = g0 * val[0][v] + g1 * val[1][v] + g2 * val[2][v];  // diff.cu:118-120
      }
      *reinterpret_cast<float4*>(on + c * V + c0) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

static bool fast3_ok(const void* p0, const void* p1, const void* p2, int64_t N, const int64_t* sh) {
  if (((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2) & 15) return false;
  if (sh[2] % 4 != 0 || sh[0] < 2 || sh[1] < 2 || sh[2] < 4) return false;
  if (sh[0] * sh[1] * sh[2] >= (1LL << 31) / 4) return false;  // 32-bit offsets incl. channel stride
  if (N * sh[0] > 65535 || sh[1] > 8 * 65535LL) return false;
  return true;
}

// returns LGM_EUNSUP when the fast path does not apply (caller falls back to the generic kernel)
int Ad_star3_f32(void* out, const void* phi, const void* m, int64_t N, const int64_t* sh, cudaStream_t s) {
  if (!fast3_ok(out, phi, m, N, sh)) return LGM_EUNSUP;
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0])), block(32, 8);
  gather3_kernel<0><<<grid, block, 0, s>>>((float*)out, (const float*)phi, (const float*)m, (int)sh[0],
                                           (int)sh[1], (int)sh[2], 1.f, 0.f, 0.f, 0.f);
  count_launch("Ad_star", s);
  return finish(s, "lgm_Ad_star_fwd");
}

int compose3_f32(void* out, const void* u, const void* v, int64_t N, const int64_t* sh, double ds, double dt,
                 cudaStream_t s) {
  if (!fast3_ok(out, u, v, N, sh)) return LGM_EUNSUP;
  const float dh = (float)ds, dl = (float)(ds - (double)dh);
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0])), block(32, 8);
  gather3_kernel<1><<<grid, block, 0, s>>>((float*)out, (const float*)u, (const float*)v, (int)sh[0],
                                           (int)sh[1], (int)sh[2], dh, dl, (float)ds, (float)dt);
  count_launch("compose", s);
  return finish(s, "lgm_compose_fwd");
}

}  // namespace lgm
