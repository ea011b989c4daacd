// gather3.cu -- fp32 3-D fast paths of the two gather kernels of an EPDiff step:
//   Ad_star : out_c = sum_d (D_d phi_c + delta_cd) * m_d(x + phi(x))          (adjrep.py:86-97)
//   compose : out_c = ds*u_c + dt*v_c(x + ds*u(x))                            (deform.py:53-55)
//
// Layout of the work: lane = z. A warp walks one 128-voxel z row in 4 chunks of 32, a CTA covers 8
// neighbouring y rows of one x slab, the batch and x are the grid's z dimension, so that
//   - every direct load / store of a warp is one aligned 128-byte row segment;
//   - the 8-corner gathers of a warp land on runs of consecutive addresses for smooth flows (two
//     cache lines per request); the upper-z corner is the +1 neighbour of the lower one, i.e. an
//     immediate offset on the same address register;
//   - all index arithmetic is 32-bit, one IMAD.WIDE per corner row, no division in the hot loop.
// (A float4-per-thread mapping was measured and rejected: it quadruples the L1 wavefronts of every
// gather request.) Sample coordinates reproduce the reference's "form in double, round to float"
// (cuda/interp.cu:68-73) with error-free float transformations instead of fp64/conversion
// instructions; floor() is a magic-number add.
#include "gather_common.cuh"

#ifndef LGM_ADSTAR_MINB
#define LGM_ADSTAR_MINB 4  /* CTAs/SM the Ad_star instantiation is compiled for (64 registers) */
#endif
// cache hints (kernel experiments): 1 = streaming stores, 2 = + streaming centre loads
#ifndef LGM_GATHER_HINTS
#define LGM_GATHER_HINTS 0
#endif
#if LGM_GATHER_HINTS >= 1
#define LGM_ST(p, v) __stcs((p), (v))
#else
#define LGM_ST(p, v) (*(p) = (v))
#endif
#if LGM_GATHER_HINTS >= 2
#define LGM_LDC(p) __ldcs(p)
#else
#define LGM_LDC(p) __ldg(p)
#endif
#ifndef LGM_GATHER_NR
#define LGM_GATHER_NR 1  /* y rows per thread in Ad_star / compose */
#endif
#ifndef LGM_GATHER_NV
#define LGM_GATHER_NV 4  /* 32-voxel chunks of a z row per thread in Ad_star / compose */
#endif
#ifndef LGM_COMPOSE_MINB
#define LGM_COMPOSE_MINB 5
#endif
#ifndef LGM_GATHER_BX
#define LGM_GATHER_BX 1  /* 2,4,8 measured equal on B200: the gathers are L1-data-pipe bound, not L2 bound */
#endif

namespace lgm {

// MODE 0: Ad_star (a = phiinv, b = m0); MODE 1: compose (a = u, b = v).
// blockDim = (32, 8): a warp walks one z row (lane = z, NV chunks of 32), a CTA covers 8 y rows.
template <int MODE, int NV, int BX, int NR>
__global__ void __launch_bounds__(256, MODE == 0 ? LGM_ADSTAR_MINB : LGM_COMPOSE_MINB)
gather3_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
               int X, int Y, int Z, float dh, float dl, float dsr, float dtr, int rev) {
  // blockDim = (32, 8/BX, BX): BX neighbouring x slabs share a CTA so that the upper-x corner rows
  // of one slab are the lower-x rows of the next (L1 reuse instead of a second L2 fetch)
  // rev: walk the grid from its far end (the EPDiff drivers alternate it from kernel to kernel)
  const unsigned bz = rev ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
  const unsigned by = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  if (BX == 1 && NR == 1) LGM_PREFETCH_ROWS_AHEAD(6, (threadIdx.x < 3 ? a : b), X, Y, Z, rev)
  // a thread owns NR rows (j, j + 8/BX, ...) x NV chunks: NR * NV voxels, all centre loads up front
  const int j = by * (8 / BX) * NR + threadIdx.y;
  const int XB = (X + BX - 1) / BX;
  const int i = (bz % XB) * BX + threadIdx.z;
  if (j >= Y || i >= X) return;
  const int n = bz / XB;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const float* an = a + (size_t)n * 3 * V;
  const float* bn = b + (size_t)n * 3 * V;
  const float* bn1 = bn + V;
  const float* bn2 = bn1 + V;
  float* on = out + (size_t)n * 3 * V;
  float* on1 = on + V;
  float* on2 = on1 + V;
  const float* an1 = an + V;
  const float* an2 = an1 + V;
  // keep the per-subject channel bases as plain 64-bit registers so that every address (gather,
  // stencil neighbour, store) is ONE IMAD.WIDE (base + 4*index) on a 32-bit in-volume index instead
  // of a 64-bit add / LEA chain (the stencil loads alone were 4 address instructions each)
  asm volatile("" : "+l"(bn), "+l"(bn1), "+l"(bn2));
  asm volatile("" : "+l"(an), "+l"(an1), "+l"(an2));
  asm volatile("" : "+l"(on), "+l"(on1), "+l"(on2));
  const unsigned four = opaque_four();
  const float hiX = (float)X - 0.5f, hiY = (float)Y - 0.5f, hiZ = (float)Z - 0.5f;
  const int row0 = i * sx + j * sy;
  const float fi = (float)i;
  const int xm = (i > 0) ? -sx : 0, xp = (i < X - 1) ? sx : 0;
  // the centre loads of ALL chunks are issued up front: they gate everything else of a chunk, and
  // 3*NV registers buy one full memory round trip of overlap per chunk
  float Apre[NR * NV][3];
#pragma unroll
  for (int sl = 0; sl < NR * NV; ++sl) {
    const int r = sl / NV, v = sl % NV;
    const int k = (blockIdx.x * NV + v) * 32 + threadIdx.x;
    const int row = row0 + r * (8 / BX) * sy;
    if (k < Z && j + r * (8 / BX) < Y) {
      Apre[sl][0] = LGM_LDC(an + (row + k));  // literal addressing: must not wait for `four`
      Apre[sl][1] = LGM_LDC(an1 + (row + k));
      Apre[sl][2] = LGM_LDC(an2 + (row + k));
    }
  }
#pragma unroll
  for (int sl = 0; sl < NR * NV; ++sl) {
    const int r = sl / NV, v = sl % NV;
    const int k = (blockIdx.x * NV + v) * 32 + threadIdx.x;
    const int jr = j + r * (8 / BX);
    if (k >= Z || jr >= Y) continue;
    const int row = row0 + r * (8 / BX) * sy;
    const float fj = (float)jr;
    const int ym = (jr > 0) ? -sy : 0, yp = (jr < Y - 1) ? sy : 0;
    const int c0 = row + k;
    const float A0 = Apre[sl][0], A1 = Apre[sl][1], A2 = Apre[sl][2];
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
    const Ax3 ax = axis_fwd(hx, X, hiX), ay = axis_fwd(hy, Y, hiY), az = axis_fwd(hz, Z, hiZ);
    int zs;
    float wv;
    z_pair(az, Z, zs, wv);
    const unsigned rx0 = ax.i0 * sx + zs, rx1 = ax.i1 * sx + zs;
    const unsigned ry0 = ay.i0 * sy, ry1 = ay.i1 * sy;
    const unsigned i00 = rx0 + ry0, i01 = rx0 + ry1, i10 = rx1 + ry0, i11 = rx1 + ry1;
    const float omt = 1.f - ax.t, omu = 1.f - ay.t, omv = 1.f - wv;
    const float m0v = trilerp(bn, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv, four);
    const float m1v = trilerp(bn1, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv, four);
    const float m2v = trilerp(bn2, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv, four);
    if (MODE == 1) {
      LGM_ST(at4(on, c0, four), __fadd_rn(__fmul_rn(dsr, A0), __fmul_rn(dtr, m0v)));
      LGM_ST(at4(on1, c0, four), __fadd_rn(__fmul_rn(dsr, A1), __fmul_rn(dtr, m1v)));
      LGM_ST(at4(on2, c0, four), __fadd_rn(__fmul_rn(dsr, A2), __fmul_rn(dtr, m2v)));
    } else {
      const int ixp = c0 + xp, ixm = c0 + xm, iyp = c0 + yp, iym = c0 + ym;
      const int izp = c0 + ((k < Z - 1) ? 1 : 0), izm = c0 - ((k > 0) ? 1 : 0);
      const float* ac[3] = {an, an1, an2};
      float* oc[3] = {on, on1, on2};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float g0 = 0.5f * (__ldg(at4(ac[c], ixp, four)) - __ldg(at4(ac[c], ixm, four)));
        float g1 = 0.5f * (__ldg(at4(ac[c], iyp, four)) - __ldg(at4(ac[c], iym, four)));
        float g2 = 0.5f * (__ldg(at4(ac[c], izp, four)) - __ldg(at4(ac[c], izm, four)));
        if (c == 0) g0 += 1.f;
        if (c == 1) g1 += 1.f;
        if (c == 2) g2 += 1.f;
        LGM_ST(at4(oc[c], c0, four), g0 * m0v + g1 * m1v + g2 * m2v);  // diff.cu:118-120
      }
    }
  }
}

// Plain interp (cuda/interp.cu:47-78), any channel count, optional broadcast image: same thread
// mapping and corner-pair gather as above, weights computed once per voxel for all channels.
// CC: compile-time channel count (0 = runtime C): the common C = 1 / C = 3 cases unroll the channel loop,
// so all 8*C corner loads of a voxel are in flight together.
template <int NV, bool UNIT_DT, int CC>
__global__ void __launch_bounds__(256)
interp3_kernel(float* __restrict__ out, const float* __restrict__ I, const float* __restrict__ u, int X,
               int Y, int Z, int C_rt, size_t I_batch_stride, float dh, float dl) {
  const int C = CC ? CC : C_rt;
  LGM_PREFETCH_ROWS_AHEAD(3, u, X, Y, Z, 0)
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (j >= Y) return;
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const float* un = u + (size_t)n * 3 * V;
  const float* In = I + (size_t)n * I_batch_stride;
  float* on = out + (size_t)n * C * V;
  asm volatile("" : "+l"(In));
  const unsigned four = opaque_four();
  const float hiX = (float)X - 0.5f, hiY = (float)Y - 0.5f, hiZ = (float)Z - 0.5f;
  const int row = i * sx + j * sy;
  const float fi = (float)i, fj = (float)j;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = (blockIdx.x * NV + v) * 32 + threadIdx.x;
    if (k >= Z) break;
    const int c0 = row + k;
    const float A0 = __ldg(un + c0), A1 = __ldg(un + c0 + V), A2 = __ldg(un + c0 + 2 * V);
    float hx, hy, hz;
    const float fk = (float)k;
    if (UNIT_DT) {
      hx = __fadd_rn(fi, A0);
      hy = __fadd_rn(fj, A1);
      hz = __fadd_rn(fk, A2);
    } else {
      hx = coord_f32(fi, A0, dh, dl);
      hy = coord_f32(fj, A1, dh, dl);
      hz = coord_f32(fk, A2, dh, dl);
    }
    const Ax3 ax = axis_fwd(hx, X, hiX), ay = axis_fwd(hy, Y, hiY), az = axis_fwd(hz, Z, hiZ);
    int zs;
    float wv;
    z_pair(az, Z, zs, wv);
    const unsigned rx0 = ax.i0 * sx + zs, rx1 = ax.i1 * sx + zs;
    const unsigned ry0 = ay.i0 * sy, ry1 = ay.i1 * sy;
    const unsigned i00 = rx0 + ry0, i01 = rx0 + ry1, i10 = rx1 + ry0, i11 = rx1 + ry1;
    const float omt = 1.f - ax.t, omu = 1.f - ay.t, omv = 1.f - wv;
    if constexpr (CC > 0) {
      float r[CC];
#pragma unroll
      for (int c = 0; c < CC; ++c) r[c] = trilerp(In + c * V, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv, four);
#pragma unroll
      for (int c = 0; c < CC; ++c) on[c0 + c * V] = r[c];
    } else {
      const float* Ic = In;
      float* oc = on + c0;
      for (int c = 0; c < C; ++c, Ic += V, oc += V)
        *oc = trilerp(Ic, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv, four);
    }
  }
}

// d_u of interp (cuda/interp.cu:224-233): d_u[d] = sum_c (gout_c * dt) * d/dx_d I_c(h), with the
// corner-difference gradient of include/interp.h:315-326. Same mapping / gathers as interp3_kernel.
template <int NV, bool UNIT_DT, int CC>
__global__ void __launch_bounds__(256)
interp_du3_kernel(float* __restrict__ d_u, const float* __restrict__ go, const float* __restrict__ I,
                  const float* __restrict__ u, int X, int Y, int Z, int C_rt, size_t I_batch_stride, float dh,
                  float dl, double dt) {
  const int C = CC ? CC : C_rt;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (j >= Y) return;
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const float* un = u + (size_t)n * 3 * V;
  const float* gn = go + (size_t)n * C * V;
  const float* In = I + (size_t)n * I_batch_stride;
  float* dn = d_u + (size_t)n * 3 * V;
  asm volatile("" : "+l"(In));
  const int row = i * sx + j * sy;
  const float fi = (float)i, fj = (float)j;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = (blockIdx.x * NV + v) * 32 + threadIdx.x;
    if (k >= Z) break;
    const int c0 = row + k;
    const float A0 = __ldg(un + c0), A1 = __ldg(un + c0 + V), A2 = __ldg(un + c0 + 2 * V);
    float hx, hy, hz;
    const float fk = (float)k;
    if (UNIT_DT) {
      hx = __fadd_rn(fi, A0);
      hy = __fadd_rn(fj, A1);
      hz = __fadd_rn(fk, A2);
    } else {
      hx = coord_f32(fi, A0, dh, dl);
      hy = coord_f32(fj, A1, dh, dl);
      hz = coord_f32(fk, A2, dh, dl);
    }
    const Ax3 ax = axis_fast(hx, X), ay = axis_fast(hy, Y), az = axis_fast(hz, Z);
    // the gradient needs the true (possibly coincident) corner values, so use the clamped pair
    const unsigned rx0 = ax.i0 * sx, rx1 = ax.i1 * sx;
    const unsigned ry0 = ay.i0 * sy, ry1 = ay.i1 * sy;
    const unsigned i00 = rx0 + ry0 + az.i0, i01 = rx0 + ry1 + az.i0, i10 = rx1 + ry0 + az.i0, i11 = rx1 + ry1 + az.i0;
    const int dz = az.i1 - az.i0;
    const float t = ax.t, uu = ay.t, w = az.t;
    const float omt = 1.f - t, omu = 1.f - uu, omv = 1.f - w;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    const float* Ic = In;
#pragma unroll
    for (int c = 0; c < C; ++c, Ic += V) {
      const float g = __ldg(gn + (size_t)c * V + c0);
      const float gd = (float)((double)g * dt);  // "diff *= dt" in double, cuda/interp.cu:230
      const float v0 = __ldg(Ic + i00), v4 = __ldg(Ic + i00 + dz);
      const float v3 = __ldg(Ic + i01), v7 = __ldg(Ic + i01 + dz);
      const float v1 = __ldg(Ic + i10), v5 = __ldg(Ic + i10 + dz);
      const float v2 = __ldg(Ic + i11), v6 = __ldg(Ic + i11 + dz);
      const float gx = omv * (omu * (v1 - v0) + uu * (v2 - v3)) + w * (omu * (v5 - v4) + uu * (v6 - v7));
      const float gy = omv * (omt * (v3 - v0) + t * (v2 - v1)) + w * (omt * (v7 - v4) + t * (v6 - v5));
      const float gz = omu * (omt * (v4 - v0) + t * (v5 - v1)) + uu * (omt * (v7 - v3) + t * (v6 - v2));
      a0 = a0 + gx * gd;
      a1 = a1 + gy * gd;
      a2 = a2 + gz * gd;
    }
    dn[c0] = a0;
    dn[c0 + V] = a1;
    dn[c0 + 2 * V] = a2;
  }
}

// Adjoint of interp (splat, cuda/interp.cu:185-244 d_I part): every voxel adds w_corner * gout to
// its 8 corner voxels. Lanes are consecutive in z, so for smooth flows the upper-z corner of lane L
// is the lower-z corner of lane L+1: those two contributions are merged with one warp shuffle and
// leave as ONE red.global.add, which halves the L2 atomic traffic (the limiter of this kernel).
// Corner weights follow the reference's alternating "d = 1 - d" sequence (include/interp.h:437-453).
template <int NV, bool UNIT_DT, int CC>
__global__ void __launch_bounds__(256)
splat3_kernel(float* __restrict__ d_I, const float* __restrict__ go, const float* __restrict__ u, int X,
              int Y, int Z, int C_rt, size_t I_batch_stride, float dh, float dl) {
  const int C = CC ? CC : C_rt;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (j >= Y) return;  // warp-uniform (a warp is one row)
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const float* un = u + (size_t)n * 3 * V;
  const float* gn = go + (size_t)n * C * V;
  float* dn = d_I + (size_t)n * I_batch_stride;
  const int row = i * sx + j * sy;
  const float fi = (float)i, fj = (float)j;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x;
#pragma unroll 1
  for (int v = 0; v < NV; ++v) {
    const int k = (blockIdx.x * NV + v) * 32 + lane;
    if ((blockIdx.x * NV + v) * 32 >= Z) break;  // whole chunk out of range (Z % 32 == 0)
    const int c0 = row + k;
    const float A0 = __ldg(un + c0), A1 = __ldg(un + c0 + V), A2 = __ldg(un + c0 + 2 * V);
    float hx, hy, hz;
    const float fk = (float)k;
    if (UNIT_DT) {
      hx = __fadd_rn(fi, A0);
      hy = __fadd_rn(fj, A1);
      hz = __fadd_rn(fk, A2);
    } else {
      hx = coord_f32(fi, A0, dh, dl);
      hy = coord_f32(fj, A1, dh, dl);
      hz = coord_f32(fk, A2, dh, dl);
    }
    const Ax3 ax = axis_fast(hx, X), ay = axis_fast(hy, Y), az = axis_fast(hz, Z);
    // weight sequences: w0 = 1-t, then alternately 1-w
    const float wx0 = 1.f - ax.t, wx1 = 1.f - wx0;
    const float wy0 = 1.f - ay.t, wy1 = 1.f - wy0, wy2 = 1.f - wy1, wy3 = 1.f - wy2;
    const float wz0 = 1.f - az.t, wz1 = 1.f - wz0, wz2 = 1.f - wz1, wz3 = 1.f - wz2;
    const float wr[4] = {wx0 * wy0, wx0 * wy1, wx1 * wy2, wx1 * wy3};  // rows (x0,y0),(x0,y1),(x1,y0),(x1,y1)
    const float wlo[4] = {wr[0] * wz0, wr[1] * wz2, wr[2] * wz2, wr[3] * wz2};
    const float whi[4] = {wr[0] * wz1, wr[1] * wz3, wr[2] * wz3, wr[3] * wz3};
    const unsigned rb[4] = {(unsigned)(ax.i0 * sx + ay.i0 * sy), (unsigned)(ax.i0 * sx + ay.i1 * sy),
                            (unsigned)(ax.i1 * sx + ay.i0 * sy), (unsigned)(ax.i1 * sx + ay.i1 * sy)};
    bool give[4], took[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const unsigned alo = rb[r] + az.i0, ahi = rb[r] + az.i1;
      const unsigned nxt = __shfl_down_sync(full, alo, 1);
      give[r] = (lane < 31) && (nxt == ahi) && (ahi != alo);
      took[r] = __shfl_up_sync(full, (int)give[r], 1) != 0 && lane > 0;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float d = __ldg(gn + (size_t)c * V + c0);
      float* dc = dn + (size_t)c * V;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float vlo = wlo[r] * d;
        const float vhi = whi[r] * d;
        const float recv = __shfl_up_sync(full, give[r] ? vhi : 0.f, 1);
        if (took[r]) vlo += recv;
        atomicAdd(dc + rb[r] + az.i0, vlo);
        if (!give[r]) atomicAdd(dc + rb[r] + az.i1, vhi);
      }
    }
  }
}

static bool fast3_ok(const void* p0, const void* p1, const void* p2, int64_t N, const int64_t* sh) {
  (void)p0; (void)p1; (void)p2;  // scalar accesses: no alignment requirement beyond the element
  if (sh[0] < 2 || sh[1] < 2 || sh[2] < 2) return false;
  if (sh[0] * sh[1] * sh[2] >= (1LL << 31) / 4) return false;  // 32-bit offsets incl. channel stride
  if (N * sh[0] > 65535 || sh[1] > 65535LL) return false;
  return true;
}

// returns LGM_EUNSUP when the fast path does not apply (caller falls back to the generic kernel)
int Ad_star3_ring_f32(void* out, const void* phi, const void* m, int64_t N, const int64_t* sh, int rev,
                      cudaStream_t s);  // adstar_ring.cu

int Ad_star3_f32(void* out, const void* phi, const void* m, int64_t N, const int64_t* sh, int rev, cudaStream_t s) {
  if (!fast3_ok(out, phi, m, N, sh)) return LGM_EUNSUP;
  {  // power-of-two rows: the stencil operand staged in shared memory (TMA ring)
    const int rc = Ad_star3_ring_f32(out, phi, m, N, sh, rev, s);
    if (rc != LGM_EUNSUP) return rc;
  }
  constexpr int BX = LGM_GATHER_BX;
  dim3 grid((unsigned)cdiv(sh[2], 32 * LGM_GATHER_NV), (unsigned)cdiv(sh[1], (8 / BX) * LGM_GATHER_NR), (unsigned)(N * cdiv(sh[0], BX))), block(32, 8 / BX, BX);
  gather3_kernel<0, LGM_GATHER_NV, BX, LGM_GATHER_NR><<<grid, block, 0, s>>>((float*)out, (const float*)phi, (const float*)m, (int)sh[0],
                                           (int)sh[1], (int)sh[2], 1.f, 0.f, 0.f, 0.f, rev);
  count_launch("Ad_star", s);
  return finish(s, "lgm_Ad_star_fwd");
}

int compose3_ring_f32(void* out, const void* u, const void* v, int64_t N, const int64_t* sh, double ds, double dt,
                      int rev, cudaStream_t s);  // compose_ring.cu

int compose3_f32(void* out, const void* u, const void* v, int64_t N, const int64_t* sh, double ds, double dt,
                 int rev, cudaStream_t s) {
  if (!fast3_ok(out, u, v, N, sh)) return LGM_EUNSUP;
  {  // sub-voxel displacements on power-of-two rows: gather source staged in shared memory (TMA ring)
    const int rc = compose3_ring_f32(out, u, v, N, sh, ds, dt, rev, s);
    if (rc != LGM_EUNSUP) return rc;
  }
  const float dh = (float)ds, dl = (float)(ds - (double)dh);
  constexpr int BX = LGM_GATHER_BX;
  dim3 grid((unsigned)cdiv(sh[2], 32 * LGM_GATHER_NV), (unsigned)cdiv(sh[1], (8 / BX) * LGM_GATHER_NR), (unsigned)(N * cdiv(sh[0], BX))), block(32, 8 / BX, BX);
  gather3_kernel<1, LGM_GATHER_NV, BX, LGM_GATHER_NR><<<grid, block, 0, s>>>((float*)out, (const float*)u, (const float*)v, (int)sh[0],
                                           (int)sh[1], (int)sh[2], dh, dl, (float)ds, (float)dt, rev);
  count_launch("compose", s);
  return finish(s, "lgm_compose_fwd");
}

int interp3_f32(void* out, const void* I, const void* u, int64_t N, int64_t NI, int64_t C, const int64_t* sh,
                double dt, cudaStream_t s) {
  if (!fast3_ok(out, I, u, N, sh) || C < 1 || C > 0x7fffffff / (sh[0] * sh[1] * sh[2])) return LGM_EUNSUP;
  const size_t ibs = (NI < N) ? 0 : (size_t)C * sh[0] * sh[1] * sh[2];
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0])), block(32, 8);
  const float dh = (dt == 1.0) ? 1.f : (float)dt, dl = (dt == 1.0) ? 0.f : (float)(dt - (double)dh);
#define LGM_INTERP3(UNIT, CC)                                                                                   \
  interp3_kernel<4, UNIT, CC><<<grid, block, 0, s>>>((float*)out, (const float*)I, (const float*)u, (int)sh[0], \
                                                     (int)sh[1], (int)sh[2], (int)C, ibs, dh, dl)
  if (dt == 1.0) {
    if (C == 1) LGM_INTERP3(true, 1); else if (C == 3) LGM_INTERP3(true, 3); else LGM_INTERP3(true, 0);
  } else {
    if (C == 1) LGM_INTERP3(false, 1); else if (C == 3) LGM_INTERP3(false, 3); else LGM_INTERP3(false, 0);
  }
#undef LGM_INTERP3
  count_launch("interp_fwd", s);
  return finish(s, "lgm_interp_fwd");
}

// d_I must be zero-filled by the caller. LGM_EUNSUP when the fast path does not apply.
int splat3_f32(void* d_I, const void* go, const void* u, int64_t N, int64_t NI, int64_t C, const int64_t* sh,
               double dt, cudaStream_t s) {
  if (!fast3_ok(d_I, go, u, N, sh) || sh[2] % 32 != 0 || C < 1 || C > 0x7fffffff / (sh[0] * sh[1] * sh[2]))
    return LGM_EUNSUP;
  const size_t ibs = (NI < N) ? 0 : (size_t)C * sh[0] * sh[1] * sh[2];
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0])), block(32, 8);
  const float dh = (dt == 1.0) ? 1.f : (float)dt, dl = (dt == 1.0) ? 0.f : (float)(dt - (double)dh);
#define LGM_SPLAT3(UNIT, CC)                                                                                    \
  splat3_kernel<4, UNIT, CC><<<grid, block, 0, s>>>((float*)d_I, (const float*)go, (const float*)u, (int)sh[0], \
                                                    (int)sh[1], (int)sh[2], (int)C, ibs, dh, dl)
  if (dt == 1.0) {
    if (C == 1) LGM_SPLAT3(true, 1); else if (C == 3) LGM_SPLAT3(true, 3); else LGM_SPLAT3(true, 0);
  } else {
    if (C == 1) LGM_SPLAT3(false, 1); else if (C == 3) LGM_SPLAT3(false, 3); else LGM_SPLAT3(false, 0);
  }
#undef LGM_SPLAT3
  count_launch("interp_splat", s);
  return finish(s, "lgm_interp_bwd");
}

int interp_du3_f32(void* d_u, const void* go, const void* I, const void* u, int64_t N, int64_t NI, int64_t C,
                   const int64_t* sh, double dt, cudaStream_t s) {
  if (!fast3_ok(d_u, go, u, N, sh) || C < 1 || C > 0x7fffffff / (sh[0] * sh[1] * sh[2])) return LGM_EUNSUP;
  const size_t ibs = (NI < N) ? 0 : (size_t)C * sh[0] * sh[1] * sh[2];
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0])), block(32, 8);
  const float dh = (dt == 1.0) ? 1.f : (float)dt, dl = (dt == 1.0) ? 0.f : (float)(dt - (double)dh);
#define LGM_DU3(UNIT, CC)                                                                                                 \
  interp_du3_kernel<4, UNIT, CC><<<grid, block, 0, s>>>((float*)d_u, (const float*)go, (const float*)I, (const float*)u, \
                                                        (int)sh[0], (int)sh[1], (int)sh[2], (int)C, ibs, dh, dl, dt)
  if (dt == 1.0) {
    if (C == 1) LGM_DU3(true, 1); else if (C == 3) LGM_DU3(true, 3); else LGM_DU3(true, 0);
  } else {
    if (C == 1) LGM_DU3(false, 1); else if (C == 3) LGM_DU3(false, 3); else LGM_DU3(false, 0);
  }
#undef LGM_DU3
  count_launch("interp_du", s);
  return finish(s, "lgm_interp_bwd");
}

}  // namespace lgm
