// epdiff_bwd.cu -- backward of one EPDiff step (lagomorph/lddmm.py:39-44) as three fused kernels
// around one FluidMetric application, fp32 3-D.
//
// Forward step (phi = phiinv displacement, ds = -dt):
//   mi = m0(x + phi)            m_c = sum_d (D_d phi_c + delta_cd) mi_d      [* mommask]
//   v  = sharp(m)               phi' = ds*v + phi(x + ds*v)
// Given G = dL/dphi', the reference's autograd chain (deform.py:31-41 interp backward,
// diff.py:29-35 jtvf backward, metric.py:21-34) runs, per step: a 3-channel splat + an interp d_u
// kernel + 2 pointwise kernels (compose), sharp, the jtvf backward kernel, an interp recompute, a
// second splat + d_u pair and 3 gradient-accumulation adds. Here:
//   compose_bwd3 : one pass over (G, v, phi): d_v = ds*G + ds*sum_c G_c grad phi_c(h) (written) and
//                  splat of G at h = x + ds*v into the accumulator S (the d_phi of this step); the
//                  8 corner addresses / weights are computed once for the gather and the splat.
//   sharp        : d_m = sharp(d_v) in place (fluid.cu) [* mommask]
//   adstar_bwd3  : one pass over (phi, d_m, m0): q = (D phi + I)^T d_m; gathers m0's corners once
//                  for mi (written, needed by the stencil pass), for d_phi += sum_c q_c grad m0_c(h)
//                  (added to S by the owning thread) and splats q into d_m0, which accumulates over
//                  ALL steps of the shoot in place (no per-step gradient adds).
//   stencil_bwd3 : G <- S + sum_d D_d^T (mi_d * d_m_c), S <- 0 (ready for the next step).
// Arithmetic follows the unfused kernels (interp.cu / diff.cu / gather3.cu) term by term; only the
// order in which the three d_phi contributions are summed differs (fp32 rounding, <= 1e-6 rel).
#include "gather_common.cuh"

namespace lgm {

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct Corners {
  unsigned lo[4], hi[4];  // element offsets of the (x0,y0),(x0,y1),(x1,y0),(x1,y1) corner rows at z0 / z1
  float t, u, w;          // fractions along x, y, z (from the unclamped coordinate)
  float wlo[4], whi[4];   // splat weights of the rows at z0 / z1 ("d = 1 - d" sequence, interp.h:437-453)
  bool give, took;        // upper-z contributions handed to lane+1 / received from lane-1
};

template <bool SPLAT>
__device__ __forceinline__ void corner_setup(Corners& cs, float hx, float hy, float hz, int X, int Y, int Z,
                                             int sx, int sy, int lane) {
  const Ax3 ax = axis_fast(hx, X), ay = axis_fast(hy, Y), az = axis_fast(hz, Z);
  cs.t = ax.t; cs.u = ay.t; cs.w = az.t;
  const unsigned x0 = ax.i0 * sx, x1 = ax.i1 * sx, y0 = ay.i0 * sy, y1 = ay.i1 * sy;
  const unsigned r0 = x0 + y0, r1 = x0 + y1, r2 = x1 + y0, r3 = x1 + y1;
  cs.lo[0] = r0 + az.i0; cs.lo[1] = r1 + az.i0; cs.lo[2] = r2 + az.i0; cs.lo[3] = r3 + az.i0;
  cs.hi[0] = r0 + az.i1; cs.hi[1] = r1 + az.i1; cs.hi[2] = r2 + az.i1; cs.hi[3] = r3 + az.i1;
  if (SPLAT) {
    const float wx0 = 1.f - ax.t, wx1 = 1.f - wx0;
    const float wy0 = 1.f - ay.t, wy1 = 1.f - wy0, wy2 = 1.f - wy1, wy3 = 1.f - wy2;
    const float wz0 = 1.f - az.t, wz1 = 1.f - wz0, wz2 = 1.f - wz1, wz3 = 1.f - wz2;
    const float wr[4] = {wx0 * wy0, wx0 * wy1, wx1 * wy2, wx1 * wy3};
    cs.wlo[0] = wr[0] * wz0; cs.wlo[1] = wr[1] * wz2; cs.wlo[2] = wr[2] * wz2; cs.wlo[3] = wr[3] * wz2;
    cs.whi[0] = wr[0] * wz1; cs.whi[1] = wr[1] * wz3; cs.whi[2] = wr[2] * wz3; cs.whi[3] = wr[3] * wz3;
    // lane+1's four lower-z corners are this lane's four upper-z corners iff rows 0 and 3 match
    // (row offset = x*sx + y*sy decomposes uniquely, so rows 1 and 2 follow)
    const unsigned n0 = __shfl_down_sync(kFull, cs.lo[0], 1), n3 = __shfl_down_sync(kFull, cs.lo[3], 1);
    cs.give = (lane < 31) && (n0 == cs.hi[0]) && (n3 == cs.hi[3]) && (az.i1 != az.i0);
    cs.took = (__shfl_up_sync(kFull, (int)cs.give, 1) != 0) && (lane > 0);
  }
}

// the 8 corner values of one channel, numbered as include/interp.h:91-122
struct Vals { float v0, v1, v2, v3, v4, v5, v6, v7; };

// img: per-channel base held in a 64-bit register; every address is one IMAD.WIDE.U32
__device__ __forceinline__ Vals corner_load(const float* img, const Corners& cs) {
  Vals q;
  q.v0 = __ldg(img + cs.lo[0]); q.v4 = __ldg(img + cs.hi[0]);
  q.v3 = __ldg(img + cs.lo[1]); q.v7 = __ldg(img + cs.hi[1]);
  q.v1 = __ldg(img + cs.lo[2]); q.v5 = __ldg(img + cs.hi[2]);
  q.v2 = __ldg(img + cs.lo[3]); q.v6 = __ldg(img + cs.hi[3]);
  return q;
}

__device__ __forceinline__ float corner_value(const Vals& q, const Corners& cs) {
  const float t = cs.t, u = cs.u, w = cs.w, omt = 1.f - t, omu = 1.f - u, omv = 1.f - w;
  return omv * (omu * (omt * q.v0 + t * q.v1) + u * (omt * q.v3 + t * q.v2)) +
         w * (omu * (omt * q.v4 + t * q.v5) + u * (omt * q.v7 + t * q.v6));
}

// corner-difference gradient, include/interp.h:315-326
__device__ __forceinline__ void corner_grad(const Vals& q, const Corners& cs, float& gx, float& gy, float& gz) {
  const float t = cs.t, u = cs.u, w = cs.w, omt = 1.f - t, omu = 1.f - u, omv = 1.f - w;
  gx = omv * (omu * (q.v1 - q.v0) + u * (q.v2 - q.v3)) + w * (omu * (q.v5 - q.v4) + u * (q.v6 - q.v7));
  gy = omv * (omt * (q.v3 - q.v0) + t * (q.v2 - q.v1)) + w * (omt * (q.v7 - q.v4) + t * (q.v6 - q.v5));
  gz = omu * (omt * (q.v4 - q.v0) + t * (q.v5 - q.v1)) + u * (omt * (q.v7 - q.v3) + t * (q.v6 - q.v2));
}

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_if(float* p, float v, bool on) {  // predicated, no branch
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q red.global.add.f32 [%0], %1; }" ::"l"(p), "f"(v),
               "r"((int)on)
               : "memory");
}

// adds w_corner * d to the 8 corners of one channel; the upper-z share of lane L rides to lane L+1
// when that lane's lower-z corners are the same voxels (gather3.cu splat3_kernel). Whole warp calls.
__device__ __forceinline__ void corner_splat(float* dc, const Corners& cs, float d) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float vlo = cs.wlo[r] * d;
    const float vhi = cs.whi[r] * d;
    const float recv = __shfl_up_sync(kFull, vhi, 1);
    if (cs.took) vlo += recv;
    red_add(dc + cs.lo[r], vlo);
    red_add_if(dc + cs.hi[r], vhi, !cs.give);
  }
}

// ---- compose backward: phi' = ds*v + phi(x + ds*v) ------------------------------------------------
// blockDim = (32, 8): lane = z, a warp walks one z row in NV chunks of 32, a CTA covers 8 y rows.
template <int NV, bool NEED_PHI>
__global__ void __launch_bounds__(256)
compose_bwd3_kernel(float* __restrict__ dv, float* __restrict__ S, const float* __restrict__ G,
                    const float* __restrict__ phi, const float* __restrict__ vel, int X, int Y, int Z,
                    float dh, float dl, float dsf) {
  LGM_PREFETCH_ROWS_AHEAD(9, (threadIdx.x < 3 ? vel : (threadIdx.x < 6 ? G : phi)), X, Y, Z, 0)
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (j >= Y) return;  // warp-uniform
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const size_t nb = (size_t)n * 3 * V;
  const float* vn = vel + nb;
  const float* gn = G + nb;
  const float* pn = phi + nb;
  float* dvn = dv + nb;
  // per-channel bases of the gathered / splatted arrays as opaque 64-bit registers
  const float* pb[3] = {pn, pn + V, pn + 2 * (size_t)V};
  float* Sb[3] = {S + nb, S + nb + V, S + nb + 2 * (size_t)V};
  asm volatile("" : "+l"(pb[0]), "+l"(pb[1]), "+l"(pb[2]), "+l"(Sb[0]), "+l"(Sb[1]), "+l"(Sb[2]));
  const int row = i * sx + j * sy;
  const float fi = (float)i, fj = (float)j;
  const int lane = threadIdx.x;
  // Software pipeline over the chunks: the centre loads of chunk c4+1 are issued BEFORE chunk c4 is
  // processed (the REDs are asm volatile with a memory clobber, so the compiler never moves a load
  // across them itself): per chunk one dependent memory round trip (the corner gather) instead of two.
  float nA[3], nG[3];
  {
    const int c0 = row + blockIdx.x * NV * 32 + lane;
    if ((int)(blockIdx.x * NV * 32) < Z) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        nA[c] = __ldg(vn + c0 + (size_t)c * V);
        nG[c] = __ldg(gn + c0 + (size_t)c * V);
      }
    }
  }
#pragma unroll 1
  for (int c4 = 0; c4 < NV; ++c4) {
    const int kb = (blockIdx.x * NV + c4) * 32;
    if (kb >= Z) break;  // Z % 32 == 0: chunks are whole
    const int k = kb + lane;
    const int c0 = row + k;
    const float A0 = nA[0], A1 = nA[1], A2 = nA[2];
    const float g0 = nG[0], g1 = nG[1], g2 = nG[2];
    if (c4 + 1 < NV && kb + 32 < Z) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        nA[c] = __ldg(vn + c0 + 32 + (size_t)c * V);
        nG[c] = __ldg(gn + c0 + 32 + (size_t)c * V);
      }
    }
    Corners cs;
    corner_setup<NEED_PHI>(cs, coord_f32(fi, A0, dh, dl), coord_f32(fj, A1, dh, dl),
                           coord_f32((float)k, A2, dh, dl), X, Y, Z, sx, sy, lane);
    const float g[3] = {g0, g1, g2};
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const Vals q = corner_load(pb[c], cs);
      float gx, gy, gz;
      corner_grad(q, cs, gx, gy, gz);
      const float gd = mul_f32_by_double(g[c], dh, dl);  // "diff *= dt", cuda/interp.cu:230
      a0 = a0 + gx * gd;
      a1 = a1 + gy * gd;
      a2 = a2 + gz * gd;
      if (NEED_PHI) corner_splat(Sb[c], cs, g[c]);
    }
    dvn[c0] = a0 + dsf * g0;  // d_u + ds*g (deform.py ComposeFunction.backward)
    dvn[c0 + V] = a1 + dsf * g1;
    dvn[c0 + 2 * V] = a2 + dsf * g2;
  }
}

// ---- Ad_star backward: m_c = sum_d (D_d phi_c + delta_cd) mi_d, mi = m0(x + phi) ---------------------
template <int NV, bool NEED_M0, bool NEED_PHI>
__global__ void __launch_bounds__(256)
adstar_bwd3_kernel(float* __restrict__ mi_out, float* __restrict__ d_m0, float* __restrict__ S,
                   const float* __restrict__ phi, const float* __restrict__ dm, const float* __restrict__ m0,
                   int X, int Y, int Z) {
  LGM_PREFETCH_ROWS_AHEAD(9, (threadIdx.x < 3 ? phi : (threadIdx.x < 6 ? dm : m0)), X, Y, Z, 0)
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (j >= Y) return;
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const size_t nb = (size_t)n * 3 * V;
  const float* pn = phi + nb;
  const float* dmn = dm + nb;
  const float* mn = m0 + nb;
  float* min_ = mi_out + nb;
  float* Sn = S + nb;
  const float* mb[3] = {mn, mn + V, mn + 2 * (size_t)V};
  float* db[3] = {d_m0 + nb, d_m0 + nb + V, d_m0 + nb + 2 * (size_t)V};
  asm volatile("" : "+l"(mb[0]), "+l"(mb[1]), "+l"(mb[2]), "+l"(db[0]), "+l"(db[1]), "+l"(db[2]));
  const int row = i * sx + j * sy;
  const float fi = (float)i, fj = (float)j;
  const int xm = (i > 0) ? -sx : 0, xp = (i < X - 1) ? sx : 0;
  const int ym = (j > 0) ? -sy : 0, yp = (j < Y - 1) ? sy : 0;
  const int lane = threadIdx.x;
  // software pipeline over the chunks (see compose_bwd3_kernel): the displacement of chunk c4+1 is
  // loaded before chunk c4 is processed, so a chunk's corner gather does not wait for a second round trip
  float nA[3];
  {
    const int c0 = row + blockIdx.x * NV * 32 + lane;
    if ((int)(blockIdx.x * NV * 32) < Z) {
#pragma unroll
      for (int c = 0; c < 3; ++c) nA[c] = __ldg(pn + c0 + (size_t)c * V);
    }
  }
#pragma unroll 1
  for (int c4 = 0; c4 < NV; ++c4) {
    const int kb = (blockIdx.x * NV + c4) * 32;
    if (kb >= Z) break;
    const int k = kb + lane;
    const int c0 = row + k;
    const int zm = (k > 0) ? -1 : 0, zp = (k < Z - 1) ? 1 : 0;
    const float* pc = pn + c0;
    const float A0 = nA[0], A1 = nA[1], A2 = nA[2];
    if (c4 + 1 < NV && kb + 32 < Z) {
#pragma unroll
      for (int c = 0; c < 3; ++c) nA[c] = __ldg(pc + 32 + (size_t)c * V);
    }
    // S was completed by compose_bwd3 (previous launch): its three values of this voxel are read with
    // the other loads of the chunk, not in a round trip of their own after the splats
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (NEED_PHI) {
      s0 = Sn[c0];
      s1 = Sn[c0 + V];
      s2 = Sn[c0 + 2 * V];
    }
    // q_d = sum_c (D_d phi_c + delta_cd) dm_c   (jtvf backward d_w, cuda/diff.cu:417-431)
    float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* f = pc + (size_t)c * V;
      float d0 = 0.5f * (__ldg(f + xp) - __ldg(f + xm));
      float d1 = 0.5f * (__ldg(f + yp) - __ldg(f + ym));
      float d2 = 0.5f * (__ldg(f + zp) - __ldg(f + zm));
      if (c == 0) d0 += 1.f;
      if (c == 1) d1 += 1.f;
      if (c == 2) d2 += 1.f;
      const float gc = __ldg(dmn + c0 + (size_t)c * V);
      q0 += d0 * gc;
      q1 += d1 * gc;
      q2 += d2 * gc;
    }
    Corners cs;
    corner_setup<NEED_M0>(cs, __fadd_rn(fi, A0), __fadd_rn(fj, A1), __fadd_rn((float)k, A2), X, Y, Z, sx, sy, lane);
    const float q[3] = {q0, q1, q2};
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const Vals val = corner_load(mb[c], cs);
      if (NEED_PHI) {
        min_[c0 + (size_t)c * V] = corner_value(val, cs);
        float gx, gy, gz;
        corner_grad(val, cs, gx, gy, gz);
        a0 = a0 + gx * q[c];  // dt == 1: "diff *= dt" is exact
        a1 = a1 + gy * q[c];
        a2 = a2 + gz * q[c];
      }
      if (NEED_M0) corner_splat(db[c], cs, q[c]);
    }
    if (NEED_PHI) {  // this thread owns voxel c0 of S now (the splats into S finished in compose_bwd3)
      Sn[c0] = s0 + a0;
      Sn[c0 + V] = s1 + a1;
      Sn[c0 + 2 * V] = s2 + a2;
    }
  }
}

// ---- stencil part of the jtvf backward: G_c <- S_c + sum_d D_d^T (mi_d * dm_c); S <- 0 --------------
// D_d^T incl. boundary rows (cuda/diff.cu:432-460) written branch-free:
//   -0.5 * (s_hi * P(hi) - s_lo * P(lo)),  hi/lo clamped, s = -1 where the clamp is active.
template <int NV>
__global__ void __launch_bounds__(256)
stencil_bwd3_kernel(float* __restrict__ G, float* __restrict__ S, const float* __restrict__ mi,
                    const float* __restrict__ dm, int X, int Y, int Z) {
  LGM_PREFETCH_ROWS_AHEAD(9, (threadIdx.x < 3 ? mi : (threadIdx.x < 6 ? dm : S)), X, Y, Z, 0)
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (j >= Y) return;
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const size_t nb = (size_t)n * 3 * V;
  const float* min_ = mi + nb;
  const float* dmn = dm + nb;
  float* Gn = G + nb;
  float* Sn = S + nb;
  const int row = i * sx + j * sy;
  const int off_lo[2] = {(i > 0) ? -sx : 0, (j > 0) ? -sy : 0};
  const int off_hi[2] = {(i < X - 1) ? sx : 0, (j < Y - 1) ? sy : 0};
  const float s_lo[2] = {(i > 0) ? 1.f : -1.f, (j > 0) ? 1.f : -1.f};
  const float s_hi[2] = {(i < X - 1) ? 1.f : -1.f, (j < Y - 1) ? 1.f : -1.f};
#pragma unroll
  for (int c4 = 0; c4 < NV; ++c4) {
    const int k = (blockIdx.x * NV + c4) * 32 + threadIdx.x;
    if (k >= Z) break;
    const int c0 = row + k;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int lo = (d < 2) ? off_lo[d] : ((k > 0) ? -1 : 0);
      const int hi = (d < 2) ? off_hi[d] : ((k < Z - 1) ? 1 : 0);
      const float sl = (d < 2) ? s_lo[d] : ((k > 0) ? 1.f : -1.f);
      const float sh = (d < 2) ? s_hi[d] : ((k < Z - 1) ? 1.f : -1.f);
      const float* w = min_ + (size_t)d * V + c0;
      const float wh = sh * __ldg(w + hi), wl = sl * __ldg(w + lo);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* z = dmn + (size_t)c * V + c0;
        acc[c] += -0.5f * (wh * __ldg(z + hi) - wl * __ldg(z + lo));
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Gn[c0 + (size_t)c * V] = Sn[c0 + (size_t)c * V] + acc[c];
      Sn[c0 + (size_t)c * V] = 0.f;
    }
  }
}

}  // namespace

bool epdiff_bwd3_ok(int64_t N, const int64_t* sh) {
  if (sh[0] < 2 || sh[1] < 2 || sh[2] < 2 || sh[2] % 32 != 0) return false;
  if (sh[0] * sh[1] * sh[2] >= (1LL << 31) / 4) return false;
  if (N * sh[0] > 65535 || sh[1] > 65535LL) return false;
  return true;
}

static dim3 bwd_grid(int64_t N, const int64_t* sh) {
  return dim3((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0]));
}

int compose_bwd3_f32(void* dv, void* S, const void* G, const void* phi, const void* v, int64_t N,
                     const int64_t* sh, double ds, bool need_phi, cudaStream_t s) {
  const float dh = (float)ds, dl = (float)(ds - (double)dh);
  const dim3 grid = bwd_grid(N, sh), block(32, 8);
  if (need_phi)
    compose_bwd3_kernel<4, true><<<grid, block, 0, s>>>((float*)dv, (float*)S, (const float*)G, (const float*)phi,
                                                        (const float*)v, (int)sh[0], (int)sh[1], (int)sh[2], dh, dl, (float)ds);
  else
    compose_bwd3_kernel<4, false><<<grid, block, 0, s>>>((float*)dv, (float*)S, (const float*)G, (const float*)phi,
                                                         (const float*)v, (int)sh[0], (int)sh[1], (int)sh[2], dh, dl, (float)ds);
  count_launch("compose_bwd", s);
  return finish(s, "lgm_epdiff_step_bwd(compose)");
}

int adstar_bwd3_f32(void* mi, void* d_m0, void* S, const void* phi, const void* dm, const void* m0, int64_t N,
                    const int64_t* sh, bool need_m0, bool need_phi, cudaStream_t s) {
  const dim3 grid = bwd_grid(N, sh), block(32, 8);
#define LGM_ADB(M0, PHI)                                                                                       \
  adstar_bwd3_kernel<4, M0, PHI><<<grid, block, 0, s>>>((float*)mi, (float*)d_m0, (float*)S, (const float*)phi, \
                                                        (const float*)dm, (const float*)m0, (int)sh[0],         \
                                                        (int)sh[1], (int)sh[2])
  if (need_m0 && need_phi) LGM_ADB(true, true);
  else if (need_m0) LGM_ADB(true, false);
  else LGM_ADB(false, true);
#undef LGM_ADB
  count_launch("Ad_star_bwd", s);
  return finish(s, "lgm_epdiff_step_bwd(Ad_star)");
}

int stencil_bwd3_f32(void* G, void* S, const void* mi, const void* dm, int64_t N, const int64_t* sh,
                     cudaStream_t s) {
  stencil_bwd3_kernel<4><<<bwd_grid(N, sh), dim3(32, 8), 0, s>>>((float*)G, (float*)S, (const float*)mi,
                                                                (const float*)dm, (int)sh[0], (int)sh[1], (int)sh[2]);
  count_launch("jtvf_bwd_stencil", s);
  return finish(s, "lgm_epdiff_step_bwd(stencil)");
}

}  // namespace lgm
