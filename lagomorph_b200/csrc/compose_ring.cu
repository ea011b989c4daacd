// compose_ring.cu -- compose (out = ds*u + dt*v(x + ds*u), lagomorph/deform.py:53-62) for SMALL
// displacements ds*u, i.e. the last stage of an EPDiff step (ds*u = -dt*velocity, a fraction of a
// voxel): the gather source v is staged in shared memory by bulk asynchronous copies
// (cp.async.bulk + mbarrier, the TMA engine; UBLKCP in SASS) instead of being gathered through L1.
//
// Why: the planar gather kernel (gather3.cu) is bound by L1 wavefronts -- each of its 24 corner loads
// is a warp-wide run of 32 consecutive but UNALIGNED floats = two cache lines = two wavefronts. Shared
// memory has no lines: the same run is one wavefront. A CTA owns TY y rows x all z of XS consecutive x
// slabs and marches along x with a RING of 4 x planes of v (rows y-1 .. y+TY, all z, 3 channels):
// while slab x is sampled from planes x-1, x, x+1, plane x+2 streams in. Every plane row is fetched
// (TY+2)/TY times from L2 instead of ~8 times through L1.
// A warp whose samples all fall inside the staged window reads its corners with LDS; any other warp
// (large displacement) takes the global gather of gather3.cu for that chunk. Arithmetic and
// evaluation order are those of gather3_kernel<1>: results are bit-identical.
#include <cmath>
#include <cstdlib>
#include "gather_common.cuh"
#include "ring_common.cuh"

#ifndef LGM_RING_XS
#define LGM_RING_XS 16  /* x slabs marched by one CTA */
#endif
#ifndef LGM_RING_PF
#define LGM_RING_PF 2  /* L2 prefetch of the velocity rows this many slabs ahead (0 = off); 2 measured best of 2/4/8 */
#endif
#ifndef LGM_RING_MINB
#define LGM_RING_MINB 2  /* 2 measured best (C2 0.269 ms vs 0.283 at 3, 0.282 at 4): registers for loads in flight */
#endif

namespace lgm {

namespace {

constexpr int kRing = 4;
constexpr unsigned kFullMask = 0xffffffffu;

// WPR warps share one z row (Z = 128 * WPR for Z > 128), a CTA of 8 warps covers TY = 8 / WPR rows;
// NV chunks of 32 per thread. blockDim = (32, 8).
// Z is a compile-time constant (32 * NV * WPR): every offset inside the ring is an immediate.
template <int NV, int WPR>
__global__ void __launch_bounds__(256, LGM_RING_MINB)
compose_ring_kernel(float* __restrict__ out, const float* __restrict__ u, const float* __restrict__ v, int X, int Y,
                    float dh, float dl, float dsr, float dtr, int xs, int rev) {
  constexpr int TY = 8 / WPR, ROWS = TY + 2, Z = 32 * NV * WPR;
  extern __shared__ __align__(128) unsigned char ring_raw[];
  float* ring = reinterpret_cast<float*>(ring_raw);                 // [kRing][3][ROWS][Z]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (size_t)kRing * 3 * ROWS * Z);
  const int lane = threadIdx.x, w = threadIdx.y, tid = w * 32 + lane;
  const unsigned bxi = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const unsigned byi = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int n = blockIdx.z;
  const int y0t = byi * TY, yb = y0t - 1;
  const int j = y0t + w / WPR;
  const int zoff = (w % WPR) * (32 * NV);
  const int xs0 = bxi * xs, xs1 = min(X, xs0 + xs);  // slabs [xs0, xs1)
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const int ylo = max(yb, 0), yhi = min(y0t + TY, Y - 1);             // staged rows of every plane
  const unsigned plane_bytes = (unsigned)((yhi - ylo + 1) * Z * 4);
  constexpr int CH = ROWS * Z;                                         // channel stride inside a ring slot
  const float* un = u + (size_t)n * 3 * V;
  const float* vn = v + (size_t)n * 3 * V;
  const float* vn1 = vn + V;
  const float* vn2 = vn1 + V;
  float* on = out + (size_t)n * 3 * V;
  asm volatile("" : "+l"(vn), "+l"(vn1), "+l"(vn2));
  const unsigned four = opaque_four();
  const float hiX = (float)X - 0.5f, hiY = (float)Y - 0.5f, hiZ = (float)Z - 0.5f;
  const int plo = max(xs0 - 1, 0), phi_ = min(xs1, X - 1);            // planes this CTA ever needs

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kRing; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int p) {  // one thread: plane p of the three channels -> ring slot p % kRing
    if (p < plo || p > phi_) return;
    const int slot = p & (kRing - 1);
    mbar_expect_tx(&full[slot], 3 * plane_bytes);
    float* dst = ring + (size_t)slot * 3 * CH + (ylo - yb) * Z;
    const float* src = vn + (size_t)p * sx + (size_t)ylo * Z;
#pragma unroll
    for (int c = 0; c < 3; ++c) bulk_g2s(dst + c * CH, src + (size_t)c * V, plane_bytes, &full[slot]);
  };
  if (tid == 0) {
    issue(xs0 - 1);
    issue(xs0);
    issue(xs0 + 1);
  }
  unsigned phase = 0;  // parity to wait for, per slot
  int waited = plo - 1;
  const bool rowok = j < Y;
  const float fj = (float)j;
  const int row_j = j * sy;

  for (int x = xs0; x < xs1; ++x) {
    __syncthreads();  // slab x-1 is done everywhere: the slot of plane x-2 may be overwritten by plane x+2
    if (tid == 0) issue(x + 2);  // look-ahead (planes up to xs0+1 were issued in the prologue)
    if (LGM_RING_PF > 0 && tid >= 32 && tid < 35 && x + LGM_RING_PF < xs1 && y0t + TY <= Y) {
      // the centre loads of slab x + LGM_RING_PF (velocity rows of this tile) -> L2
      const float* src = un + (size_t)(tid - 32) * V + (size_t)(x + LGM_RING_PF) * sx + (size_t)y0t * Z;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((unsigned)(TY * Z * 4)) : "memory");
    }
    const int need = min(x + 1, phi_);
    while (waited < need) {
      ++waited;
      const int slot = waited & (kRing - 1);
      mbar_wait(&full[slot], (phase >> slot) & 1u);
      phase ^= 1u << slot;
    }
    if (!rowok) continue;
    const float fi = (float)x;
    const int row0 = x * sx + row_j;
    float Apre[NV][3];
#pragma unroll
    for (int c4 = 0; c4 < NV; ++c4) {
      const int k = zoff + c4 * 32 + lane;
      Apre[c4][0] = __ldg(un + (row0 + k));
      Apre[c4][1] = __ldg(un + V + (row0 + k));
      Apre[c4][2] = __ldg(un + 2 * (size_t)V + (row0 + k));
    }
#pragma unroll
    for (int c4 = 0; c4 < NV; ++c4) {
      const int k = zoff + c4 * 32 + lane;
      const int c0 = row0 + k;
      const float A0 = Apre[c4][0], A1 = Apre[c4][1], A2 = Apre[c4][2];
      const float hx = coord_f32(fi, A0, dh, dl);
      const float hy = coord_f32(fj, A1, dh, dl);
      const float hz = coord_f32((float)k, A2, dh, dl);
      const Ax3 ax = axis_fwd(hx, X, hiX), ay = axis_fwd(hy, Y, hiY), az = axis_fwd(hz, Z, hiZ);
      int zs;
      float wv;
      z_pair(az, Z, zs, wv);
      const float t = ax.t, uu = ay.t;
      const float omt = 1.f - t, omu = 1.f - uu, omv = 1.f - wv;
      const bool inwin = (ax.i0 >= x - 1) && (ax.i1 <= x + 1) && (ay.i0 >= yb) && (ay.i1 <= y0t + TY);
      float m0v, m1v, m2v;
      if (__all_sync(kFullMask, inwin)) {
        const int s0 = (ax.i0 & (kRing - 1)) * 3 * CH, s1 = (ax.i1 & (kRing - 1)) * 3 * CH;
        const int r0 = (ay.i0 - yb) * Z + zs, r1 = (ay.i1 - yb) * Z + zs;
        const float* p00 = ring + s0 + r0;
        const float* p01 = ring + s0 + r1;
        const float* p10 = ring + s1 + r0;
        const float* p11 = ring + s1 + r1;
        float r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v0 = p00[c * CH], v4 = p00[c * CH + 1];
          const float v3 = p01[c * CH], v7 = p01[c * CH + 1];
          const float v1 = p10[c * CH], v5 = p10[c * CH + 1];
          const float v2 = p11[c * CH], v6 = p11[c * CH + 1];
          r[c] = lerp8(v0, v1, v2, v3, v4, v5, v6, v7, t, uu, wv, omt, omu, omv);
        }
        m0v = r[0];
        m1v = r[1];
        m2v = r[2];
      } else {  // some sample of this warp leaves the staged window: gather from global memory
        const unsigned rx0 = ax.i0 * sx + zs, rx1 = ax.i1 * sx + zs;
        const unsigned ry0 = ay.i0 * sy, ry1 = ay.i1 * sy;
        const unsigned i00 = rx0 + ry0, i01 = rx0 + ry1, i10 = rx1 + ry0, i11 = rx1 + ry1;
        m0v = trilerp(vn, i00, i01, i10, i11, t, uu, wv, omt, omu, omv, four);
        m1v = trilerp(vn1, i00, i01, i10, i11, t, uu, wv, omt, omu, omv, four);
        m2v = trilerp(vn2, i00, i01, i10, i11, t, uu, wv, omt, omu, omv, four);
      }
      on[c0] = __fadd_rn(__fmul_rn(dsr, A0), __fmul_rn(dtr, m0v));
      on[c0 + V] = __fadd_rn(__fmul_rn(dsr, A1), __fmul_rn(dtr, m1v));
      on[c0 + 2 * (size_t)V] = __fadd_rn(__fmul_rn(dsr, A2), __fmul_rn(dtr, m2v));
    }
  }
}

}  // namespace

// LGM_EUNSUP when the ring kernel does not apply (caller uses the planar gather kernel)
int compose3_ring_f32(void* out, const void* u, const void* v, int64_t N, const int64_t* sh, double ds, double dt,
                      int rev, cudaStream_t s) {
  static const bool off = getenv("LGM_NO_RING") != nullptr;  // kernel experiments
  const int64_t X = sh[0], Y = sh[1], Z = sh[2];
  if (off || X < 4 || Y < 2 || N > 65535 || X * Y * Z >= (1LL << 31) / 4) return LGM_EUNSUP;
  if (!(Z == 32 || Z == 64 || Z == 128 || Z == 256)) return LGM_EUNSUP;
  if (((uintptr_t)v & 15) != 0) return LGM_EUNSUP;  // bulk copies need 16-byte aligned rows
  if (fabs(ds) > 1.0) return LGM_EUNSUP;            // staging pays for sub-voxel displacements only
  const float dh = (float)ds, dl = (float)(ds - (double)dh);
  const int wpr = Z == 256 ? 2 : 1, TY = 8 / wpr;
  const size_t smem = (size_t)kRing * 3 * (TY + 2) * Z * 4 + kRing * 8;
  // x slabs marched by one CTA: LGM_RING_XS for large batches (each plane row is fetched (TY+2)/TY * (xs+2)/xs times);
  // shorter marches when the grid would not fill the GPU twice over (small batches: chunks of expmap_host,
  // single registrations), so that CTAs = N * Y/TY * X/xs stays above ~6 per SM
  int xs = LGM_RING_XS;
  while (xs > 4 && N * cdiv(Y, TY) * cdiv(X, xs) < 6 * 148) xs /= 2;
  dim3 grid((unsigned)cdiv(X, xs), (unsigned)cdiv(Y, TY), (unsigned)N), block(32, 8);
#define LGM_RING(NV_, WPR_)                                                                                          \
  do {                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(compose_ring_kernel<NV_, WPR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                                 \
    if (e != cudaSuccess) return set_error((int)e, "compose ring smem: %s", cudaGetErrorString(e));                  \
    compose_ring_kernel<NV_, WPR_><<<grid, block, smem, s>>>((float*)out, (const float*)u, (const float*)v, (int)X,   \
                                                             (int)Y, dh, dl, (float)ds, (float)dt, xs, rev);             \
  } while (0)
  if (Z == 32) LGM_RING(1, 1);
  else if (Z == 64) LGM_RING(2, 1);
  else if (Z == 128) LGM_RING(4, 1);
  else LGM_RING(4, 2);
#undef LGM_RING
  count_launch("compose", s);
  return finish(s, "lgm_compose_fwd(ring)");
}

}  // namespace lgm
