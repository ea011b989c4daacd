// adstar_ring.cu -- Ad_star (out_c = sum_d (D_d phi_c + delta_cd) * m_d(x + phi(x)), adjrep.py:86-97) with
// the STENCIL operand phi staged in shared memory by bulk asynchronous copies (cp.async.bulk + mbarrier,
// the TMA engine), same marching scheme as compose_ring.cu.
//
// Why: gather3_kernel<0> is bound by L1 wavefronts. Of its ~79 wavefronts per 32 voxels, 27-30 are the
// 21 loads of phi: the 3 centre values and their 18 stencil neighbours, where every z+-1 neighbour is an
// unaligned 128-byte run (two cache lines = two wavefronts) and every x+-1 row misses L1. From shared
// memory each of the 21 reads is one wavefront whatever its alignment, nothing of phi passes through
// L1 (which is left to the m0 gather), and the centre values no longer gate a chunk with a DRAM round
// trip. A CTA owns TY y rows x all z of XS consecutive x slabs and marches along x with a ring of 4 x
// planes of phi (rows y-1 .. y+TY, 3 channels): while slab x is worked on from planes x-1, x, x+1,
// plane x+2 streams in. The m0 gather (displacements of several voxels) stays the L1 gather of
// gather3.cu. Arithmetic and evaluation order are those of gather3_kernel<0>: results are bit-identical.
// Rows of 256 voxels: the row is cut into two z tiles of 128 + 8 words (template parameter NZT) so that a
// CTA keeps 8 + 2 rows in 64 KB, and each plane's box (3 channels x 10 rows x 136 words, strided in global
// memory) is staged by one TMA TENSOR copy through a tensor map of the phi field (TM = true;
// cp.async.bulk.tensor.4d, UTMALDG in SASS; ring_common.cuh: make_field_tmap / tma_box4_g2s).
#include <cstdlib>
#include <cstring>
#include "gather_common.cuh"
#include "ring_common.cuh"

#ifndef LGM_ARING_XS
#define LGM_ARING_XS 16  /* x slabs marched by one CTA */
#endif
#ifndef LGM_ARING_PF
#define LGM_ARING_PF 2   /* L2 prefetch of the m0 rows of this tile this many slabs ahead (0 = off) */
#endif
#ifndef LGM_ARING_MINB
#define LGM_ARING_MINB 3
#endif

namespace lgm {

namespace {

constexpr int kARing = 4;

// WPR warps share one z row (Z = 128 * WPR for Z > 128), a CTA of 8 warps covers TY = 8 / WPR rows;
// NV chunks of 32 per thread; blockDim = (32, 8). Z = 32 * NV * WPR * NZT is a compile-time constant.
// NZT > 1: the row is cut into NZT z tiles (gridDim.y = y tiles * NZT), each staged with kZHalo words of
// halo on the inner side(s) as one bulk copy per row (ZW = 32 NV WPR + 2 kZHalo words, 16-byte aligned
// start) instead of one per plane: 8 + 2 rows per CTA at Z = 256 with the shared memory of the Z = 128 kernel.
constexpr int kZHalo = 4;
template <int NV, int WPR, int NZT, bool TM>
__global__ void __launch_bounds__(256, LGM_ARING_MINB)
adstar_ring_kernel(float* __restrict__ out, const float* __restrict__ phi, const float* __restrict__ m, int X, int Y,
                   int xs, int rev, const __grid_constant__ CUtensorMap tmap, int use_tmap) {
  constexpr int TY = 8 / WPR, ROWS = TY + 2, ZT = 32 * NV * WPR, Z = ZT * NZT;
  constexpr int ZW = (NZT == 1) ? Z : ZT + 2 * kZHalo;               // staged words per row
  constexpr int CH = ROWS * ZW;                                      // channel stride inside a ring slot
  constexpr int SLOT = TM ? (3 * CH + 31) / 32 * 32 : 3 * CH;        // slot stride: 128-byte aligned for TMA tensor copies
  extern __shared__ __align__(128) unsigned char aring_raw[];
  float* ring = reinterpret_cast<float*>(aring_raw);                 // [kARing][3][ROWS][ZW]
  // mbarriers: behind the ring, or (z tiles) in the padding of slot 0 so that the CTA stays at 64 KB:
  // three CTAs per SM inside the 196 KB carve-out (one KB more and the SM drops to two, or to a smaller L1)
  static_assert(!TM || SLOT - 3 * CH >= 2 * kARing, "no room for the mbarriers in the slot padding");
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (TM ? (size_t)3 * CH : (size_t)kARing * SLOT));
  const int lane = threadIdx.x, w = threadIdx.y, tid = w * 32 + lane;
  const unsigned bxi = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const unsigned byz = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const unsigned byi = byz / NZT, zt = byz % NZT;
  const int n = blockIdx.z;
  const int y0t = byi * TY, yb = y0t - 1;
  const int j = y0t + w / WPR;
  const int zoff = zt * ZT + (w % WPR) * (32 * NV);
  // first staged z of this tile: its halo, shifted inwards at the two ends of the row
  const int zlo = (NZT == 1) ? 0 : min(max((int)zt * ZT - kZHalo, 0), Z - ZW);
  const int xs0 = bxi * xs, xs1 = min(X, xs0 + xs);  // slabs [xs0, xs1)
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const int ylo = max(yb, 0), yhi = min(y0t + TY, Y - 1);             // staged rows of every plane
  const unsigned plane_bytes = (unsigned)((yhi - ylo + 1) * ZW * 4);
  const float* pn = phi + (size_t)n * 3 * V;
  const float* bn = m + (size_t)n * 3 * V;
  const float* bn1 = bn + V;
  const float* bn2 = bn1 + V;
  float* on = out + (size_t)n * 3 * V;
  asm volatile("" : "+l"(bn), "+l"(bn1), "+l"(bn2));
  const unsigned four = opaque_four();
  const float hiX = (float)X - 0.5f, hiY = (float)Y - 0.5f, hiZ = (float)Z - 0.5f;
  const int plo = max(xs0 - 1, 0), phi_ = min(xs1, X - 1);            // planes this CTA ever needs

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kARing; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // warp 0: plane p of the three channels -> ring slot p % kARing. NZT == 1: lane 0 issues one bulk copy per
  // channel (the staged rows are contiguous); NZT > 1: one bulk copy per (channel, row), spread over the lanes.
  auto issue = [&](int p) {
    if (p < plo || p > phi_) return;
    const int slot = p & (kARing - 1);
    if constexpr (TM) {
      // one tensor copy: box (3 channels, 1 plane, ROWS rows from yb, ZW words from zlo); rows outside the
      // volume arrive as zeros and are never read (clamped stencil rows)
      if (lane == 0) {
        mbar_expect_tx(&full[slot], (unsigned)(3 * CH * 4));
        tma_box4_g2s(ring + (size_t)slot * SLOT, &tmap, zlo, yb, p, 3 * n, &full[slot]);
      }
      return;
    }
    if (lane == 0) mbar_expect_tx(&full[slot], 3 * plane_bytes);
    float* dst = ring + (size_t)slot * SLOT + (ylo - yb) * ZW;
    const float* src = pn + (size_t)p * sx + (size_t)ylo * Z + zlo;
    if constexpr (NZT == 1) {
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) bulk_g2s(dst + c * CH, src + (size_t)c * V, plane_bytes, &full[slot]);
      }
    } else {
      __syncwarp();
      const int nr = yhi - ylo + 1;
      for (int i = lane; i < 3 * nr; i += 32) {
        const int c = i / nr, r = i - c * nr;
        bulk_g2s(dst + c * CH + r * ZW, src + (size_t)c * V + (size_t)r * Z, (unsigned)(ZW * 4), &full[slot]);
      }
    }
  };
  if (w == 0) {
    issue(xs0 - 1);
    issue(xs0);
    issue(xs0 + 1);
  }
  unsigned phase = 0;  // parity to wait for, per slot
  int waited = plo - 1;
  const bool rowok = j < Y;
  const float fj = (float)j;
  // ring row offsets of this thread's y row and its clamped y neighbours (diff.h: clamped indices)
  const int rj = (j - yb) * ZW - zlo;   // + global z = column of the staged row
  const int rjm = (j > 0) ? rj - ZW : rj, rjp = (j < Y - 1) ? rj + ZW : rj;

  for (int x = xs0; x < xs1; ++x) {
    __syncthreads();  // slab x-1 is done everywhere: the slot of plane x-2 may be overwritten by plane x+2
    if (w == 0) issue(x + 2);  // look-ahead (planes up to xs0+1 were issued in the prologue)
    if (LGM_ARING_PF > 0 && tid >= 32 && tid < 35 && x + LGM_ARING_PF < xs1 && y0t + TY <= Y) {
      // the undisplaced rows of m0 for slab x + LGM_ARING_PF (where most of its gather lands) -> L2
      // (NZT z tiles share the TY rows: tile zt takes rows [zt, zt + 1) * TY / NZT)
      const float* src = bn + (size_t)(tid - 32) * V + (size_t)(x + LGM_ARING_PF) * sx +
                         (size_t)(y0t + (int)zt * (TY / NZT)) * Z;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((unsigned)(TY / NZT * Z * 4)) : "memory");
    }
    const int need = min(x + 1, phi_);
    while (waited < need) {
      ++waited;
      const int slot = waited & (kARing - 1);
      mbar_wait(&full[slot], (phase >> slot) & 1u);
      phase ^= 1u << slot;
    }
    if (!rowok) continue;
    const float fi = (float)x;
    const int row0 = x * sx + j * sy;
    const float* sc = ring + (x & (kARing - 1)) * SLOT;                               // plane x
    const float* sm = (x > 0) ? ring + ((x - 1) & (kARing - 1)) * SLOT : sc;          // plane x-1 (clamped)
    const float* sp = (x < X - 1) ? ring + ((x + 1) & (kARing - 1)) * SLOT : sc;      // plane x+1 (clamped)
#pragma unroll
    for (int c4 = 0; c4 < NV; ++c4) {
      const int k = zoff + c4 * 32 + lane;
      const int c0 = row0 + k;
      const float A0 = sc[rj + k], A1 = sc[CH + rj + k], A2 = sc[2 * CH + rj + k];
      const float hx = __fadd_rn(fi, A0);  // dt == 1: the double sum is exact before rounding
      const float hy = __fadd_rn(fj, A1);
      const float hz = __fadd_rn((float)k, A2);
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
      const int kp = (k < Z - 1) ? k + 1 : k, km = (k > 0) ? k - 1 : k;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float g0 = 0.5f * (sp[c * CH + rj + k] - sm[c * CH + rj + k]);
        float g1 = 0.5f * (sc[c * CH + rjp + k] - sc[c * CH + rjm + k]);
        float g2 = 0.5f * (sc[c * CH + rj + kp] - sc[c * CH + rj + km]);
        if (c == 0) g0 += 1.f;
        if (c == 1) g1 += 1.f;
        if (c == 2) g2 += 1.f;
        on[c0 + (size_t)c * V] = g0 * m0v + g1 * m1v + g2 * m2v;  // diff.cu:118-120
      }
    }
  }
}

}  // namespace

// LGM_EUNSUP when the ring kernel does not apply (caller uses the planar gather kernel)
int Ad_star3_ring_f32(void* out, const void* phi, const void* m, int64_t N, const int64_t* sh, int rev,
                      cudaStream_t s) {
  static const bool off = getenv("LGM_NO_ADSTAR_RING") != nullptr;  // kernel experiments
  const int64_t X = sh[0], Y = sh[1], Z = sh[2];
  if (off || X < 4 || Y < 2 || N > 65535 || X * Y * Z >= (1LL << 31) / 4) return LGM_EUNSUP;
  // Z = 256: two z tiles of 128 + 8 words, 8 + 2 rows per CTA, each plane's box (3 channels x 10 rows x 136
  // words) staged by ONE TMA tensor copy (cp.async.bulk.tensor.4d, UTMALDG). 8 x 256^3 on a B200
  // (r3_aring256_tmap_ab3.log, r3_ring256_tmap_ab4.log): planar kernel 1.23-1.25 ms, z tiles with one bulk
  // copy per row 1.22-1.27, with the tensor copy 1.09-1.14; shoot 22.86 -> 22.01 ms. LGM_ADSTAR_RING_256=0:
  // planar kernel; =2: two warps per row, 4 + 2 full rows per CTA (round 2; slower than planar: 1.35).
  // The same z tiles for compose measured SLOWER than its two-warps-per-row layout (0.97 vs 0.83 ms): not built in.
  static const int mode256 = getenv("LGM_ADSTAR_RING_256") ? atoi(getenv("LGM_ADSTAR_RING_256")) : 1;
  if (!(Z == 32 || Z == 64 || Z == 128 || (Z == 256 && mode256 != 0))) return LGM_EUNSUP;
  if (((uintptr_t)phi & 15) != 0) return LGM_EUNSUP;  // bulk copies need 16-byte aligned rows
  const bool wide256 = mode256 == 2;
  const int wpr = (Z == 256 && wide256) ? 2 : 1, TY = 8 / wpr;
  const int nzt = (Z == 256 && !wide256) ? 2 : 1;
  const int zw = nzt == 1 ? (int)Z : (int)Z / nzt + 2 * kZHalo;

  // z tiles: the staged box (3 channels x 10 rows x 136 words of one x plane) is one TMA tensor copy
  // (LGM_ADSTAR_RING_TMAP=0: one bulk copy per row instead)
  static const bool tmap_off = getenv("LGM_ADSTAR_RING_TMAP") != nullptr && atoi(getenv("LGM_ADSTAR_RING_TMAP")) == 0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  const int use_tmap = (nzt > 1 && !tmap_off && make_field_tmap(&tmap, phi, N * 3, X, Y, Z, 3, TY + 2, zw)) ? 1 : 0;
  const size_t slot_words = use_tmap ? ((size_t)3 * (TY + 2) * zw + 31) / 32 * 32 : (size_t)3 * (TY + 2) * zw;
  const size_t smem = (size_t)kARing * slot_words * 4 + (use_tmap ? 0 : kARing * 8);
  // x slabs marched by one CTA: LGM_ARING_XS for large batches (each plane row is fetched (TY+2)/TY * (xs+2)/xs times);
  // shorter marches when the grid would not fill the GPU twice over (small batches: chunks of expmap_host,
  // single registrations), so that CTAs = N * Y/TY * X/xs stays above ~6 per SM
  int xs = LGM_ARING_XS;
  while (xs > 4 && N * cdiv(Y, TY) * nzt * cdiv(X, xs) < 6 * 148) xs /= 2;
  dim3 grid((unsigned)cdiv(X, xs), (unsigned)(cdiv(Y, TY) * nzt), (unsigned)N), block(32, 8);
#define LGM_ARING(NV_, WPR_, NZT_, TM_)                                                                                   \
  do {                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(adstar_ring_kernel<NV_, WPR_, NZT_, TM_>,                                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                    \
    if (e != cudaSuccess) return set_error((int)e, "Ad_star ring smem: %s", cudaGetErrorString(e));                  \
    adstar_ring_kernel<NV_, WPR_, NZT_, TM_><<<grid, block, smem, s>>>((float*)out, (const float*)phi, (const float*)m,   \
                                                                  (int)X, (int)Y, xs, rev, tmap, use_tmap);          \
  } while (0)
  if (Z == 32) LGM_ARING(1, 1, 1, false);
  else if (Z == 64) LGM_ARING(2, 1, 1, false);
  else if (Z == 128) LGM_ARING(4, 1, 1, false);
  else if (wide256) LGM_ARING(4, 2, 1, false);
  else if (use_tmap) LGM_ARING(4, 1, 2, true);
  else LGM_ARING(4, 1, 2, false);
#undef LGM_ARING
  count_launch("Ad_star", s);
  return finish(s, "lgm_Ad_star_fwd(ring)");
}

}  // namespace lgm
