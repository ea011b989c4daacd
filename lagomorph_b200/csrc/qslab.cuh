// qslab.cuh -- "quarter slab" FFT passes for (y, z) planes too large for one SM's shared memory
// (fp32, Y = Z = 256: the 256 x 129 spectrum slab of one (n, c, x) is 265 KB).
//
// The Y transform is split by decimation in time, Y = 4 * YQ:
//     F[k + YQ r] = sum_q  w4^{rq} w_Y^{kq}  S_q[k],      S_q[k] = sum_j f[4j + q] w_YQ^{jk}
// * qslab_fwd : one CTA per (n, c, x, q): the YQ real lines y = 4j + q (1 KB each, every fourth line of
//               the slab) -> Z real transform -> YQ-point Y transform -> S_q as a [q][kpos][kz] block of
//               the spectrum workspace. The tile is 129 x 65 words = 67 KB (3 CTAs per SM), exactly the
//               shape economy of the single-CTA slab kernels at 128^2.
// * xpassq    : the X pass absorbs the missing radix-4 Y stage. Its tile holds, for 8 consecutive
//               (kpos, kz) columns, all four q: [X][4][8]. X is transformed as (X/8) x 8: the outer stage
//               loads straight from global memory (a 32-point register DFT at X = 256), the inner stage
//               takes 8 rows x 4 q per thread: 8-point X DFTs, twiddle w_Y^{kq}, 4-point DFT over q
//               -> the true frequencies (kx, k + YQ r, kz) -> multiplier -> the whole way back, all in
//               registers. Shared-memory round trips per element: 2, as in the plain X pass.
// * qslab_inv : mirror of qslab_fwd.
// Replaces the 4-CTA cluster slab kernels (cslab_*: all-to-all through distributed shared memory, bound by
// cluster barriers and remote latency: 0.41-0.45 of the HBM peak) for beta == 0.
// Reference op being replaced: lagomorph/metric.py:11-19 + cuda/metric.cu:220-306.
#pragma once
#include "fft.cuh"

namespace lgm {

#ifndef LGM_PF_QSLAB
#define LGM_PF_QSLAB 148  /* CTAs ahead whose input is prefetched into L2 (0 = off) */
#endif
#ifndef LGM_PF_XPASSQ
#define LGM_PF_XPASSQ 111  /* X pass: tile of the CTA this many positions ahead prefetched into L2 (0 = off) */
#endif
#ifndef LGM_XPASSQ_MINBLOCKS
#define LGM_XPASSQ_MINBLOCKS 2
#endif

#ifndef LGM_XPASSQ_THREADS
#define LGM_XPASSQ_THREADS 256
#endif
constexpr int kQsThreads = 256;
constexpr int kQxThreads = LGM_XPASSQ_THREADS;
constexpr int kQxT = 8;               // (kpos, kz) columns per X-pass tile
constexpr int kQxL = 4 * kQxT;        // tile row length in words: [q][w]
constexpr int kQxP = kQxL + 1;        // tile row pitch (odd: the inner stage is bank-conflict free)

__device__ __forceinline__ void qs_l2_prefetch(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int Y, int Z>
__global__ void __launch_bounds__(kQsThreads, 3)
qslab_fwd_kernel(float2* __restrict__ spec, const float* __restrict__ in, const float2* __restrict__ twz_g,
                 const float2* __restrict__ twy_g, int rev) {
  constexpr int YQ = Y / 4, M = Z / 2, ZC = M + 1, P = YQ + 1;
  static_assert(YQ % 32 == 0, "quarter slab: the Z edge stage needs a multiple of 32 lines");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);  // ZC x P
  float2* twz = tile + ZC * P;                         // Z entries
  float2* twM = twz + Z;                               // M entries
  float2* twy = twM + M;                               // YQ entries e^{-2 pi i j / YQ}
  const int tid = threadIdx.x;
  const unsigned b = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const size_t slab = b >> 2;
  const unsigned q = b & 3;
  if (LGM_PF_QSLAB > 0 && tid == 0 && blockIdx.x + LGM_PF_QSLAB < gridDim.x) {
    // the four CTAs of a slab prefetch the four quarters of the slab LGM_PF_QSLAB/4 positions ahead
    const long long ns = rev ? (long long)slab - LGM_PF_QSLAB / 4 : (long long)slab + LGM_PF_QSLAB / 4;
    qs_l2_prefetch(in + ((size_t)ns * Y + q * YQ) * Z, (unsigned)(YQ * Z * sizeof(float)));
  }
  for (int j = tid; j < Z; j += kQsThreads) twz[j] = twz_g[j];
  for (int j = tid; j < M; j += kQsThreads) twM[j] = twz_g[2 * j];
  for (int j = tid; j < YQ; j += kQsThreads) twy[j] = twy_g[4 * j];
  __syncthreads();
  // Z: lines y = 4j + q, first radix stage straight from global memory, real split fused into the last
  real_fft_fwd_g<float, M, YQ>(in + (slab * Y + q) * Z, tile, P, twM, twz, tid, kQsThreads, 4 * M);
  __syncthreads();
  // YQ-point transform over j; its last stage stores straight to the block [q][kpos][kz]
  GSide<float2> gout{spec + (slab * Y + q * YQ) * ZC, ZC, ZC};
  ColFFT<float, YQ, YQ, 0, ZC>::template fwd_g<false, true>(tile, 1, P, twy, tid, kQsThreads, gout, gout);
}

template <int Y, int Z, bool POST>
__global__ void __launch_bounds__(kQsThreads, 3)
qslab_inv_kernel(float* __restrict__ out, const float2* __restrict__ spec, const float2* __restrict__ twz_g,
                 const float2* __restrict__ twy_g, int rev, float post) {
  constexpr int YQ = Y / 4, M = Z / 2, ZC = M + 1, P = YQ + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* twz = tile + ZC * P;
  float2* twM = twz + Z;
  float2* twy = twM + M;
  const int tid = threadIdx.x;
  const unsigned b = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const size_t slab = b >> 2;
  const unsigned q = b & 3;
  if (LGM_PF_QSLAB > 0 && tid == 0 && blockIdx.x + LGM_PF_QSLAB < gridDim.x) {
    const long long nb = rev ? (long long)b - LGM_PF_QSLAB : (long long)b + LGM_PF_QSLAB;
    qs_l2_prefetch(spec + (size_t)nb * YQ * ZC, (unsigned)(YQ * ZC * sizeof(float2)));
  }
  for (int j = tid; j < Z; j += kQsThreads) twz[j] = twz_g[j];
  for (int j = tid; j < M; j += kQsThreads) twM[j] = twz_g[2 * j];
  for (int j = tid; j < YQ; j += kQsThreads) twy[j] = twy_g[4 * j];
  __syncthreads();
  GSide<float2> gin{const_cast<float2*>(spec) + (slab * Y + q * YQ) * ZC, ZC, ZC};
  ColFFT<float, YQ, YQ, 0, ZC>::template inv_g<true, false>(tile, 1, P, twy, tid, kQsThreads, gin, gin);
  __syncthreads();
  real_fft_inv_g<float, M, YQ, POST>(out + (slab * Y + q) * Z, tile, P, twM, twz, tid, kQsThreads, 4 * M, post);
}

// 4-point DFT over q (forward: w4 = -i) and its inverse, unnormalised
__device__ __forceinline__ void dft4_fwd(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1.x = d02.x + d13.y; a1.y = d02.y - d13.x;   // d02 - i d13
  a3.x = d02.x - d13.y; a3.y = d02.y + d13.x;   // d02 + i d13
}
__device__ __forceinline__ void dft4_inv(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1.x = d02.x - d13.y; a1.y = d02.y + d13.x;   // d02 + i d13
  a3.x = d02.x + d13.y; a3.y = d02.y - d13.x;   // d02 - i d13
}

// X pass of the quarter-slab path, beta == 0 (scalar symbol, channels independent).
//   spec : (N*3, NX, 4, YQ, Zc) words; grid = (YQ*Zc / 8, N*3)
//   lxq  : 2(1-cos) of the X frequency held by tile row rho = k1*8 + k2  (kx = k1 + (NX/8) k2)
//   wyq  : [r][kpos] 2(1-cos) of the Y frequency fft_freq<YQ>(kpos) + YQ r
//   wlz  : Z LUT in the storage order of the real transform
template <int NX, int YQ, bool INVERSE, bool SAFE = false>
__global__ void __launch_bounds__(kQxThreads, LGM_XPASSQ_MINBLOCKS)
xpassq_kernel(float2* __restrict__ spec, int Zc, const float2* __restrict__ twx_g, const float2* __restrict__ twy_g,
              const float* __restrict__ lxq, const float* __restrict__ wyq, const float* __restrict__ wlz,
              double alpha, double gamma, float scale, int rev) {
  using C = float2;
  constexpr int T = kQxT, L = kQxL, P = kQxP, SUB = 8, RAD0 = NX / SUB, B0 = ilog2(RAD0);
  static_assert(RAD0 >= 2 && RAD0 <= 32, "X pass of the quarter-slab path: 16 <= NX <= 256");
  const unsigned bx = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const unsigned by = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);   // NX x P
  C* tw = tile + NX * P;                      // NX entries e^{-2 pi i j / NX}
  float* lx = reinterpret_cast<float*>(tw + NX);
  const int tid = threadIdx.x;
  for (int j = tid; j < NX; j += kQxThreads) {
    tw[j] = twx_g[j];
    lx[j] = lxq[j];
  }
  const long long QS = (long long)YQ * Zc;     // words per quarter block
  const long long plane = 4 * QS;              // words per x row
  const long long q0 = (long long)bx * T;
  C* base = spec + (long long)by * NX * plane + q0;
  if (LGM_PF_XPASSQ > 0) {  // the NX x 4 pieces of 64 bytes of a later CTA's tile: 4 prefetches per thread
    const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + LGM_PF_XPASSQ;
    if (lin < (long long)gridDim.x * gridDim.y) {
      const unsigned pbx0 = (unsigned)(lin % gridDim.x), pby0 = (unsigned)(lin / gridDim.x);
      const unsigned pbx = rev ? gridDim.x - 1 - pbx0 : pbx0, pby = rev ? gridDim.y - 1 - pby0 : pby0;
      const C* pb = spec + (long long)pby * NX * plane + (long long)pbx * T;
#ifndef LGM_PF_XPASSQ_SECT
#define LGM_PF_XPASSQ_SECT 1
#endif
      for (int pi = tid; pi < NX * 4 * LGM_PF_XPASSQ_SECT; pi += kQxThreads) {
        const int pc = pi / LGM_PF_XPASSQ_SECT, sc = pi % LGM_PF_XPASSQ_SECT;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + (long long)(pc >> 2) * plane + (pc & 3) * QS + sc * (T / LGM_PF_XPASSQ_SECT)));
      }
    }
  }
  __syncthreads();
  // outer forward stage: global -> registers (RAD0-point DFT over x = rest + 8 n) -> tile
  for (int it = tid; it < L * SUB; it += kQxThreads) {
    const int l = it % L, rest = it / L;
    const C* gp = base + (long long)rest * plane + (l / T) * QS + (l % T);
    C* p = tile + rest * P + l;
    C x[RAD0];
#pragma unroll
    for (int n = 0; n < RAD0; ++n) x[n] = gp[(long long)n * SUB * plane];
    reg_fft<RAD0, false>(x);
#pragma unroll
    for (int i = 0; i < RAD0; ++i) {
      const int kk = bitrev(i, B0);
      C v = x[i];
      if (kk != 0) v = cmul(v, tw[rest * kk]);
      p[kk * SUB * P] = v;
    }
  }
  // per-thread column of the inner stage: (kpos, kz) -> Y twiddles w_Y^{kq} and the (y, z) symbol part
  // (looked up here, not at the top: nothing of it has to live through the 32-point outer stage)
  const int w = tid % T;
  const int kpos = (int)((q0 + w) / Zc), pz = (int)((q0 + w) - (long long)kpos * Zc);
  const int k = fft_freq<YQ>(kpos);
  const C t1 = twy_g[k], t2 = twy_g[2 * k], t3 = twy_g[3 * k];
  const float wz = wlz[pz];
  float wy[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) wy[r] = wyq[r * YQ + kpos];
  __syncthreads();
  // inner stage: 8 rows x 4 q per thread
  for (int it = tid; it < T * (NX / SUB); it += kQxThreads) {
    const int blk = it / T;   // it % T == w
    C* p = tile + blk * SUB * P + w;
    C v[4][SUB];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int n = 0; n < SUB; ++n) v[q][n] = p[n * P + q * T];
      reg_fft<SUB, false>(v[q]);   // v[q][i] = X frequency row blk*8 + bitrev(i)
    }
    C y[4][SUB];
#pragma unroll
    for (int i = 0; i < SUB; ++i) {
      const int k2 = bitrev(i, 3);
      C a[4] = {v[0][i], cmul(v[1][i], t1), cmul(v[2][i], t2), cmul(v[3][i], t3)};
      dft4_fwd(a[0], a[1], a[2], a[3]);   // a[r] = F[kx, k + YQ r, kz]
      const float lxr = lx[blk * SUB + k2];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float sw = lxr + wy[r] + wz;
        const float lambda = (float)(gamma + alpha * (double)sw);
        const float Lm = lambda * lambda;
        C u = a[r];
        if (INVERSE) {
          const float f = oo_lambda_fast<float, SAFE>(lambda, Lm);
          u.x = ((u.x * f) * f) * scale;
          u.y = ((u.y * f) * f) * scale;
        } else {
          u.x = (Lm * u.x) * scale;
          u.y = (Lm * u.y) * scale;
        }
        a[r] = u;
      }
      dft4_inv(a[0], a[1], a[2], a[3]);
      y[0][k2] = a[0];
      y[1][k2] = cmulc(a[1], t1);
      y[2][k2] = cmulc(a[2], t2);
      y[3][k2] = cmulc(a[3], t3);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      reg_fft<SUB, true>(y[q]);
#pragma unroll
      for (int i = 0; i < SUB; ++i) p[bitrev(i, 3) * P + q * T] = y[q][i];
    }
  }
  __syncthreads();
  // outer inverse stage: tile -> registers -> global. (The base pointer is laundered: without it the
  // compiler keeps the 32 addresses of the first stage alive -- in local memory -- for these stores.)
  C* base2 = base;
  asm volatile("" : "+l"(base2));
  for (int it = tid; it < L * SUB; it += kQxThreads) {
    const int l = it % L, rest = it / L;
    C* gp = base2 + (long long)rest * plane + (l / T) * QS + (l % T);
    const C* p = tile + rest * P + l;
    C x[RAD0];
#pragma unroll
    for (int kk = 0; kk < RAD0; ++kk) {
      C v = p[kk * SUB * P];
      if (kk != 0) v = cmulc(v, tw[rest * kk]);
      x[kk] = v;
    }
    reg_fft<RAD0, true>(x);
#pragma unroll
    for (int i = 0; i < RAD0; ++i) gp[(long long)bitrev(i, B0) * SUB * plane] = x[i];
  }
}

}  // namespace lgm
