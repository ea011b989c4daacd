// fluid.cu -- FluidMetric sharp/flat: batched real 2-D/3-D FFT passes with the
// (alpha*lap + beta*grad div + gamma)^2 Fourier multiplier fused into the middle pass.
//
// Replaces, for the reference, torch.rfft -> lagomorph_ext.fluid_operator ->
// torch.irfft (lagomorph/metric.py:11-19, cuda/metric.cu:162-355): no cuFFT, no
// separate multiplier launch, no normalisation passes.
//
// Pass structure (3-D; 2-D drops the Y pass). Spectrum = half spectrum along the
// last axis, Zc = Z/2+1 complex per line, kept in the digit-reversed storage
// order the in-place FFT stages leave behind (fft.cuh):
//   Z-fwd : real lines -> half-length complex FFT + split -> spectrum lines
//   Y     : in-place column FFT on [Y x T] tiles (T contiguous spectrum words)
//   X     : [X x T] tiles: forward FFT -> multiplier (all `dim` channels of one
//           frequency together when beta != 0) -> inverse FFT, in place
//   Y-inv, Z-inv : mirrors.
// Sizes that are not powers of two (the reference's own tests use 3) take a
// direct-DFT path with the same semantics (unitary scaling in the transforms,
// multiplier on the natural-order spectrum).
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include <cooperative_groups.h>
#include "fft.cuh"
#include "mixfft.cuh"

namespace lgm {

// ------------------------------------------------------------------------------------------
// Fourier multiplier, arithmetic as cuda/metric.cu:162-306 (double alpha/beta/gamma, symbol
// squared, Cholesky solve with safe_sqrt for sharp, plain multiply for flat).
// ------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ R safe_sqrt(R x) {  // cuda/metric.cu:14-18
  if ((double)x < 1e-8) return (R)1e-4;
  return sqrt(x);
}
template <typename R>
__device__ __forceinline__ R oo_sqrt(R x) {  // 1./safe_sqrt(x), rounded to Real
  return (R)(1. / (double)safe_sqrt(x));
}

template <typename R, int D>
struct Symbol {  // per-frequency operator, built once, applied to re and im parts of every sample
  R L00, L10, L11, L20, L21, L22;
  R ooG00, G10, ooG11, G20, G21, ooG22;
};

template <typename R, int D, bool INVERSE>
__device__ __forceinline__ Symbol<R, D> make_symbol(const R (&w)[3], const R (&s)[3], double alpha,
                                                    double beta, double gamma) {
  Symbol<R, D> S;
  if constexpr (D == 2) {
    const R lambda = (R)(gamma + alpha * (double)(w[0] + w[1]));
    R l00 = (R)((double)lambda - beta * (double)w[0]);
    R l11 = (R)((double)lambda - beta * (double)w[1]);
    R l10 = (R)(beta * (double)s[0] * (double)s[1]);
    S.L00 = l00 * l00 + l10 * l10;
    S.L10 = l00 * l10 + l10 * l11;
    S.L11 = l11 * l11 + l10 * l10;
    if (INVERSE) {
      S.ooG00 = oo_sqrt<R>(S.L00);
      S.G10 = S.L10 * S.ooG00;
      S.ooG11 = oo_sqrt<R>(S.L11 - S.G10 * S.G10);
    }
  } else {
    const R lambda = (R)(gamma + alpha * (double)(w[0] + w[1] + w[2]));
    R l00 = (R)((double)lambda - beta * (double)w[0]);
    R l11 = (R)((double)lambda - beta * (double)w[1]);
    R l22 = (R)((double)lambda - beta * (double)w[2]);
    R l10 = (R)(beta * (double)s[0] * (double)s[1]);
    R l20 = (R)(beta * (double)s[0] * (double)s[2]);
    R l21 = (R)(beta * (double)s[1] * (double)s[2]);
    S.L00 = l00 * l00 + l10 * l10 + l20 * l20;
    S.L10 = l00 * l10 + l10 * l11 + l20 * l21;
    S.L11 = l10 * l10 + l11 * l11 + l21 * l21;
    S.L20 = l00 * l20 + l10 * l21 + l20 * l22;
    S.L21 = l10 * l20 + l11 * l21 + l21 * l22;
    S.L22 = l20 * l20 + l21 * l21 + l22 * l22;
    if (INVERSE) {
      S.ooG00 = oo_sqrt<R>(S.L00);
      S.G10 = S.L10 * S.ooG00;
      S.G20 = S.L20 * S.ooG00;
      S.ooG11 = oo_sqrt<R>(S.L11 - S.G10 * S.G10);
      S.G21 = (S.L21 - S.G20 * S.G10) * S.ooG11;
      S.ooG22 = oo_sqrt<R>(S.L22 - S.G20 * S.G20 - S.G21 * S.G21);
    }
  }
  return S;
}

template <typename R, int D, bool INVERSE>
__device__ __forceinline__ void apply_symbol(const Symbol<R, D>& S, R (&b)[D]) {
  if constexpr (D == 2) {
    if (INVERSE) {  // cuda/metric.cu:80-100
      R y0 = b[0] * S.ooG00;
      R y1 = (b[1] - S.G10 * y0) * S.ooG11;
      b[1] = y1 * S.ooG11;
      b[0] = (y0 - S.G10 * b[1]) * S.ooG00;
    } else {  // :132-144
      R x = S.L00 * b[0] + S.L10 * b[1];
      b[1] = S.L10 * b[0] + S.L11 * b[1];
      b[0] = x;
    }
  } else {
    if (INVERSE) {  // :102-130
      R y0 = b[0] * S.ooG00;
      R y1 = (b[1] - S.G10 * y0) * S.ooG11;
      R y2 = (b[2] - S.G20 * y0 - S.G21 * y1) * S.ooG22;
      b[2] = y2 * S.ooG22;
      b[1] = (y1 - S.G21 * b[2]) * S.ooG11;
      b[0] = (y0 - S.G10 * b[1] - S.G20 * b[2]) * S.ooG00;
    } else {  // :146-160
      R x = S.L00 * b[0] + S.L10 * b[1] + S.L20 * b[2];
      R y = S.L10 * b[0] + S.L11 * b[1] + S.L21 * b[2];
      b[2] = S.L20 * b[0] + S.L21 * b[1] + S.L22 * b[2];
      b[0] = x;
      b[1] = y;
    }
  }
}

// beta == 0: the symbol is lambda^2 * Identity. Returns the factor pair so that
//   sharp: v = (v*f)*f with f = 1/sqrt(lambda^2)   (what the Cholesky path reduces to)
//   flat : v = f*v     with f = lambda^2
template <typename R, int D, bool INVERSE>
__device__ __forceinline__ R scalar_symbol(const R (&w)[3], double alpha, double gamma) {
  R sw = (D == 2) ? (w[0] + w[1]) : (w[0] + w[1] + w[2]);
  const R lambda = (R)(gamma + alpha * (double)sw);
  const R L = lambda * lambda;
  return INVERSE ? oo_sqrt<R>(L) : L;
}

// Standalone multiplier on a natural-order interleaved half spectrum (N, D, X, Y[, Zc], 2):
// the reference's fluid_operator boundary, also the middle step of the direct-DFT path.
template <typename R, int D, bool INVERSE>
__global__ void __launch_bounds__(256)
fluid_operator_kernel(R* __restrict__ Fm, const R* __restrict__ c0, const R* __restrict__ s0,
                      const R* __restrict__ c1, const R* __restrict__ s1, const R* __restrict__ c2,
                      const R* __restrict__ s2, double alpha, double beta, double gamma, int N,
                      Geom<D> g) {
  const long long fid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (fid >= g.V) return;
  int pos[D];
  decode<D>(fid, g, pos);
  R w[3] = {c0[pos[0]], c1[pos[1]], D == 3 ? c2[pos[D - 1]] : R(0)};
  R s[3] = {s0[pos[0]], s1[pos[1]], D == 3 ? s2[pos[D - 1]] : R(0)};
  Symbol<R, D> S = make_symbol<R, D, INVERSE>(w, s, alpha, beta, gamma);
  for (int n = 0; n < N; ++n) {
    R* base = Fm + 2 * ((long long)n * D * g.V + fid);
    R re[D], im[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      re[c] = base[2 * c * g.V];
      im[c] = base[2 * c * g.V + 1];
    }
    apply_symbol<R, D, INVERSE>(S, re);
    apply_symbol<R, D, INVERSE>(S, im);
#pragma unroll
    for (int c = 0; c < D; ++c) {
      base[2 * c * g.V] = re[c];
      base[2 * c * g.V + 1] = im[c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Plan: device LUTs (storage order) and twiddle tables, cached per (device, dtype, dim, shape).
// ------------------------------------------------------------------------------------------
struct FluidPlan {
  bool fast = false;
  int dim = 0;
  int n[3] = {1, 1, 1};
  int rows[3] = {1, 1, 1};   // spectrum extent per axis (last axis: n/2+1)
  void* wl[3] = {nullptr, nullptr, nullptr};  // 2(1-cos) LUT, storage order
  void* sl[3] = {nullptr, nullptr, nullptr};  // sin LUT, storage order
  void* tw[3] = {nullptr, nullptr, nullptr};  // e^{-2 pi i j/n}, n entries (complex)
  // quarter-slab path (qslab.cuh): X LUT in the (X/8) x 8 storage order, Y LUT as [r][kpos]
  bool qslab = false;
  void* q_lx = nullptr;
  void* q_wy = nullptr;
};

// shapes served by the quarter-slab kernels (fp32, beta == 0 decided per call)
template <typename R>
static bool qslab_shape(int dim, const int64_t* shape) {
  return sizeof(R) == 4 && dim == 3 && shape[1] == 256 && shape[2] == 256 &&
         (shape[0] == 64 || shape[0] == 128 || shape[0] == 256);
}

static bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }

template <typename R>
static bool fast_ok(int dim, const int64_t* shape) {
  const int maxn = sizeof(R) == 4 ? 512 : 256;
  for (int a = 0; a < dim; ++a) {
    long long n = shape[a];
    if (!is_pow2(n)) return false;
    if (a == dim - 1) {
      if (n < 16 || n > 2 * maxn) return false;  // half-length complex FFT of n/2 >= 8 points
    } else if (n < 8 || n > maxn) return false;
  }
  return true;
}

static std::mutex g_plan_mu;
static std::map<std::tuple<int, int, int, long long, long long, long long>, FluidPlan> g_plans;

template <typename R>
static int get_plan(int dim, const int64_t* shape, cudaStream_t s, const FluidPlan** out) {
  int dev = 0;
  cudaGetDevice(&dev);
  auto key = std::make_tuple(dev, (int)sizeof(R), dim, (long long)shape[0], (long long)shape[1],
                             dim == 3 ? (long long)shape[2] : 0LL);
  std::lock_guard<std::mutex> lk(g_plan_mu);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) {
    *out = &it->second;
    return LGM_OK;
  }
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &st);
  if (st != cudaStreamCaptureStatusNone)
    return set_error(LGM_EUNSUP, "lgm_fluid_apply: first call for a shape builds its tables and cannot be stream-captured; call once outside capture");
  FluidPlan p;
  p.dim = dim;
  p.fast = fast_ok<R>(dim, shape);
  using C = typename Cx<R>::T;
  for (int a = 0; a < dim; ++a) {
    const int n = (int)shape[a];
    const bool last = (a == dim - 1);
    p.n[a] = n;
    p.rows[a] = last ? n / 2 + 1 : n;
    std::vector<R> wl(p.rows[a]), sl(p.rows[a]);
    // frequency held by each storage row (digit-reversed on the fast path)
    std::vector<int> freq(p.rows[a]);
    for (int pos = 0; pos < p.rows[a]; ++pos) freq[pos] = pos;
    if (p.fast) {
      const int m = last ? n / 2 : n;
      for (int k = 0; k < m; ++k) freq[fft_pos_rt(m, k)] = k;
      if (last) freq[m] = m;
    }
    for (int pos = 0; pos < p.rows[a]; ++pos) {
      // lagomorph/metric.py:65-75: float64 numpy -> torch.Tensor (float32!) -> .type(dtype)
      const double ang = 2.0 * M_PI * (double)freq[pos] / (double)n;
      wl[pos] = (R)(float)(2.0 * (1.0 - cos(ang)));
      sl[pos] = (R)(float)sin(ang);
    }
    std::vector<C> tw(n);
    for (int j = 0; j < n; ++j) {
      // exact octant symmetry is not needed; cos/sin of the reduced angle in double
      const double ang = 2.0 * M_PI * (double)j / (double)n;
      tw[j].x = (R)cos(ang);
      tw[j].y = (R)(-sin(ang));
    }
    cudaError_t e;
    if ((e = cudaMalloc(&p.wl[a], sizeof(R) * wl.size())) != cudaSuccess ||
        (e = cudaMalloc(&p.sl[a], sizeof(R) * sl.size())) != cudaSuccess ||
        (e = cudaMalloc(&p.tw[a], sizeof(C) * tw.size())) != cudaSuccess)
      return set_error((int)e, "lgm_fluid_apply: table allocation failed: %s", cudaGetErrorString(e));
    cudaMemcpy(p.wl[a], wl.data(), sizeof(R) * wl.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.sl[a], sl.data(), sizeof(R) * sl.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.tw[a], tw.data(), sizeof(C) * tw.size(), cudaMemcpyHostToDevice);
  }
  if (p.fast && qslab_shape<R>(dim, shape)) {
    const int NX = (int)shape[0], NY = (int)shape[1], YQ = NY / 4, R0 = NX / 8;
    std::vector<float> lx(NX), wy(4 * YQ);
    for (int rho = 0; rho < NX; ++rho) {
      const int kx = rho / 8 + R0 * (rho % 8);
      lx[rho] = (float)(2.0 * (1.0 - cos(2.0 * M_PI * (double)kx / (double)NX)));
    }
    for (int kk = 0; kk < YQ; ++kk)
      for (int r = 0; r < 4; ++r)
        wy[r * YQ + fft_pos_rt(YQ, kk)] = (float)(2.0 * (1.0 - cos(2.0 * M_PI * (double)(kk + YQ * r) / (double)NY)));
    cudaError_t e;
    if ((e = cudaMalloc(&p.q_lx, sizeof(float) * lx.size())) != cudaSuccess ||
        (e = cudaMalloc(&p.q_wy, sizeof(float) * wy.size())) != cudaSuccess)
      return set_error((int)e, "lgm_fluid_apply: table allocation failed: %s", cudaGetErrorString(e));
    cudaMemcpy(p.q_lx, lx.data(), sizeof(float) * lx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.q_wy, wy.data(), sizeof(float) * wy.size(), cudaMemcpyHostToDevice);
    p.qslab = true;
  }
  auto ins = g_plans.emplace(key, p);
  *out = &ins.first->second;
  return LGM_OK;
}

// ------------------------------------------------------------------------------------------
// Fast path kernels
// ------------------------------------------------------------------------------------------
// L2 prefetch of the input of a LATER CTA (bulk prefetch, no shared memory, one instruction issued
// by one thread): the slab kernels start with a burst of global loads whose DRAM latency only the
// other CTAs of the SM hide; pulling the slab that the CTA a few hundred positions ahead
// will read into L2 now turns those loads into L2 hits. Distances (in CTAs, 0 = off) per kernel:
// measured on B200 (profiles/r2_notes.md): slab_fwd / slab_inv 0.190 -> 0.172 ms at C2 with 74..148,
// cslab_fwd 1.09 -> 1.02 and cslab_inv 1.21 -> 1.06 ms at C3 with 148..296; the X pass gets slower (off).
#ifndef LGM_PF_SLAB_FWD
#define LGM_PF_SLAB_FWD 111
#endif
#ifndef LGM_PF_SLAB_INV
#define LGM_PF_SLAB_INV 111
#endif
#ifndef LGM_PF_CSLAB_FWD
#define LGM_PF_CSLAB_FWD 222
#endif
#ifndef LGM_PF_CSLAB_INV
#define LGM_PF_CSLAB_INV 222
#endif
#ifndef LGM_PF_XPASS
#define LGM_PF_XPASS 0
#endif
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

#ifndef LGM_SLAB_IO_UNROLL
#define LGM_SLAB_IO_UNROLL 8  /* loads in flight per thread in slab_fwd's fill loop: 4 -> 8 = 0.244 -> 0.224 ms */
#endif
#ifndef LGM_FFT_THREADS
#define LGM_FFT_THREADS 256
#endif
constexpr int kFftThreads = LGM_FFT_THREADS;
constexpr int kSlabIoUnroll = LGM_SLAB_IO_UNROLL;
#ifndef LGM_XPASS_MULTI_RADMAX
#define LGM_XPASS_MULTI_RADMAX 8  /* largest innermost radix fused in registers for beta != 0 (3 channels) */
#endif
#ifndef LGM_XPASS_TX32_MAX
#define LGM_XPASS_TX32_MAX 128  /* largest NCH*NX whose X-pass tile is 32 words wide (else 16) */
#endif
#ifndef LGM_XPASS_MINBLOCKS
#define LGM_XPASS_MINBLOCKS 4  /* 64 registers: measured 0.254 -> 0.240 ms per C2 X pass vs 80 registers */
#endif

// Z transforms whose outermost radix stage works on global memory (fft.cuh zedge_stage): needs two
// radix stages and a multiple of 32 lines per CTA.
#ifndef LGM_ZEDGE
#define LGM_ZEDGE 1
#endif
template <int M, int L>
constexpr bool kZEdge = LGM_ZEDGE && ((ilog2(M) + 3) / 4 >= 2) && (L % 32 == 0);

// Z forward: L real lines of Z points -> L spectrum lines of Z/2+1 words.
template <typename R, int Z, int L>
__global__ void __launch_bounds__(kFftThreads)
zfwd_kernel(typename Cx<R>::T* __restrict__ spec, const R* __restrict__ in, long long rows_total,
            const typename Cx<R>::T* __restrict__ tw_g) {
  using C = typename Cx<R>::T;
  constexpr int M = Z / 2, P = L + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);  // (M+1) x P
  C* tw = tile + (M + 1) * P;                // Z entries
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * L;
  for (int j = tid; j < Z; j += kFftThreads) tw[j] = tw_g[j];
  C* twM = tw + Z;  // M entries W_M^j = W_Z^{2j}, compacted (fft_stage indexes tw[j*(M/BLOCK)])
  for (int j = tid; j < M; j += kFftThreads) twM[j] = tw_g[2 * j];
  bool edge = false;
  if constexpr (kZEdge<M, L>) edge = (row0 + L <= rows_total);
  if (edge) {
    if constexpr (kZEdge<M, L>) {
      __syncthreads();
      real_fft_fwd_g<R, M, L>(in + row0 * Z, tile, P, twM, tw, tid, kFftThreads);
    }
  } else {
    const C* in2 = reinterpret_cast<const C*>(in);
#pragma unroll kSlabIoUnroll
    for (int idx = tid; idx < L * M; idx += kFftThreads) {
      const int l = idx / M, j = idx % M;
      C v;
      v.x = v.y = R(0);
      if (row0 + l < rows_total) v = in2[(row0 + l) * M + j];
      tile[j * P + l] = v;
    }
    __syncthreads();
    // half-length complex FFT with the real split fused into its last radix stage (fft.cuh)
    real_fft_fwd<R, M, L>(tile, P, 1, twM, tw, tid, kFftThreads);
  }
  __syncthreads();
  for (int idx = tid; idx < L * (M + 1); idx += kFftThreads) {
    const int l = idx / (M + 1), p = idx % (M + 1);
    if (row0 + l < rows_total) spec[(row0 + l) * (M + 1) + p] = tile[p * P + l];
  }
}

// Z inverse: spectrum lines -> real lines (unnormalised; scaling lives in the multiplier).
template <typename R, int Z, int L>
__global__ void __launch_bounds__(kFftThreads)
zinv_kernel(R* __restrict__ out, const typename Cx<R>::T* __restrict__ spec, long long rows_total,
            const typename Cx<R>::T* __restrict__ tw_g) {
  using C = typename Cx<R>::T;
  constexpr int M = Z / 2, P = L + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);
  C* tw = tile + (M + 1) * P;
  C* twM = tw + Z;
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * L;
  for (int j = tid; j < Z; j += kFftThreads) tw[j] = tw_g[j];
  for (int j = tid; j < M; j += kFftThreads) twM[j] = tw_g[2 * j];
#pragma unroll kSlabIoUnroll
  for (int idx = tid; idx < L * (M + 1); idx += kFftThreads) {
    const int l = idx / (M + 1), p = idx % (M + 1);
    C v;
    v.x = v.y = R(0);
    if (row0 + l < rows_total) v = spec[(row0 + l) * (M + 1) + p];
    tile[p * P + l] = v;
  }
  __syncthreads();
  // (storing the last radix stage straight to global memory, as zfwd loads its first one, was
  // measured slower here: 64-byte store runs, 0.65 -> 0.69 ms at 256^3 x 8)
  // unsplit fused into the first radix stage of the inverse half-length FFT (fft.cuh)
  real_fft_inv<R, M, L>(tile, P, 1, twM, tw, tid, kFftThreads);
  __syncthreads();
  C* out2 = reinterpret_cast<C*>(out);
  for (int idx = tid; idx < L * M; idx += kFftThreads) {
    const int l = idx / M, j = idx % M;
    if (row0 + l < rows_total) out2[(row0 + l) * M + j] = tile[j * P + l];
  }
}

// Slab forward (3-D): one CTA owns one x-slab of one channel: Y real lines of Z points are
// staged in shared memory as tile[rz*P + y] (P = Y+1, odd), transformed along z (half-length
// complex FFT + split, lanes over y) and then along y (lanes over rz, stride P => conflict
// free), and leave as the spectrum slab [ry][rz]. Replaces zfwd + ypass (one global round trip
// of the spectrum less) whenever (Z/2+1)*(Y+1) complex words fit in shared memory.
template <typename R, int Y, int Z>
__global__ void __launch_bounds__(kFftThreads)
slab_fwd_kernel(typename Cx<R>::T* __restrict__ spec, const R* __restrict__ in,
                const typename Cx<R>::T* __restrict__ twz_g, const typename Cx<R>::T* __restrict__ twy_g, int rev) {
  const unsigned bx = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  using C = typename Cx<R>::T;
  constexpr int M = Z / 2, P = Y + 1, ZC = M + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);  // ZC x P
  C* twz = tile + ZC * P;                    // Z entries
  C* twM = twz + Z;                          // M entries (W_M^j = W_Z^2j)
  C* twy = twM + M;                          // Y entries
  const int tid = threadIdx.x;
  if (LGM_PF_SLAB_FWD > 0 && tid == 0 && blockIdx.x + LGM_PF_SLAB_FWD < gridDim.x) {
    const long long nb = rev ? (long long)bx - LGM_PF_SLAB_FWD : (long long)bx + LGM_PF_SLAB_FWD;
    l2_prefetch(in + (size_t)nb * Y * Z, (unsigned)(Y * Z * sizeof(R)));
  }
  for (int j = tid; j < Z; j += kFftThreads) twz[j] = twz_g[j];
  for (int j = tid; j < M; j += kFftThreads) twM[j] = twz_g[2 * j];
  for (int j = tid; j < Y; j += kFftThreads) twy[j] = twy_g[j];
  if constexpr (kZEdge<M, Y>) {
    // Z: first radix stage straight from global memory (no fill loop), then the fused split stage
    __syncthreads();
    real_fft_fwd_g<R, M, Y>(in + (size_t)bx * Y * Z, tile, P, twM, twz, tid, kFftThreads);
  } else {
    const C* in2 = reinterpret_cast<const C*>(in) + (size_t)bx * Y * M;
#pragma unroll kSlabIoUnroll
    for (int idx = tid; idx < Y * M; idx += kFftThreads) {
      const int y = idx / M, j = idx % M;
      tile[j * P + y] = in2[idx];
    }
    __syncthreads();
    real_fft_fwd<R, M, Y>(tile, P, 1, twM, twz, tid, kFftThreads);  // Z: half-length FFT + fused split
  }
  __syncthreads();
  // Y transform; its last stage stores straight to the spectrum slab [ry][rz] (lanes over rz)
  GSide<C> gout{spec + (size_t)bx * Y * ZC, ZC, ZC};
  ColFFT<R, Y, Y, 0, ZC>::template fwd_g<false, true>(tile, 1, P, twy, tid, kFftThreads, gout, gout);
}

// Slab inverse: mirror of slab_fwd_kernel (spectrum slab -> Y real lines), unnormalised.
template <typename R, int Y, int Z, bool POST>
__global__ void __launch_bounds__(kFftThreads)
slab_inv_kernel(R* __restrict__ out, const typename Cx<R>::T* __restrict__ spec,
                const typename Cx<R>::T* __restrict__ twz_g, const typename Cx<R>::T* __restrict__ twy_g, int rev,
                R post) {
  const unsigned bx = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  using C = typename Cx<R>::T;
  constexpr int M = Z / 2, P = Y + 1, ZC = M + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);
  C* twz = tile + ZC * P;
  C* twM = twz + Z;
  C* twy = twM + M;
  const int tid = threadIdx.x;
  if (LGM_PF_SLAB_INV > 0 && tid == 0 && blockIdx.x + LGM_PF_SLAB_INV < gridDim.x) {
    const long long nb = rev ? (long long)bx - LGM_PF_SLAB_INV : (long long)bx + LGM_PF_SLAB_INV;
    l2_prefetch(spec + (size_t)nb * Y * ZC, (unsigned)(Y * ZC * sizeof(C)));
  }
  for (int j = tid; j < Z; j += kFftThreads) twz[j] = twz_g[j];
  for (int j = tid; j < M; j += kFftThreads) twM[j] = twz_g[2 * j];
  for (int j = tid; j < Y; j += kFftThreads) twy[j] = twy_g[j];
  __syncthreads();
  // inverse Y transform; its first stage loads straight from the spectrum slab [ry][rz]
  GSide<C> gin{const_cast<C*>(spec) + (size_t)bx * Y * ZC, ZC, ZC};
  ColFFT<R, Y, Y, 0, ZC>::template inv_g<true, false>(tile, 1, P, twy, tid, kFftThreads, gin, gin);
  __syncthreads();
  if constexpr (kZEdge<M, Y>) {
    // fused unsplit stage, then the last radix stage stores straight to global memory (no drain loop)
    real_fft_inv_g<R, M, Y, POST>(out + (size_t)bx * Y * Z, tile, P, twM, twz, tid, kFftThreads, M, post);
  } else {
    real_fft_inv<R, M, Y>(tile, P, 1, twM, twz, tid, kFftThreads);  // fused unsplit + inverse half-length FFT
    __syncthreads();
    C* o2 = reinterpret_cast<C*>(out) + (size_t)bx * Y * M;
#pragma unroll 4
    for (int idx = tid; idx < Y * M; idx += kFftThreads) {
      const int y = idx / M, j = idx % M;
      C v = tile[j * P + y];
      if (POST) {
        v.x = post_scale(post, v.x);
        v.y = post_scale(post, v.y);
      }
      o2[idx] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Cluster slab (fp32, Y = Z = 256): the 256 x 129 spectrum slab of one (n, c, x) is 265 KB, more than
// one SM's shared memory, so a thread-block CLUSTER of four CTAs owns it (69 KB each, 3 CTAs per SM
// like the single-CTA slab kernels) and the transposition between the Z and the Y transform is an
// all-to-all through distributed shared memory:
//   Z phase : CTA r holds the real lines y in [64r, 64r+64) as tile[row(p)*P + yl] and transforms
//             them along z. The 129 spectrum rows p are stored as four PANELS of 33 tile rows:
//             panel q = rows p in [32q, 32q+32) plus one spare row; the spare row of panel 0 holds the
//             Nyquist row p = 128.
//   swap    : panel q of CTA r <-> panel r of CTA q (DSMEM), after which CTA r owns the spectrum rows
//             of range r for ALL 256 lines: panel q now is (range r, y in [64q, 64q+64)).
//   Y phase : 256-point transforms along y for its 32 lines (33 in CTA 0: the Nyquist line), the
//             last stage stores straight to global memory.
// Replaces zfwd + ypass (and ypass + zinv) at 256^3: one global round trip of the spectrum less in
// each direction, 5 passes -> 3. (A two-CTA variant with 134 KB tiles, one CTA per SM, was measured
// slower than the unfused passes: nothing overlaps the load phase.)
// ------------------------------------------------------------------------------------------
constexpr int kCsNC = 4;              // CTAs per cluster
constexpr int kCsThreads = 256;
constexpr int kCsYL = 256 / kCsNC;    // real lines per CTA in the Z phase (64)
constexpr int kCsPR = 33;             // tile rows per panel
constexpr int kCsP = kCsYL + 1;       // row pitch in complex words (65, odd: conflict free both ways)
constexpr int kCsRows = kCsNC * kCsPR;
constexpr int kCsNyq = 32 * kCsP;     // Nyquist row = spare row of panel 0
constexpr int kCsJumpY = kCsPR * kCsP - kCsYL;  // Y phase: line y lives in panel y >> 6

// panel q of rank r <-> panel r of rank q; every (row, yl) element pair is moved by exactly one thread
// (rows of parity [r < q] by the lower rank, the others by the higher rank).
__device__ __forceinline__ void cslab_swap(float2* tile, unsigned rank, int tid) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
  for (unsigned d = 1; d < kCsNC; ++d) {
    const unsigned q = (rank + d) % kCsNC;
    float2* mine = tile + q * kCsPR * kCsP;
    float2* theirs = cluster.map_shared_rank(tile, q) + rank * kCsPR * kCsP;
    const int par = (rank < q) ? 0 : 1;
#pragma unroll 4
    for (int e = tid; e < 17 * kCsYL; e += kCsThreads) {
      const int i = 2 * (e / kCsYL) + par, yl = e % kCsYL;
      if (i < kCsPR) {
        const int o = i * kCsP + yl;
        const float2 a = mine[o], b = theirs[o];
        mine[o] = b;
        theirs[o] = a;
      }
    }
  }
}

template <bool INV, int L>
__device__ __forceinline__ void cslab_ypass(float2* tile, const float2* twy, int tid, GSide<float2> g) {
  constexpr int Y = 256;
  if (!INV) {
    fft_stage<float, Y, Y, 16, L, false, false, false, 6>(tile, 1, kCsP, twy, tid, kCsThreads, GSide<float2>(), kCsJumpY);
    __syncthreads();
    fft_stage<float, Y, 16, 16, L, false, false, true, 6>(tile, 1, kCsP, twy, tid, kCsThreads, g, kCsJumpY);
  } else {
    fft_stage<float, Y, 16, 16, L, true, true, false, 6>(tile, 1, kCsP, twy, tid, kCsThreads, g, kCsJumpY);
    __syncthreads();
    fft_stage<float, Y, Y, 16, L, true, false, false, 6>(tile, 1, kCsP, twy, tid, kCsThreads, GSide<float2>(), kCsJumpY);
  }
}

__global__ void __cluster_dims__(kCsNC, 1, 1) __launch_bounds__(kCsThreads, 3)
cslab_fwd_kernel(float2* __restrict__ spec, const float* __restrict__ in, const float2* __restrict__ twz_g,
                 const float2* __restrict__ twy_g, int rev) {
  namespace cg = cooperative_groups;
  constexpr int Y = 256, Z = 256, M = Z / 2, ZC = M + 1, P = kCsP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);  // kCsRows x P
  float2* twz = tile + kCsRows * P;                    // Z entries
  float2* twM = twz + Z;                               // M entries
  float2* twy = twM + M;                               // Y entries
  const int tid = threadIdx.x;
  const unsigned rank = cg::this_cluster().block_rank();
  const size_t slab = rev ? (gridDim.x - 1 - blockIdx.x) / kCsNC : blockIdx.x / kCsNC;
  for (int j = tid; j < Z; j += kCsThreads) twz[j] = twz_g[j];
  for (int j = tid; j < M; j += kCsThreads) twM[j] = twz_g[2 * j];
  for (int j = tid; j < Y; j += kCsThreads) twy[j] = twy_g[j];
  float2* in2 = reinterpret_cast<float2*>(const_cast<float*>(in)) + (slab * Y + rank * kCsYL) * M;
  if (LGM_PF_CSLAB_FWD > 0 && tid == 0 && blockIdx.x + LGM_PF_CSLAB_FWD < gridDim.x) {
    const long long d = LGM_PF_CSLAB_FWD / kCsNC;
    const long long ns = rev ? (long long)slab - d : (long long)slab + d;
    l2_prefetch(in + ((size_t)ns * Y + rank * kCsYL) * Z, (unsigned)(kCsYL * Z * sizeof(float)));
  }
  __syncthreads();
  // Z: half-length complex FFT (16 x 8) on panel rows; first stage straight from global memory, the
  // real split fused into the last one
  zedge_stage<float, M, 16, kCsYL, false, 5>(in2, tile, P, twM, tid, kCsThreads, P);
  __syncthreads();
  real_edge_stage<float, M, kCsYL, false, 5>(tile, P, 1, twz, tid, kCsThreads, P, kCsNyq);
  cg::this_cluster().sync();
  cslab_swap(tile, rank, tid);
  cg::this_cluster().sync();
  // Y transform over the four panels; the last stage stores straight to the spectrum slab [ry][rz]
  GSide<float2> gout{spec + slab * Y * ZC + rank * 32, ZC, rank == 0 ? 33 : 32, 32, M - 32};
  if (rank == 0) cslab_ypass<false, 33>(tile, twy, tid, gout);
  else cslab_ypass<false, 32>(tile, twy, tid, gout);
}

__global__ void __cluster_dims__(kCsNC, 1, 1) __launch_bounds__(kCsThreads, 3)
cslab_inv_kernel(float* __restrict__ out, const float2* __restrict__ spec, const float2* __restrict__ twz_g,
                 const float2* __restrict__ twy_g, int rev) {
  namespace cg = cooperative_groups;
  constexpr int Y = 256, Z = 256, M = Z / 2, ZC = M + 1, P = kCsP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  float2* twz = tile + kCsRows * P;
  float2* twM = twz + Z;
  float2* twy = twM + M;
  const int tid = threadIdx.x;
  const unsigned rank = cg::this_cluster().block_rank();
  const size_t slab = rev ? (gridDim.x - 1 - blockIdx.x) / kCsNC : blockIdx.x / kCsNC;
  for (int j = tid; j < Z; j += kCsThreads) twz[j] = twz_g[j];
  for (int j = tid; j < M; j += kCsThreads) twM[j] = twz_g[2 * j];
  for (int j = tid; j < Y; j += kCsThreads) twy[j] = twy_g[j];
  __syncthreads();
  if (LGM_PF_CSLAB_INV > 0 && tid == 0 && rank == 0 && blockIdx.x + LGM_PF_CSLAB_INV < gridDim.x) {
    const long long d = LGM_PF_CSLAB_INV / kCsNC;
    const long long ns = rev ? (long long)slab - d : (long long)slab + d;
    l2_prefetch(spec + (size_t)ns * Y * ZC, (unsigned)(Y * ZC * sizeof(float2)));
  }
  GSide<float2> gin{const_cast<float2*>(spec) + slab * Y * ZC + rank * 32, ZC, rank == 0 ? 33 : 32, 32, M - 32};
  if (rank == 0) cslab_ypass<true, 33>(tile, twy, tid, gin);
  else cslab_ypass<true, 32>(tile, twy, tid, gin);
  cg::this_cluster().sync();
  cslab_swap(tile, rank, tid);
  cg::this_cluster().sync();
  real_edge_stage<float, M, kCsYL, true, 5>(tile, P, 1, twz, tid, kCsThreads, P, kCsNyq);
  __syncthreads();
  float2* o2 = reinterpret_cast<float2*>(out) + (slab * Y + rank * kCsYL) * M;
  zedge_stage<float, M, 16, kCsYL, true, 5>(o2, tile, P, twM, tid, kCsThreads, P);
}

// Y pass (3-D only): in-place column FFT over the middle axis on [NY x T] tiles.
// grid = (ceil(Zc/T), X, N*dim)
template <typename R, int NY, int T, bool INV>
__global__ void __launch_bounds__(kFftThreads)
ypass_kernel(typename Cx<R>::T* __restrict__ spec, int X, int Zc,
             const typename Cx<R>::T* __restrict__ tw_g) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);  // NY x T
  C* tw = tile + NY * T;
  const int tid = threadIdx.x;
  for (int j = tid; j < NY; j += kFftThreads) tw[j] = tw_g[j];
  const int z0 = blockIdx.x * T;
  C* base = spec + (((long long)blockIdx.z * X + blockIdx.y) * NY) * Zc + z0;
  for (int idx = tid; idx < NY * T; idx += kFftThreads) {
    const int l = idx % T, r = idx / T;
    C v;
    v.x = v.y = R(0);
    if (z0 + l < Zc) v = base[(long long)r * Zc + l];
    tile[idx] = v;
  }
  __syncthreads();
  if (!INV) col_fft_fwd<R, NY, T>(tile, T, 1, tw, tid, kFftThreads);
  else col_fft_inv<R, NY, T>(tile, T, 1, tw, tid, kFftThreads);
  __syncthreads();
  for (int idx = tid; idx < NY * T; idx += kFftThreads) {
    const int l = idx % T, r = idx / T;
    if (z0 + l < Zc) base[(long long)r * Zc + l] = tile[idx];
  }
}

// Y pass, second generation: same tiles, but the first stage reads global memory and the last
// stage writes it (in place), so the tile is only the inter-stage exchange buffer.
template <typename R, int NY, int T, bool INV>
__global__ void __launch_bounds__(kFftThreads)
ypass2_kernel(typename Cx<R>::T* __restrict__ spec, int X, int Zc,
              const typename Cx<R>::T* __restrict__ tw_g) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);  // NY x T
  C* tw = tile + NY * T;
  const int tid = threadIdx.x;
  for (int j = tid; j < NY; j += kFftThreads) tw[j] = tw_g[j];
  const int z0 = blockIdx.x * T;
  GSide<C> gs{spec + (((long long)blockIdx.z * X + blockIdx.y) * NY) * Zc + z0, Zc,
              (Zc - z0 < T) ? (Zc - z0) : T};
  __syncthreads();
  if (!INV) ColFFT<R, NY, NY, 0, T>::template fwd_g<true, true>(tile, T, 1, tw, tid, kFftThreads, gs, gs);
  else ColFFT<R, NY, NY, 0, T>::template inv_g<true, true>(tile, T, 1, tw, tid, kFftThreads, gs, gs);
}

// X pass with the multiplier: tiles of T consecutive words of the (Y x Zc) plane (contiguous
// for fixed x), all NX rows. NCH = 1 (beta == 0: channels independent, grid.y = N*dim) or
// NCH = dim (beta != 0: the dim channels of a subject together, grid.y = N).
template <typename R, int NX, int T, int D, int NCH, bool INVERSE>
__global__ void __launch_bounds__(kFftThreads)
xpass_kernel(typename Cx<R>::T* __restrict__ spec, long long plane, int Zc,
             const typename Cx<R>::T* __restrict__ tw_g, const R* __restrict__ wl0,
             const R* __restrict__ sl0, const R* __restrict__ wl1, const R* __restrict__ sl1,
             const R* __restrict__ wl2, const R* __restrict__ sl2, double alpha, double beta,
             double gamma, R scale) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);  // NCH x NX x T
  C* tw = tile + NCH * NX * T;
  const int tid = threadIdx.x;
  for (int j = tid; j < NX; j += kFftThreads) tw[j] = tw_g[j];
  const long long q0 = (long long)blockIdx.x * T;
  C* base = spec + (long long)blockIdx.y * NCH * NX * plane + q0;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
    for (int idx = tid; idx < NX * T; idx += kFftThreads) {
      const int l = idx % T, r = idx / T;
      C v;
      v.x = v.y = R(0);
      if (q0 + l < plane) v = base[((long long)ch * NX + r) * plane + l];
      tile[ch * NX * T + idx] = v;
    }
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    col_fft_fwd<R, NX, T>(tile + ch * NX * T, T, 1, tw, tid, kFftThreads);
  }
  __syncthreads();
  for (int idx = tid; idx < NX * T; idx += kFftThreads) {
    const int l = idx % T, r = idx / T;
    const long long q = q0 + l;
    if (q >= plane) continue;
    R w[3], s[3];
    w[0] = wl0[r];
    s[0] = sl0[r];
    if constexpr (D == 2) {
      w[1] = wl1[q]; s[1] = sl1[q];
      w[2] = R(0); s[2] = R(0);
    } else {
      const int py = (int)(q / Zc), pz = (int)(q - (long long)py * Zc);
      w[1] = wl1[py]; s[1] = sl1[py];
      w[2] = wl2[pz]; s[2] = sl2[pz];
    }
    if constexpr (NCH == 1) {
      const R f = scalar_symbol<R, D, INVERSE>(w, alpha, gamma);
      C v = tile[idx];
      if (INVERSE) { v.x = ((v.x * f) * f) * scale; v.y = ((v.y * f) * f) * scale; }
      else { v.x = (f * v.x) * scale; v.y = (f * v.y) * scale; }
      tile[idx] = v;
    } else {
      Symbol<R, D> S = make_symbol<R, D, INVERSE>(w, s, alpha, beta, gamma);
      R re[D], im[D];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        C v = tile[c * NX * T + idx];
        re[c] = v.x;
        im[c] = v.y;
      }
      apply_symbol<R, D, INVERSE>(S, re);
      apply_symbol<R, D, INVERSE>(S, im);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        C v;
        v.x = re[c] * scale;
        v.y = im[c] * scale;
        tile[c * NX * T + idx] = v;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    col_fft_inv<R, NX, T>(tile + ch * NX * T, T, 1, tw, tid, kFftThreads);
  }
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
    for (int idx = tid; idx < NX * T; idx += kFftThreads) {
      const int l = idx % T, r = idx / T;
      if (q0 + l < plane) base[((long long)ch * NX + r) * plane + l] = tile[ch * NX * T + idx];
    }
}

// ------------------------------------------------------------------------------------------
// X pass, second generation. Same tiling as xpass_kernel, but
//  - the first forward stage loads its radix inputs straight from global memory and the last
//    inverse stage stores straight to global memory (no staging copy of the tile, two fewer
//    shared-memory round trips per element);
//  - a thread keeps one column l = tid % T for the whole kernel, so the (y,z) part of the symbol
//    (LUT lookups, the q -> (ry, rz) division) is computed once per thread, not once per element;
//  - beta == 0 path: reciprocal square root in fp32 round-to-nearest intrinsics (equal to the
//    reference's double division rounded to float except on double-rounding ties).
template <typename R>
__device__ __forceinline__ R oo_sqrt_fast(R x) {
  if constexpr (sizeof(R) == 4) {
    float s = ((double)x < 1e-8) ? 1e-4f : __fsqrt_rn((float)x);
    return (R)__frcp_rn(s);
  } else {
    return oo_sqrt<R>(x);
  }
}

// 1 / safe_sqrt(lambda^2) of the beta == 0 symbol without the square root: in binary round-to-nearest
// arithmetic sqrt(fl(x*x)) == |x| exactly (no over-/underflow: Lm <= 1e-8f is the safe_sqrt branch --
// for a float x, (double)x < 1e-8 <=> x <= 1e-8f -- and an infinite Lm keeps sqrt's infinity), so the
// result is bit-identical to oo_sqrt_fast(Lm) at a third of its instructions (no MUFU.RSQ + fix-up).
// The reciprocal itself: MUFU.RCP + one FMA Newton step, which IS the round-to-nearest reciprocal for every
// float in [2^-64, 2^64] (scripts/micro/rcp_rn.cu checks all 1.07e9 of them against __frcp_rn on the GPU:
// 0 mismatches) without __frcp_rn's range-check branch and slow-path call around every element; arguments
// outside that range (alpha or gamma beyond 1e19) take __frcp_rn.
__device__ __noinline__ float rcp_rn_slow(float s) { return __frcp_rn(s); }  // out of line: cold path
// beta == 0: lambda = fl(gamma + alpha * sw) with sw = sum of up to three 2(1 - cos) in [0, 12.000002]. True when every
// lambda of a launch is in [2e-4, 1e15] (lambda^2 > 1e-8 and finite, lambda < 2^64); false for NaN parameters.
// LGM_NO_SAFE_LAMBDA=1 keeps the guarded multiplier (kernel experiments, tests).
inline bool lambda_range_safe(double alpha, double gamma) {
  static const bool off = getenv("LGM_NO_SAFE_LAMBDA") != nullptr;
  return !off && alpha >= 0.0 && gamma >= 2e-4 && gamma + 13.0 * alpha <= 1e15;
}
// SAFE: the host has checked (lambda_range_safe) that every lambda of this launch lies in [2e-4, 1e15]: neither
// safe_sqrt guard can fire and the reciprocal's argument is inside the range where MUFU.RCP + one Newton step
// equals __frcp_rn, so the two selects, the range test and the out-of-line call go -- same bits, 16-18 % fewer
// instructions in the X passes (xpassq 3608 -> 2960 SASS instructions, X pass at 8 x 256^3 1.10 -> 1.00 ms).
template <typename R, bool SAFE = false>
__device__ __forceinline__ R oo_lambda_fast(R lambda, R Lm) {
  if constexpr (sizeof(R) == 4) {
    float s;
    if constexpr (SAFE) {
      s = fabsf(lambda);
    } else {
      s = (Lm <= 1e-8f) ? 1e-4f : (Lm == INFINITY ? Lm : fabsf(lambda));
      if (__builtin_expect(!(s <= 18446744073709551616.f), 0)) return (R)rcp_rn_slow(s);   // > 2^64, inf, NaN
    }
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    const float e = __fmaf_rn(-s, r, 1.f);
    return (R)__fmaf_rn(r, e, r);
  } else {
    return oo_sqrt<R>(Lm);
  }
}

}  // namespace lgm
#include "qslab.cuh"
namespace lgm {

template <typename R, int NX, int T, int D, int NCH, bool INVERSE, bool SAFE = false>
__global__ void __launch_bounds__(kFftThreads, (sizeof(R) == 4 && NCH == 1) ? (NX <= 128 ? LGM_XPASS_MINBLOCKS : 3) : 2)
xpass2_kernel(typename Cx<R>::T* __restrict__ spec, long long plane, int Zc,
              const typename Cx<R>::T* __restrict__ tw_g, const R* __restrict__ wl0,
              const R* __restrict__ sl0, const R* __restrict__ wl1, const R* __restrict__ sl1,
              const R* __restrict__ wl2, const R* __restrict__ sl2, double alpha, double beta,
              double gamma, R scale, int rev) {
  using C = typename Cx<R>::T;
  static_assert(kFftThreads % T == 0, "a thread must own one column");
  const unsigned bx = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const unsigned by = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);  // NCH x NX x T
  C* tw = tile + NCH * NX * T;
  R* lx = reinterpret_cast<R*>(tw + NX);     // wl0[NX] (+ sl0[NX] when NCH > 1)
  const int tid = threadIdx.x;
  for (int j = tid; j < NX; j += kFftThreads) {
    tw[j] = tw_g[j];
    lx[j] = wl0[j];
    if (NCH > 1) lx[NX + j] = sl0[j];
  }
  const long long q0 = (long long)bx * T;
  const int lvalid = (int)((plane - q0 < T) ? (plane - q0) : T);
  C* base = spec + (long long)by * NCH * NX * plane + q0;
  if (LGM_PF_XPASS > 0 && sizeof(R) == 4) {  // rows of the tile LGM_PF_XPASS blocks ahead in launch order
    const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + LGM_PF_XPASS;
    if (lin < (long long)gridDim.x * gridDim.y) {
      const unsigned pbx0 = (unsigned)(lin % gridDim.x), pby0 = (unsigned)(lin / gridDim.x);
      const unsigned pbx = rev ? gridDim.x - 1 - pbx0 : pbx0, pby = rev ? gridDim.y - 1 - pby0 : pby0;
      const long long pq0 = (long long)pbx * T;
#ifndef LGM_PF_XPASS_MODE
#define LGM_PF_XPASS_MODE 0  /* 0: one bulk prefetch per row; n > 0: n plain prefetch.global.L2 per row */
#endif
      if (plane - pq0 >= T) {
        if (LGM_PF_XPASS_MODE == 0) {
          for (int r = tid; r < NCH * NX; r += kFftThreads)
            l2_prefetch(spec + (long long)pby * NCH * NX * plane + (long long)r * plane + pq0, (unsigned)(T * sizeof(C)));
        } else {
          constexpr int PM = LGM_PF_XPASS_MODE > 0 ? LGM_PF_XPASS_MODE : 1;
          for (int pi = tid; pi < NCH * NX * PM; pi += kFftThreads)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(spec + (long long)pby * NCH * NX * plane + (long long)(pi / PM) * plane + pq0 + (pi % PM) * (T / PM)));
        }
      }
    }
  }
  // per-thread (y,z) part of the symbol
  const int l = tid % T;
  R wy = R(0), wz = R(0), sy = R(0), sz = R(0);
  if (l < lvalid) {
    const long long q = q0 + l;
    if constexpr (D == 2) {
      wy = wl1[q];
      if (NCH > 1) sy = sl1[q];
    } else {
      const int py = (int)(q / Zc), pz = (int)(q - (long long)py * Zc);
      wy = wl1[py];
      wz = wl2[pz];
      if (NCH > 1) { sy = sl1[py]; sz = sl2[pz]; }
    }
  }
  __syncthreads();
  constexpr int NS = (ilog2(NX) + 3) / 4;  // radix stages of the X transform
  if constexpr (NCH == 1 && NS >= 2) {
    // beta == 0: forward -> scalar multiplier -> inverse with the innermost stage pair and the
    // multiplier fused in registers (fft_mid_stage): 2 shared-memory round trips and 2 barriers for
    // a two-stage transform instead of 4 and 4.
    constexpr int RAD0 = 1 << stage_bits(ilog2(NX), 0);
    constexpr int RADL = 1 << stage_bits(ilog2(NX), NS - 1);
    fft_stage_edge<R, NX, RAD0, T, false>(base, plane, tile, T, 1, tw, lvalid, tid, kFftThreads);
    __syncthreads();
    if constexpr (NS > 2) {
      ColFFT<R, NX, NX / RAD0, 1, T>::fwd_nolast(tile, T, 1, tw, tid, kFftThreads);
      __syncthreads();
    }
    fft_mid_stage<R, NX, RADL, T>(tile, T, 1, tid, kFftThreads, [&](int r, int, C v) -> C {
      const R sw = (D == 2) ? (lx[r] + wy) : (lx[r] + wy + wz);
      const R lambda = (R)(gamma + alpha * (double)sw);
      const R Lm = lambda * lambda;
      if (INVERSE) {
        const R f = oo_lambda_fast<R, SAFE>(lambda, Lm);
        v.x = ((v.x * f) * f) * scale;
        v.y = ((v.y * f) * f) * scale;
      } else {
        v.x = (Lm * v.x) * scale;
        v.y = (Lm * v.y) * scale;
      }
      return v;
    });
    __syncthreads();
    if constexpr (NS > 2) {
      ColFFT<R, NX, NX / RAD0, 1, T>::inv_nolast(tile, T, 1, tw, tid, kFftThreads);
      __syncthreads();
    }
    fft_stage_edge<R, NX, RAD0, T, true>(base, plane, tile, T, 1, tw, lvalid, tid, kFftThreads);
  } else if constexpr (sizeof(R) == 4 && NCH > 1 && NS >= 2 && (1 << stage_bits(ilog2(NX), NS - 1)) <= LGM_XPASS_MULTI_RADMAX) {
    // beta != 0 with a small innermost radix: the same register fusion for the NCH coupled channels
    // (the matrix symbol is built once per frequency and applied to the real and imaginary parts)
    constexpr int RAD0 = 1 << stage_bits(ilog2(NX), 0);
    constexpr int RADL = 1 << stage_bits(ilog2(NX), NS - 1);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
      fft_stage_edge<R, NX, RAD0, T, false>(base + (long long)ch * NX * plane, plane, tile + ch * NX * T, T, 1, tw,
                                            lvalid, tid, kFftThreads);
    __syncthreads();
    if constexpr (NS > 2) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) ColFFT<R, NX, NX / RAD0, 1, T>::fwd_nolast(tile + ch * NX * T, T, 1, tw, tid, kFftThreads);
      __syncthreads();
    }
    fft_mid_stage_multi<R, NX, RADL, T, NCH>(tile, NX * T, T, 1, tid, kFftThreads, [&](int r, int, C (&v)[NCH]) {
      R w[3] = {lx[r], wy, wz};
      R sn[3] = {lx[NX + r], sy, sz};
      Symbol<R, D> S = make_symbol<R, D, INVERSE>(w, sn, alpha, beta, gamma);
      R re[D], im[D];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        re[c] = v[c].x;
        im[c] = v[c].y;
      }
      apply_symbol<R, D, INVERSE>(S, re);
      apply_symbol<R, D, INVERSE>(S, im);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        v[c].x = re[c] * scale;
        v[c].y = im[c] * scale;
      }
    });
    __syncthreads();
    if constexpr (NS > 2) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) ColFFT<R, NX, NX / RAD0, 1, T>::inv_nolast(tile + ch * NX * T, T, 1, tw, tid, kFftThreads);
      __syncthreads();
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
      fft_stage_edge<R, NX, RAD0, T, true>(base + (long long)ch * NX * plane, plane, tile + ch * NX * T, T, 1, tw,
                                           lvalid, tid, kFftThreads);
  } else {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
    col_fft_fwd_from_global<R, NX, T>(base + (long long)ch * NX * plane, plane, tile + ch * NX * T, T, 1, tw,
                                      lvalid, tid, kFftThreads);
  __syncthreads();
  for (int idx = tid; idx < NX * T; idx += kFftThreads) {  // idx % T == l for every iteration
    const int r = idx / T;
    R w[3] = {lx[r], wy, wz};
    if constexpr (NCH == 1) {
      R sw = (D == 2) ? (w[0] + w[1]) : (w[0] + w[1] + w[2]);
      const R lambda = (R)(gamma + alpha * (double)sw);
      const R Lm = lambda * lambda;
      C v = tile[idx];
      if (INVERSE) {
        const R f = oo_lambda_fast<R, SAFE>(lambda, Lm);
        v.x = ((v.x * f) * f) * scale;
        v.y = ((v.y * f) * f) * scale;
      } else {
        v.x = (Lm * v.x) * scale;
        v.y = (Lm * v.y) * scale;
      }
      tile[idx] = v;
    } else {
      R s[3] = {lx[NX + r], sy, sz};
      Symbol<R, D> S = make_symbol<R, D, INVERSE>(w, s, alpha, beta, gamma);
      R re[D], im[D];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        C v = tile[c * NX * T + idx];
        re[c] = v.x;
        im[c] = v.y;
      }
      apply_symbol<R, D, INVERSE>(S, re);
      apply_symbol<R, D, INVERSE>(S, im);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        C v;
        v.x = re[c] * scale;
        v.y = im[c] * scale;
        tile[c * NX * T + idx] = v;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
    col_fft_inv_to_global<R, NX, T>(base + (long long)ch * NX * plane, plane, tile + ch * NX * T, T, 1, tw,
                                    lvalid, tid, kFftThreads);
  }
}


// First-axis pass of the mixed-radix path: X forward -> multiplier -> X inverse in one kernel
// (mix_xmid, mixfft.cuh). NCH = 1 for beta == 0 (scalar symbol, channels independent), NCH = D
// otherwise (the reference's 3x3 Cholesky arithmetic on the channels of one frequency).
template <typename R, int D, int NCH, bool INVERSE>
__global__ void __launch_bounds__(kMixThreads)
mix_xpass_kernel(typename Cx<R>::T* __restrict__ spec, long long plane, int Zc, MixPlan pl, int Tsh,
                 const typename Cx<R>::T* __restrict__ tw, const R* __restrict__ wl0, const R* __restrict__ sl0,
                 const R* __restrict__ wl1, const R* __restrict__ sl1, const R* __restrict__ wl2,
                 const R* __restrict__ sl2, double alpha, double beta, double gamma) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char mixx_smem[];
  const int T = 1 << Tsh, NX = pl.n;
  C* bufs = reinterpret_cast<C*>(mixx_smem);                   // 2 x NCH x T x NX
  R* ly = reinterpret_cast<R*>(bufs + (size_t)2 * NCH * T * NX);  // per line: wy, wz, sy, sz
  const long long q0 = (long long)blockIdx.x * T;
  const int lvalid = (int)((plane - q0 < T) ? (plane - q0) : T);
  for (int l = threadIdx.x; l < T; l += kMixThreads) {
    R wy = R(0), wz = R(0), sy = R(0), sz = R(0);
    if (l < lvalid) {
      const long long q = q0 + l;
      if constexpr (D == 2) {
        wy = wl1[q];
        sy = sl1[q];
      } else {
        const int py = (int)(q / Zc), pz = (int)(q - (long long)py * Zc);
        wy = wl1[py];
        wz = wl2[pz];
        sy = sl1[py];
        sz = sl2[pz];
      }
    }
    ly[l] = wy;
    ly[T + l] = wz;
    ly[2 * T + l] = sy;
    ly[3 * T + l] = sz;
  }
  C* base = spec + (long long)blockIdx.y * NCH * NX * plane + q0;
  mix_xmid<R, NCH>(base, (long long)NX * plane, plane, lvalid, pl, Tsh, tw, bufs, [&](int r, int l, C (&v)[NCH]) {
    R w[3] = {wl0[r], ly[l], ly[T + l]};
    if constexpr (NCH == 1) {
      const R sw = (D == 2) ? (w[0] + w[1]) : (w[0] + w[1] + w[2]);
      const R lambda = (R)(gamma + alpha * (double)sw);
      const R Lm = lambda * lambda;
      if (INVERSE) {
        const R f = oo_lambda_fast<R>(lambda, Lm);
        v[0].x = (v[0].x * f) * f;
        v[0].y = (v[0].y * f) * f;
      } else {
        v[0].x = Lm * v[0].x;
        v[0].y = Lm * v[0].y;
      }
    } else {
      R sn[3] = {sl0[r], ly[2 * T + l], ly[3 * T + l]};
      Symbol<R, D> S = make_symbol<R, D, INVERSE>(w, sn, alpha, beta, gamma);
      R re[D], im[D];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        re[c] = v[c].x;
        im[c] = v[c].y;
      }
      apply_symbol<R, D, INVERSE>(S, re);
      apply_symbol<R, D, INVERSE>(S, im);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        v[c].x = re[c];
        v[c].y = im[c];
      }
    }
  });
}

// ------------------------------------------------------------------------------------------
// Direct-DFT path (any size): unitary transforms, out of place between two buffers.
// ------------------------------------------------------------------------------------------
// real lines (rows x n) -> (rows x nc) complex, scaled
template <typename R>
__global__ void dft_r2c_kernel(typename Cx<R>::T* __restrict__ out, const R* __restrict__ in,
                               long long rows, int n, int nc, const typename Cx<R>::T* __restrict__ tw,
                               R scale) {
  using C = typename Cx<R>::T;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= rows * nc) return;
  const long long row = id / nc;
  const int k = (int)(id - row * nc);
  R ar = 0, ai = 0;
  for (int j = 0; j < n; ++j) {
    C w = tw[(int)(((long long)j * k) % n)];
    R x = in[row * n + j];
    ar += x * w.x;
    ai += x * w.y;
  }
  C o;
  o.x = ar * scale;
  o.y = ai * scale;
  out[id] = o;
}
// complex lines along an axis of length n with element stride `st` (inner count = st):
// element index = (outer*n + j)*st + inner
template <typename R, bool INV>
__global__ void dft_c2c_kernel(typename Cx<R>::T* __restrict__ out,
                               const typename Cx<R>::T* __restrict__ in, long long total, int n,
                               long long st, const typename Cx<R>::T* __restrict__ tw, R scale) {
  using C = typename Cx<R>::T;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  const long long inner = id % st;
  const long long t = id / st;
  const int k = (int)(t % n);
  const long long outer = t / n;
  const C* line = in + outer * n * st + inner;
  R ar = 0, ai = 0;
  for (int j = 0; j < n; ++j) {
    C w = tw[(int)(((long long)j * k) % n)];
    if (INV) w.y = -w.y;
    C x = line[(long long)j * st];
    ar += x.x * w.x - x.y * w.y;
    ai += x.x * w.y + x.y * w.x;
  }
  C o;
  o.x = ar * scale;
  o.y = ai * scale;
  out[id] = o;
}
// (rows x nc) half spectrum -> real lines (rows x n); imaginary parts of DC/Nyquist ignored
template <typename R>
__global__ void dft_c2r_kernel(R* __restrict__ out, const typename Cx<R>::T* __restrict__ in,
                               long long rows, int n, int nc, const typename Cx<R>::T* __restrict__ tw,
                               R scale) {
  using C = typename Cx<R>::T;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= rows * n) return;
  const long long row = id / n;
  const int j = (int)(id - row * n);
  const C* line = in + row * nc;
  R acc = line[0].x;
  for (int k = 1; k < nc; ++k) {
    C w = tw[(int)(((long long)j * k) % n)];  // e^{-i th}; need Re(X e^{+i th})
    C x = line[k];
    R term = x.x * w.x + x.y * w.y;
    if (2 * k == n) acc += x.x * w.x;  // Nyquist: real, counted once
    else acc += R(2) * term;
  }
  out[id] = acc * scale;
}

// ------------------------------------------------------------------------------------------
// Host drivers
// ------------------------------------------------------------------------------------------
template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define LGM_CUDA_TRY(expr, what)                                                          \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) return set_error((int)e__, "%s: %s", what, cudaGetErrorString(e__)); \
  } while (0)

template <typename R>
struct FastLaunch {
  using C = typename Cx<R>::T;
  static constexpr int ZL = 32;                        // lines per CTA in the Z passes
  static constexpr int T = sizeof(R) == 4 ? 16 : 8;    // tile width of the Y / X passes

  template <int Z>
  static int zfwd(C* spec, const R* in, long long rows, const C* tw, cudaStream_t s) {
    constexpr int M = Z / 2;
    const size_t smem = sizeof(C) * ((size_t)(M + 1) * (ZL + 1) + Z + M);
    LGM_CUDA_TRY(set_smem(zfwd_kernel<R, Z, ZL>, smem), "zfwd smem");
    zfwd_kernel<R, Z, ZL><<<(unsigned)cdiv(rows, ZL), kFftThreads, smem, s>>>(spec, in, rows, tw);
    count_launch("zfwd", s);
    return LGM_OK;
  }
  template <int Z>
  static int zinv(R* out, const C* spec, long long rows, const C* tw, cudaStream_t s) {
    constexpr int M = Z / 2;
    const size_t smem = sizeof(C) * ((size_t)(M + 1) * (ZL + 1) + Z + M);
    LGM_CUDA_TRY(set_smem(zinv_kernel<R, Z, ZL>, smem), "zinv smem");
    zinv_kernel<R, Z, ZL><<<(unsigned)cdiv(rows, ZL), kFftThreads, smem, s>>>(out, spec, rows, tw);
    count_launch("zinv", s);
    return LGM_OK;
  }
  template <int NY, bool INV>
  static int ypass(C* spec, int NC, int X, int Zc, const C* tw, cudaStream_t s) {
    const size_t smem = sizeof(C) * ((size_t)NY * T + NY);
    LGM_CUDA_TRY(set_smem(ypass2_kernel<R, NY, T, INV>, smem), "ypass smem");
    dim3 grid((unsigned)cdiv(Zc, T), (unsigned)X, (unsigned)NC);
    ypass2_kernel<R, NY, T, INV><<<grid, kFftThreads, smem, s>>>(spec, X, Zc, tw);
    count_launch("ypass", s);
    return LGM_OK;
  }
  template <int NX, int D, int NCH, bool INVERSE, bool SAFE = false>
  static int xpass1(C* spec, long long N, long long plane, int Zc, const FluidPlan& p, double alpha,
                    double beta, double gamma, R scale, int rev, cudaStream_t s) {
    // tile width: 32 words (256 B runs) while the tile stays small, else the class default
    constexpr int TX = (sizeof(R) == 4 && NCH * NX <= LGM_XPASS_TX32_MAX) ? 32 : T;  // 32 at NX=256 measured slower (regs)
    const size_t smem = sizeof(C) * ((size_t)NCH * NX * TX + NX) + sizeof(R) * 2 * NX;
    LGM_CUDA_TRY(set_smem(xpass2_kernel<R, NX, TX, D, NCH, INVERSE, SAFE>, smem), "xpass smem");
    dim3 grid((unsigned)cdiv(plane, TX), (unsigned)(NCH == 1 ? N * D : N));
    xpass2_kernel<R, NX, TX, D, NCH, INVERSE, SAFE><<<grid, kFftThreads, smem, s>>>(
        spec, plane, Zc, (const C*)p.tw[0], (const R*)p.wl[0], (const R*)p.sl[0], (const R*)p.wl[1],
        (const R*)p.sl[1], (const R*)p.wl[2], (const R*)p.sl[2], alpha, beta, gamma, scale, rev);
    count_launch("xpass", s);
    return LGM_OK;
  }
  template <int NX, int D>
  static int xpass(C* spec, long long N, long long plane, int Zc, const FluidPlan& p, int inverse,
                   double alpha, double beta, double gamma, R scale, int rev, cudaStream_t s) {
    if (beta == 0.0) {
      if constexpr (sizeof(R) == 4) {
        if (inverse && lambda_range_safe(alpha, gamma))
          return xpass1<NX, D, 1, true, true>(spec, N, plane, Zc, p, alpha, beta, gamma, scale, rev, s);
      }
      return inverse ? xpass1<NX, D, 1, true>(spec, N, plane, Zc, p, alpha, beta, gamma, scale, rev, s)
                     : xpass1<NX, D, 1, false>(spec, N, plane, Zc, p, alpha, beta, gamma, scale, rev, s);
    }
    return inverse ? xpass1<NX, D, D, true>(spec, N, plane, Zc, p, alpha, beta, gamma, scale, rev, s)
                   : xpass1<NX, D, D, false>(spec, N, plane, Zc, p, alpha, beta, gamma, scale, rev, s);
  }
};

// size switches ------------------------------------------------------------------------------
#define LGM_SWITCH_POW2(n, MAXN, CALL)                     \
  switch (n) {                                             \
    case 8: { constexpr int NN = 8; CALL; } break;         \
    case 16: { constexpr int NN = 16; CALL; } break;       \
    case 32: { constexpr int NN = 32; CALL; } break;       \
    case 64: { constexpr int NN = 64; CALL; } break;       \
    case 128: { constexpr int NN = 128; CALL; } break;     \
    case 256: { constexpr int NN = 256; CALL; } break;     \
    case 512: if constexpr (MAXN >= 512) { constexpr int NN = 512; CALL; } break; \
    default: break;                                        \
  }
#define LGM_SWITCH_Z(n, MAXN, CALL)                        \
  switch (n) {                                             \
    case 16: { constexpr int NN = 16; CALL; } break;       \
    case 32: { constexpr int NN = 32; CALL; } break;       \
    case 64: { constexpr int NN = 64; CALL; } break;       \
    case 128: { constexpr int NN = 128; CALL; } break;     \
    case 256: { constexpr int NN = 256; CALL; } break;     \
    case 512: { constexpr int NN = 512; CALL; } break;     \
    case 1024: if constexpr (MAXN >= 512) { constexpr int NN = 1024; CALL; } break; \
    default: break;                                        \
  }

// Slab path launchers (3-D, Y == Z in {16,32,64,128}); LGM_EUNSUP when the shape has no slab kernel.
template <typename R, int YZ>
static int slab_launch(bool inv, void* real, typename Cx<R>::T* spec, long long slabs, const FluidPlan& p,
                       int rev, cudaStream_t s, const double* post = nullptr) {
  using C = typename Cx<R>::T;
  constexpr int M = YZ / 2;
  const size_t smem = sizeof(C) * ((size_t)(M + 1) * (YZ + 1) + YZ + M + YZ);
  if (!inv) {
    LGM_CUDA_TRY(set_smem(slab_fwd_kernel<R, YZ, YZ>, smem), "slab_fwd smem");
    slab_fwd_kernel<R, YZ, YZ><<<(unsigned)slabs, kFftThreads, smem, s>>>(spec, (const R*)real, (const C*)p.tw[2], (const C*)p.tw[1], rev);
    count_launch("slab_fwd", s);
  } else {
    if (post) {
      LGM_CUDA_TRY(set_smem(slab_inv_kernel<R, YZ, YZ, true>, smem), "slab_inv smem");
      slab_inv_kernel<R, YZ, YZ, true><<<(unsigned)slabs, kFftThreads, smem, s>>>((R*)real, spec, (const C*)p.tw[2], (const C*)p.tw[1], rev, (R)*post);
    } else {
      LGM_CUDA_TRY(set_smem(slab_inv_kernel<R, YZ, YZ, false>, smem), "slab_inv smem");
      slab_inv_kernel<R, YZ, YZ, false><<<(unsigned)slabs, kFftThreads, smem, s>>>((R*)real, spec, (const C*)p.tw[2], (const C*)p.tw[1], rev, R(1));
    }
    count_launch("slab_inv", s);
  }
  return LGM_OK;
}
// 256 x 256 slabs (fp32): four-CTA cluster kernels, see cslab_fwd_kernel.
static int cslab_launch(bool inv, void* real, float2* spec, long long slabs, const FluidPlan& p, int rev, cudaStream_t s) {
  // LGM_NO_CLUSTER_SLAB: kernel experiments. `broken` is set when the device refuses the cluster
  // launch (e.g. a partition without enough co-schedulable SMs): the unfused passes take over.
  static const bool off = getenv("LGM_NO_CLUSTER_SLAB") != nullptr;
  static bool broken = false;
  if (off || broken) return LGM_EUNSUP;
  const size_t smem = sizeof(float2) * ((size_t)kCsRows * kCsP + 256 + 128 + 256);
  if (!inv) {
    if (set_smem(cslab_fwd_kernel, smem) != cudaSuccess) { cudaGetLastError(); broken = true; return LGM_EUNSUP; }
    cslab_fwd_kernel<<<(unsigned)(kCsNC * slabs), kCsThreads, smem, s>>>(spec, (const float*)real, (const float2*)p.tw[2], (const float2*)p.tw[1], rev);
    if (cudaPeekAtLastError() != cudaSuccess) { cudaGetLastError(); broken = true; return LGM_EUNSUP; }
    count_launch("slab_fwd", s);
  } else {
    if (set_smem(cslab_inv_kernel, smem) != cudaSuccess) { cudaGetLastError(); broken = true; return LGM_EUNSUP; }
    cslab_inv_kernel<<<(unsigned)(kCsNC * slabs), kCsThreads, smem, s>>>((float*)real, spec, (const float2*)p.tw[2], (const float2*)p.tw[1], rev);
    if (cudaPeekAtLastError() != cudaSuccess) { cudaGetLastError(); broken = true; return LGM_EUNSUP; }
    count_launch("slab_inv", s);
  }
  return LGM_OK;
}

// Quarter-slab path (qslab.cuh): fp32, Y = Z = 256, beta == 0. Three launches, same names in the launch
// counters as the passes they replace. LGM_NO_QSLAB (kernel experiments): the cluster slab kernels.
template <int NX>
static int qslab_run(float* out, const float* in, float2* spec, long long NC, const FluidPlan& p, int inverse,
                     double alpha, double gamma, float scale, int rev0, int revx, cudaStream_t s, const double* post) {
  constexpr int Y = 256, Z = 256, YQ = Y / 4, M = Z / 2, ZC = M + 1;
  const size_t smem_s = sizeof(float2) * ((size_t)ZC * (YQ + 1) + Z + M + YQ);
  const size_t smem_x = sizeof(float2) * ((size_t)NX * kQxP + NX) + sizeof(float) * NX;
  const unsigned nslab = (unsigned)(4 * NC * NX);
  LGM_CUDA_TRY(set_smem(qslab_fwd_kernel<Y, Z>, smem_s), "qslab_fwd smem");
  qslab_fwd_kernel<Y, Z><<<nslab, kQsThreads, smem_s, s>>>(spec, in, (const float2*)p.tw[2], (const float2*)p.tw[1], rev0);
  count_launch("slab_fwd", s);
  dim3 grid((unsigned)(YQ * ZC / kQxT), (unsigned)NC);
  static_assert((YQ * ZC) % kQxT == 0, "quarter block must be a whole number of X-pass tiles");
  if (inverse && lambda_range_safe(alpha, gamma)) {
    LGM_CUDA_TRY(set_smem((xpassq_kernel<NX, YQ, true, true>), smem_x), "xpassq smem");
    xpassq_kernel<NX, YQ, true, true><<<grid, kQxThreads, smem_x, s>>>(spec, ZC, (const float2*)p.tw[0], (const float2*)p.tw[1],
        (const float*)p.q_lx, (const float*)p.q_wy, (const float*)p.wl[2], alpha, gamma, scale, revx);
  } else if (inverse) {
    LGM_CUDA_TRY(set_smem(xpassq_kernel<NX, YQ, true>, smem_x), "xpassq smem");
    xpassq_kernel<NX, YQ, true><<<grid, kQxThreads, smem_x, s>>>(spec, ZC, (const float2*)p.tw[0], (const float2*)p.tw[1],
        (const float*)p.q_lx, (const float*)p.q_wy, (const float*)p.wl[2], alpha, gamma, scale, revx);
  } else {
    LGM_CUDA_TRY(set_smem(xpassq_kernel<NX, YQ, false>, smem_x), "xpassq smem");
    xpassq_kernel<NX, YQ, false><<<grid, kQxThreads, smem_x, s>>>(spec, ZC, (const float2*)p.tw[0], (const float2*)p.tw[1],
        (const float*)p.q_lx, (const float*)p.q_wy, (const float*)p.wl[2], alpha, gamma, scale, revx);
  }
  count_launch("xpass", s);
  if (post) {
    LGM_CUDA_TRY(set_smem(qslab_inv_kernel<Y, Z, true>, smem_s), "qslab_inv smem");
    qslab_inv_kernel<Y, Z, true><<<nslab, kQsThreads, smem_s, s>>>(out, spec, (const float2*)p.tw[2], (const float2*)p.tw[1], rev0, (float)*post);
  } else {
    LGM_CUDA_TRY(set_smem(qslab_inv_kernel<Y, Z, false>, smem_s), "qslab_inv smem");
    qslab_inv_kernel<Y, Z, false><<<nslab, kQsThreads, smem_s, s>>>(out, spec, (const float2*)p.tw[2], (const float2*)p.tw[1], rev0, 1.f);
  }
  count_launch("slab_inv", s);
  return LGM_OK;
}

static bool qslab_enabled() {
  static const bool on = getenv("LGM_NO_QSLAB") == nullptr;
  return on;
}

template <typename R>
static int slab_pass(bool inv, int Y, int Z, void* real, typename Cx<R>::T* spec, long long slabs,
                     const FluidPlan& p, int rev, cudaStream_t s, const double* post = nullptr, bool* posted = nullptr) {
  // post / posted: scale the real output on its way out (single-CTA slab kernels only); *posted says
  // whether that happened
  if (Y != Z) return LGM_EUNSUP;
  if constexpr (sizeof(R) == 4) {
    if (Y == 256) return cslab_launch(inv, real, spec, slabs, p, rev, s);
  }
  if (posted && inv && post && (Y == 16 || Y == 32 || Y == 64 || Y == 128)) *posted = true;
  switch (Y) {
    case 16: return slab_launch<R, 16>(inv, real, spec, slabs, p, rev, s, post);
    case 32: return slab_launch<R, 32>(inv, real, spec, slabs, p, rev, s, post);
    case 64: return slab_launch<R, 64>(inv, real, spec, slabs, p, rev, s, post);
    case 128: return slab_launch<R, 128>(inv, real, spec, slabs, p, rev, s, post);
    default: return LGM_EUNSUP;
  }
}

// Subjects are pushed through all passes in chunks small enough for the chunk's spectrum to stay
// resident in the 126 MB L2, so that only the first read and the last write of a chunk go to HBM.
static bool alternate_passes() {
  static const bool on = getenv("LGM_NO_ALTERNATE") == nullptr;  // kernel experiments
  return on;
}

static long long chunk_budget_bytes() {
  static long long v = -1;
  if (v < 0) {
    const char* e = getenv("LGM_FLUID_CHUNK_MB");
    // default: no chunking. Measured on B200 (profiles/r1_notes.md): while the passes are bound by
    // SM work rather than HBM, smaller launches only add tail and launch-gap time.
    v = (e && atoll(e) > 0) ? atoll(e) << 20 : 1LL << 50;
  }
  return v;
}

template <typename R>
static int fluid_fast(void* out, const void* in, int64_t N, int dim, const int64_t* shape,
                      int inverse, double alpha, double beta, double gamma, void* ws,
                      const FluidPlan& p, int rev0, cudaStream_t s, const double* post, bool* posted) {
  // rev0: traversal direction of the two slab passes (0 = ascending block order); the X pass between
  // them runs the other way round so that each pass starts on what its predecessor wrote last (L2)
  using C = typename Cx<R>::T;
  using FL = FastLaunch<R>;
  constexpr int MAXN = sizeof(R) == 4 ? 512 : 256;
  const int X = (int)shape[0], Y = (int)shape[1];
  const int nlast = (int)shape[dim - 1];
  const int Zc = nlast / 2 + 1;
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  C* spec = (C*)ws;
  const R scale = (R)(1.0 / (double)V);
  const long long spec_bytes_per_subject = (long long)dim * (V / nlast) * Zc * sizeof(C);
  long long G = chunk_budget_bytes() / spec_bytes_per_subject;
  if (G < 1) G = 1;
  if (G > N) G = N;
  for (long long n0 = 0; n0 < N; n0 += G) {
    const long long g = (N - n0 < G) ? (N - n0) : G;
    const R* in_g = (const R*)in + n0 * dim * V;
    R* out_g = (R*)out + n0 * dim * V;
    const long long rows = g * dim * (V / nlast);
    int rc = LGM_EUNSUP;
    const int revx = alternate_passes() ? !rev0 : rev0;
    if constexpr (sizeof(R) == 4) {
      if (p.qslab && beta == 0.0 && qslab_enabled()) {
        switch (X) {
          case 64: rc = qslab_run<64>((float*)out_g, (const float*)in_g, (float2*)spec, g * dim, p, inverse, alpha, gamma, (float)scale, rev0, revx, s, post); break;
          case 128: rc = qslab_run<128>((float*)out_g, (const float*)in_g, (float2*)spec, g * dim, p, inverse, alpha, gamma, (float)scale, rev0, revx, s, post); break;
          case 256: rc = qslab_run<256>((float*)out_g, (const float*)in_g, (float2*)spec, g * dim, p, inverse, alpha, gamma, (float)scale, rev0, revx, s, post); break;
          default: break;
        }
        if (rc != LGM_EUNSUP) {
          if (rc) return rc;
          if (post) *posted = true;
          continue;
        }
      }
    }
    if (dim == 3) rc = slab_pass<R>(false, Y, nlast, (void*)in_g, spec, g * dim * X, p, rev0, s);
    const bool slab = (rc == LGM_OK);
    if (rc != LGM_OK && rc != LGM_EUNSUP) return rc;
    if (!slab) {
      rc = LGM_EUNSUP;
      LGM_SWITCH_Z(nlast, MAXN, rc = FL::template zfwd<NN>(spec, in_g, rows, (const C*)p.tw[dim - 1], s));
      if (rc) return rc == LGM_EUNSUP ? set_error(rc, "lgm_fluid_apply: unsupported size") : rc;
    }
    if (dim == 3) {
      if (!slab) {
        rc = LGM_EUNSUP;
        LGM_SWITCH_POW2(Y, MAXN, rc = (FL::template ypass<NN, false>(spec, (int)(g * dim), X, Zc, (const C*)p.tw[1], s)));
        if (rc) return rc;
      }
      rc = LGM_EUNSUP;
      LGM_SWITCH_POW2(X, MAXN, rc = (FL::template xpass<NN, 3>(spec, g, (long long)Y * Zc, Zc, p, inverse, alpha, beta, gamma, scale, revx, s)));
      if (rc) return rc;
      if (!slab) {
        rc = LGM_EUNSUP;
        LGM_SWITCH_POW2(Y, MAXN, rc = (FL::template ypass<NN, true>(spec, (int)(g * dim), X, Zc, (const C*)p.tw[1], s)));
        if (rc) return rc;
      }
    } else {
      rc = LGM_EUNSUP;
      LGM_SWITCH_POW2(X, MAXN, rc = (FL::template xpass<NN, 2>(spec, g, (long long)Zc, Zc, p, inverse, alpha, beta, gamma, scale, rev0, s)));
      if (rc) return rc;
    }
    rc = LGM_EUNSUP;
    if (slab) {
      rc = slab_pass<R>(true, Y, nlast, (void*)out_g, spec, g * dim * X, p, rev0, s, post, posted);
      if (rc == LGM_EUNSUP) {  // cluster launch refused after the forward slab ran: unfused inverse passes
        LGM_SWITCH_POW2(Y, MAXN, rc = (FL::template ypass<NN, true>(spec, (int)(g * dim), X, Zc, (const C*)p.tw[1], s)));
        if (rc) return rc;
        rc = LGM_EUNSUP;
      }
    }
    if (rc == LGM_EUNSUP)
      LGM_SWITCH_Z(nlast, MAXN, rc = FL::template zinv<NN>(out_g, spec, rows, (const C*)p.tw[dim - 1], s));
    if (rc) return rc;
  }
  return LGM_OK;
}

template <typename R, int D>
static void launch_operator(R* Fm, int inverse, const FluidPlan& p, double alpha, double beta,
                            double gamma, int N, const Geom<D>& g, cudaStream_t s) {
  const unsigned blocks = (unsigned)cdiv(g.V, 256);
  if (inverse)
    fluid_operator_kernel<R, D, true><<<blocks, 256, 0, s>>>(Fm, (const R*)p.wl[0], (const R*)p.sl[0], (const R*)p.wl[1], (const R*)p.sl[1], (const R*)p.wl[2], (const R*)p.sl[2], alpha, beta, gamma, N, g);
  else
    fluid_operator_kernel<R, D, false><<<blocks, 256, 0, s>>>(Fm, (const R*)p.wl[0], (const R*)p.sl[0], (const R*)p.wl[1], (const R*)p.sl[1], (const R*)p.wl[2], (const R*)p.sl[2], alpha, beta, gamma, N, g);
  count_launch("fluid_operator", s);
}

template <typename R>
static int fluid_naive(void* out, const void* in, int64_t N, int dim, const int64_t* shape,
                       int inverse, double alpha, double beta, double gamma, void* ws,
                       const FluidPlan& p, cudaStream_t s) {
  using C = typename Cx<R>::T;
  const int nlast = (int)shape[dim - 1], nc = nlast / 2 + 1;
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  const long long lines = N * dim * (V / nlast);   // real lines
  const long long S = lines * nc;                  // spectrum words
  C* A = (C*)ws;
  C* B = A + S;
  const R sc = (R)(1.0 / sqrt((double)V));
  const int th = 256;
  dft_r2c_kernel<R><<<(unsigned)cdiv(S, th), th, 0, s>>>(A, (const R*)in, lines, nlast, nc, (const C*)p.tw[dim - 1], sc);
  count_launch("dft_r2c", s);
  C* cur = A;
  C* oth = B;
  // remaining axes from the second-to-last to the first
  long long st = nc;
  for (int a = dim - 2; a >= 0; --a) {
    dft_c2c_kernel<R, false><<<(unsigned)cdiv(S, th), th, 0, s>>>(oth, cur, S, (int)shape[a], st, (const C*)p.tw[a], R(1));
    count_launch("dft_c2c", s);
    C* t = cur; cur = oth; oth = t;
    st *= shape[a];
  }
  int64_t sshape[3];
  for (int a = 0; a < dim; ++a) sshape[a] = shape[a];
  sshape[dim - 1] = nc;
  if (dim == 2) launch_operator<R, 2>((R*)cur, inverse, p, alpha, beta, gamma, (int)N, make_geom<2>(sshape), s);
  else launch_operator<R, 3>((R*)cur, inverse, p, alpha, beta, gamma, (int)N, make_geom<3>(sshape), s);
  st = nc;
  for (int a = dim - 2; a >= 0; --a) {
    dft_c2c_kernel<R, true><<<(unsigned)cdiv(S, th), th, 0, s>>>(oth, cur, S, (int)shape[a], st, (const C*)p.tw[a], R(1));
    count_launch("dft_c2c", s);
    C* t = cur; cur = oth; oth = t;
    st *= shape[a];
  }
  dft_c2r_kernel<R><<<(unsigned)cdiv(lines * nlast, th), th, 0, s>>>((R*)out, cur, lines, nlast, nc, (const C*)p.tw[dim - 1], sc);
  count_launch("dft_c2r", s);
  return LGM_OK;
}


// ------------------------------------------------------------------------------------------
// Mixed-radix path (mixfft.cuh): any size whose lines fit in shared memory. Same structure as
// the reference's rfft -> fluid_operator -> irfft (metric.py:11-19): one pass per axis each way, all
// in place on one natural-order half spectrum, O(n log n) and coalesced.
// ------------------------------------------------------------------------------------------
static int mix_lines_per_cta(int n, size_t csize, bool contiguous_lines) {
  const size_t pitch = contiguous_lines ? (size_t)(n | 1) : (size_t)n;
  for (int T = 32; T >= 1; T /= 2)
    if (2 * (size_t)T * pitch * csize <= 96 * 1024) return T;
  for (int T = 1; T >= 1; --T)
    if (2 * (size_t)T * pitch * csize <= 220 * 1024) return T;
  return 0;
}

template <typename R>
static bool mixed_ok(int dim, const int64_t* shape) {
  using C = typename Cx<R>::T;
  static const bool off = getenv("LGM_NO_MIXED_FFT") != nullptr;  // kernel experiments: direct DFT
  if (off) return false;
  for (int a = 0; a < dim; ++a) {
    if (shape[a] < 2 || shape[a] > (1 << 20)) return false;
    if (mix_lines_per_cta((int)shape[a], sizeof(C), a == dim - 1) < 1) return false;
  }
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  return V / shape[dim - 1] <= 65535LL * 32;  // grid.y of the axis passes
}

template <typename R>
static int fluid_mixed(void* out, const void* in, int64_t N, int dim, const int64_t* shape, int inverse,
                       double alpha, double beta, double gamma, void* ws, const FluidPlan& p, cudaStream_t s) {
  using C = typename Cx<R>::T;
  const int nlast = (int)shape[dim - 1], nc = nlast / 2 + 1;
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  const long long lines = N * dim * (V / nlast);
  C* spec = (C*)ws;
  const R sc = (R)(1.0 / sqrt((double)V));
  // last axis: even lengths as a half-length complex transform + split, odd ones full length
  const bool even = (nlast % 2 == 0) && nlast >= 4 && (((uintptr_t)in | (uintptr_t)out) % sizeof(C) == 0);
  const int nz = even ? nlast / 2 : nlast;
  const int Tz = mix_lines_per_cta(nz, sizeof(C), true);
  const size_t smem_z = 2 * (size_t)Tz * (nz | 1) * sizeof(C);
  const MixPlan plz = mix_factor(nz);
  if (even) {
    LGM_CUDA_TRY(set_smem(mix_r2c_even_kernel<R>, smem_z), "mix_r2c smem");
    mix_r2c_even_kernel<R><<<(unsigned)cdiv(lines, Tz), kMixThreads, smem_z, s>>>(spec, (const R*)in, lines, plz, ilog2(Tz), (const C*)p.tw[dim - 1], sc);
  } else {
    LGM_CUDA_TRY(set_smem(mix_r2c_kernel<R>, smem_z), "mix_r2c smem");
    mix_r2c_kernel<R><<<(unsigned)cdiv(lines, Tz), kMixThreads, smem_z, s>>>(spec, (const R*)in, lines, plz, ilog2(Tz), (const C*)p.tw[dim - 1], sc);
  }
  count_launch("mix_r2c", s);
  auto axis_pass = [&](int a, long long st, bool inv) -> int {
    const int n = (int)shape[a];
    const MixPlan pl = mix_factor(n);
    const int T = mix_lines_per_cta(n, sizeof(C), false);
    const size_t smem = 2 * (size_t)T * n * sizeof(C);
    const long long outer = lines * nc / ((long long)n * st);
    if (outer > 0x7fffffffLL) return set_error(LGM_EUNSUP, "lgm_fluid_apply: too many lines");
    dim3 grid((unsigned)cdiv(st, T), 1, 1);
    // outer count goes to grid.y in chunks of 65535
    for (long long o0 = 0; o0 < outer; o0 += 65535) {
      grid.y = (unsigned)((outer - o0 < 65535) ? (outer - o0) : 65535);
      C* base = spec + o0 * n * st;
      if (inv) {
        LGM_CUDA_TRY(set_smem(mix_c2c_kernel<R, true>, smem), "mix_c2c smem");
        mix_c2c_kernel<R, true><<<grid, kMixThreads, smem, s>>>(base, st, pl, ilog2(T), (const C*)p.tw[a]);
      } else {
        LGM_CUDA_TRY(set_smem(mix_c2c_kernel<R, false>, smem), "mix_c2c smem");
        mix_c2c_kernel<R, false><<<grid, kMixThreads, smem, s>>>(base, st, pl, ilog2(T), (const C*)p.tw[a]);
      }
      count_launch("mix_c2c", s);
    }
    return LGM_OK;
  };
  long long sts[3] = {0, 0, 0};
  {
    long long t = nc;
    for (int a = dim - 2; a >= 0; --a) {
      sts[a] = t;
      t *= shape[a];
    }
  }
  // middle axes forward (3-D: Y), then the first axis with the multiplier inside, then back
  for (int a = dim - 2; a >= 1; --a) {
    int rc = axis_pass(a, sts[a], false);
    if (rc) return rc;
  }
  {
    const int NX = (int)shape[0];
    const int nch = (beta == 0.0) ? 1 : dim;
    int T = 0;
    for (int t = 32; t >= 1; t /= 2)
      if (2 * (size_t)nch * t * NX * sizeof(C) + 4 * t * sizeof(R) <= (t >= 8 ? 100 : 220) * 1024) { T = t; break; }
    const long long plane = sts[0];
    if (T >= 1 && N * dim <= 65535) {
      const MixPlan pl = mix_factor(NX);
      int Tsh = 0;
      while ((1 << Tsh) < T) ++Tsh;
      const size_t smem = 2 * (size_t)nch * T * NX * sizeof(C) + 4 * T * sizeof(R);
      dim3 grid((unsigned)cdiv(plane, T), (unsigned)(nch == 1 ? N * dim : N));
#define LGM_MIXX(D_, NCH_, INV_)                                                                                   \
  do {                                                                                                             \
    LGM_CUDA_TRY(set_smem(mix_xpass_kernel<R, D_, NCH_, INV_>, smem), "mix_xpass smem");                           \
    mix_xpass_kernel<R, D_, NCH_, INV_><<<grid, kMixThreads, smem, s>>>(                                           \
        spec, plane, nc, pl, Tsh, (const C*)p.tw[0], (const R*)p.wl[0], (const R*)p.sl[0], (const R*)p.wl[1],      \
        (const R*)p.sl[1], (const R*)p.wl[2], (const R*)p.sl[2], alpha, beta, gamma);                              \
  } while (0)
      if (dim == 2) {
        if (nch == 1) { if (inverse) LGM_MIXX(2, 1, true); else LGM_MIXX(2, 1, false); }
        else { if (inverse) LGM_MIXX(2, 2, true); else LGM_MIXX(2, 2, false); }
      } else {
        if (nch == 1) { if (inverse) LGM_MIXX(3, 1, true); else LGM_MIXX(3, 1, false); }
        else { if (inverse) LGM_MIXX(3, 3, true); else LGM_MIXX(3, 3, false); }
      }
#undef LGM_MIXX
      count_launch("mix_xpass", s);
    } else {  // a first axis too long for the fused tile: three separate passes
      int rc = axis_pass(0, sts[0], false);
      if (rc) return rc;
      int64_t sshape[3];
      for (int a = 0; a < dim; ++a) sshape[a] = shape[a];
      sshape[dim - 1] = nc;
      if (dim == 2) launch_operator<R, 2>((R*)spec, inverse, p, alpha, beta, gamma, (int)N, make_geom<2>(sshape), s);
      else launch_operator<R, 3>((R*)spec, inverse, p, alpha, beta, gamma, (int)N, make_geom<3>(sshape), s);
      rc = axis_pass(0, sts[0], true);
      if (rc) return rc;
    }
  }
  for (int a = 1; a <= dim - 2; ++a) {
    int rc = axis_pass(a, sts[a], true);
    if (rc) return rc;
  }
  if (even) {
    LGM_CUDA_TRY(set_smem(mix_c2r_even_kernel<R>, smem_z), "mix_c2r smem");
    mix_c2r_even_kernel<R><<<(unsigned)cdiv(lines, Tz), kMixThreads, smem_z, s>>>((R*)out, spec, lines, plz, ilog2(Tz), (const C*)p.tw[dim - 1], sc);
  } else {
    LGM_CUDA_TRY(set_smem(mix_c2r_kernel<R>, smem_z), "mix_c2r smem");
    mix_c2r_kernel<R><<<(unsigned)cdiv(lines, Tz), kMixThreads, smem_z, s>>>((R*)out, spec, lines, plz, ilog2(Tz), (const C*)p.tw[dim - 1], sc);
  }
  count_launch("mix_c2r", s);
  return LGM_OK;
}

template <typename R>
static int64_t fluid_ws_bytes(int64_t N, int dim, const int64_t* shape) {
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  const long long nlast = shape[dim - 1];
  if (V == 0 || N == 0) return 0;
  const long long S = N * dim * (V / nlast) * (nlast / 2 + 1);
  const int bufs = fast_ok<R>(dim, shape) ? 1 : 2;
  return (int64_t)(S * 2 * sizeof(R) * bufs);
}

template <typename R>
__global__ void __launch_bounds__(256) post_scale_kernel(R* __restrict__ x, long long total, R post) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < total) x[i] = post_scale(post, x[i]);
}

// post != nullptr: the result is additionally scaled, out <- fl(fl(*post * out) + 0) -- inside the last
// kernel where the path has a hook for it (slab kernels), by one more elementwise launch elsewhere.
template <typename R>
int fluid_apply_t(void* out, const void* in, int64_t N, int dim, const int64_t* shape, int inverse,
                  double alpha, double beta, double gamma, void* ws, int64_t ws_bytes, int rev,
                  cudaStream_t s, const double* post = nullptr) {
  if (N == 0) return LGM_OK;
  for (int a = 0; a < dim; ++a)
    if (shape[a] <= 0) return LGM_OK;
  if (ws_bytes < fluid_ws_bytes<R>(N, dim, shape))
    return set_error(LGM_ENOSPC, "lgm_fluid_apply: workspace too small (%lld < %lld bytes)",
                     (long long)ws_bytes, (long long)fluid_ws_bytes<R>(N, dim, shape));
  const FluidPlan* p = nullptr;
  int rc = get_plan<R>(dim, shape, s, &p);
  if (rc) return rc;
  bool posted = false;
  if (p->fast) {
    if (((uintptr_t)in | (uintptr_t)out | (uintptr_t)ws) & 15)
      return set_error(LGM_EINVAL, "lgm_fluid_apply: pointers must be 16-byte aligned");
    rc = fluid_fast<R>(out, in, N, dim, shape, inverse, alpha, beta, gamma, ws, *p, rev, s, post, &posted);
  } else if (mixed_ok<R>(dim, shape)) {
    rc = fluid_mixed<R>(out, in, N, dim, shape, inverse, alpha, beta, gamma, ws, *p, s);
  } else {
    rc = fluid_naive<R>(out, in, N, dim, shape, inverse, alpha, beta, gamma, ws, *p, s);
  }
  if (rc) return rc;
  if (post && !posted) {
    long long total = N * dim;
    for (int a = 0; a < dim; ++a) total *= shape[a];
    post_scale_kernel<R><<<(unsigned)cdiv(total, 256), 256, 0, s>>>((R*)out, total, (R)*post);
    count_launch("post_scale", s);
  }
  return finish(s, "lgm_fluid_apply");
}

// lgm_fluid_apply with an explicit traversal direction (used by the EPDiff step / shoot drivers)
int fluid_apply_post(int dtype, void* out, const void* in, int64_t N, int dim, const int64_t* shape, int inverse,
                     double alpha, double beta, double gamma, void* ws, int64_t ws_bytes, int rev, const double* post,
                     cudaStream_t s) {
  if (dim != 2 && dim != 3) return set_error(LGM_EINVAL, "Only two- and three-dimensional fluid metric is supported");
  if (N < 0 || N > 21845) return set_error(LGM_EINVAL, "lgm_fluid_apply: batch size out of range");
  if (!(dim == 2 ? geom_fits<2>(shape) : geom_fits<3>(shape))) return set_error(LGM_EINVAL, "lgm_fluid_apply: volume too large");
  if (dtype == LGM_F32)
    return fluid_apply_t<float>(out, in, N, dim, shape, inverse, alpha, beta, gamma, ws, ws_bytes, rev, s, post);
  if (dtype == LGM_F64)
    return fluid_apply_t<double>(out, in, N, dim, shape, inverse, alpha, beta, gamma, ws, ws_bytes, rev, s, post);
  return set_error(LGM_EINVAL, "lgm_fluid_apply: unsupported dtype %d", dtype);
}

int fluid_apply_dir(int dtype, void* out, const void* in, int64_t N, int dim, const int64_t* shape, int inverse,
                    double alpha, double beta, double gamma, void* ws, int64_t ws_bytes, int rev, cudaStream_t s) {
  return fluid_apply_post(dtype, out, in, N, dim, shape, inverse, alpha, beta, gamma, ws, ws_bytes, rev, nullptr, s);
}

int64_t fluid_workspace_bytes(int dtype, int64_t N, int dim, const int64_t* shape) {
  return dtype == LGM_F32 ? fluid_ws_bytes<float>(N, dim, shape) : fluid_ws_bytes<double>(N, dim, shape);
}

}  // namespace lgm

using namespace lgm;

extern "C" int64_t lgm_fluid_workspace_bytes(int dtype, int64_t N, int dim, const int64_t* shape) {
  if ((dim != 2 && dim != 3) || (dtype != LGM_F32 && dtype != LGM_F64)) return -1;
  return fluid_workspace_bytes(dtype, N, dim, shape);
}

extern "C" int lgm_fluid_apply(int dtype, void* out, const void* in, int64_t N, int dim,
                               const int64_t* shape, int inverse, double alpha, double beta,
                               double gamma, void* workspace, int64_t workspace_bytes, void* stream) {
  return fluid_apply_dir(dtype, out, in, N, dim, shape, inverse, alpha, beta, gamma, workspace, workspace_bytes, 0,
                         (cudaStream_t)stream);
}

template <typename R>
static int fluid_operator_t(void* Fm, int inverse, const void* const* cl, const void* const* sl,
                            double alpha, double beta, double gamma, int64_t N, int dim,
                            const int64_t* spec_shape, cudaStream_t s) {
  FluidPlan p;
  for (int a = 0; a < dim; ++a) {
    p.wl[a] = const_cast<void*>(cl[a]);
    p.sl[a] = const_cast<void*>(sl[a]);
  }
  if (dim == 2) {
    Geom<2> g = make_geom<2>(spec_shape);
    if (g.V == 0 || N == 0) return LGM_OK;
    launch_operator<R, 2>((R*)Fm, inverse, p, alpha, beta, gamma, (int)N, g, s);
  } else {
    Geom<3> g = make_geom<3>(spec_shape);
    if (g.V == 0 || N == 0) return LGM_OK;
    launch_operator<R, 3>((R*)Fm, inverse, p, alpha, beta, gamma, (int)N, g, s);
  }
  return finish(s, "lgm_fluid_operator");
}

extern "C" int lgm_fluid_operator(int dtype, void* Fm, int inverse, const void* const* cos_luts,
                                  const void* const* sin_luts, double alpha, double beta,
                                  double gamma, int64_t N, int dim, const int64_t* spec_shape,
                                  void* stream) {
  LGM_REQUIRE(dim == 2 || dim == 3, "Only two- and three-dimensional fluid metric is supported");
  if (dtype == LGM_F32)
    return fluid_operator_t<float>(Fm, inverse, cos_luts, sin_luts, alpha, beta, gamma, N, dim, spec_shape, (cudaStream_t)stream);
  if (dtype == LGM_F64)
    return fluid_operator_t<double>(Fm, inverse, cos_luts, sin_luts, alpha, beta, gamma, N, dim, spec_shape, (cudaStream_t)stream);
  return set_error(LGM_EINVAL, "lgm_fluid_operator: unsupported dtype %d", dtype);
}
