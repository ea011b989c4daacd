// mixfft.cuh -- mixed-radix line FFTs for the FluidMetric passes of sizes that are NOT powers of two
// (the reference takes any size through cuFFT: lagomorph/metric.py:17-19; medical volumes are
// 160 x 192 x 160, 182 x 218 x 182, 192^3, ...).
//
// One CTA transforms T neighbouring lines of n points in shared memory with the Stockham autosort
// scheme: n = r_1 * r_2 * ... with radices 4, 2, 3, 5, 7 done as register butterflies and any other
// prime factor p as a generic stage (one thread per output, p terms each: O(n p), e.g. 218 = 2 * 109).
// Every stage reads buffer A and writes buffer B in natural order (no digit reversal), so the
// spectrum comes out in natural order and the Fourier multiplier reads its LUTs directly.
//   element (point q, line l) lives at buf[q*rs + l*ls]; lanes run over l (fastest) then over the
//   butterflies, so a warp touches runs of consecutive words (rs = T, ls = 1) or, for lines that are
//   contiguous in memory (the last axis), a tile with an odd pitch (rs = 1, ls = n | 1).
#pragma once
#include "fft.cuh"

namespace lgm {

struct MixPlan {
  int n = 0;
  int nst = 0;
  int rad[24];
  unsigned ns_magic[24];  // ceil(2^32 / Ns) of each stage, Ns = product of the radices before it
  unsigned m_magic[24];   // ceil(2^32 / (n / rad)) (generic prime stages)
};

// q = j / d for j * d < 2^32 with magic = ceil(2^32 / d); d == 1 is handled by the callers
__device__ __forceinline__ unsigned fast_div(unsigned j, unsigned magic) { return __umulhi(j, magic); }

inline MixPlan mix_factor(int n) {
  MixPlan p;
  p.n = n;
  int m = n;
  const int small[5] = {4, 2, 3, 5, 7};
  for (int i = 0; i < 5; ++i)
    while (m % small[i] == 0 && p.nst < 24) {
      p.rad[p.nst++] = small[i];
      m /= small[i];
    }
  for (int f = 11; (long long)f * f <= m; f += 2)
    while (m % f == 0 && p.nst < 24) {
      p.rad[p.nst++] = f;
      m /= f;
    }
  if (m > 1 && p.nst < 24) p.rad[p.nst++] = m;
  long long Ns = 1;
  for (int i = 0; i < p.nst; ++i) {
    p.ns_magic[i] = (unsigned)((0x100000000ULL + Ns - 1) / Ns);  // Ns == 1: wraps to 0, unused
    const long long mm = n / p.rad[i];
    p.m_magic[i] = (unsigned)((0x100000000ULL + mm - 1) / mm);
    Ns *= p.rad[i];
  }
  return p;
}

// ---- r-point DFTs in registers, natural order in and out; INV conjugates the roots -------------
template <bool INV, typename C>
__device__ __forceinline__ C mul_mi(C b) {  // forward: -i*b, inverse: +i*b
  C r;
  if (!INV) { r.x = b.y; r.y = -b.x; } else { r.x = -b.y; r.y = b.x; }
  return r;
}

template <int RAD> struct Roots;
template <> struct Roots<3> {
  static __device__ __forceinline__ double c(int k) { return k == 0 ? 1.0 : -0.5; }
  static __device__ __forceinline__ double s(int k) { return k == 0 ? 0.0 : (k == 1 ? 0.8660254037844386 : -0.8660254037844386); }
};
template <> struct Roots<5> {
  static __device__ __forceinline__ double c(int k) {
    return k == 0 ? 1.0 : ((k == 1 || k == 4) ? 0.30901699437494745 : -0.8090169943749475);
  }
  static __device__ __forceinline__ double s(int k) {
    return k == 0 ? 0.0 : (k == 1 ? 0.9510565162951535 : (k == 2 ? 0.5877852522924731 : (k == 3 ? -0.5877852522924731 : -0.9510565162951535)));
  }
};
template <> struct Roots<7> {
  static __device__ __forceinline__ double c(int k) {
    return k == 0 ? 1.0 : ((k == 1 || k == 6) ? 0.6234898018587336 : ((k == 2 || k == 5) ? -0.2225209339563144 : -0.9009688679024191));
  }
  static __device__ __forceinline__ double s(int k) {
    const double t[7] = {0.0, 0.7818314824680298, 0.9749279121818236, 0.4338837391175581, -0.4338837391175581,
                         -0.9749279121818236, -0.7818314824680298};
    return t[k];
  }
};

template <int RAD, bool INV, typename C>
__device__ __forceinline__ void small_dft(C (&v)[RAD]) {
  using R = decltype(v[0].x);
  if constexpr (RAD == 2) {
    const C a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if constexpr (RAD == 4) {
    const C a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    const C c = cadd(v[1], v[3]), d = mul_mi<INV>(csub(v[1], v[3]));
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
  } else {  // odd prime: V_q = v0 + sum_n [ (v_n + v_{r-n}) cos(2 pi n q / r) -+ i (v_n - v_{r-n}) sin(2 pi n q / r) ]
    constexpr int H = (RAD - 1) / 2;
    C sm[H], df[H];
#pragma unroll
    for (int n = 1; n <= H; ++n) {
      sm[n - 1] = cadd(v[n], v[RAD - n]);
      df[n - 1] = csub(v[n], v[RAD - n]);
    }
    C out[RAD];
    out[0] = v[0];
#pragma unroll
    for (int n = 0; n < H; ++n) out[0] = cadd(out[0], sm[n]);
#pragma unroll
    for (int q = 1; q <= H; ++q) {
      C A = v[0], B;
      B.x = B.y = R(0);
#pragma unroll
      for (int n = 1; n <= H; ++n) {
        const R cc = (R)Roots<RAD>::c((n * q) % RAD), ss = (R)Roots<RAD>::s((n * q) % RAD);
        A.x += sm[n - 1].x * cc;
        A.y += sm[n - 1].y * cc;
        B.x += df[n - 1].x * ss;
        B.y += df[n - 1].y * ss;
      }
      const C iB = mul_mi<INV>(B);
      out[q] = cadd(A, iB);
      out[RAD - q] = csub(A, iB);
    }
#pragma unroll
    for (int q = 0; q < RAD; ++q) v[q] = out[q];
  }
}

// one Stockham stage of radix RAD over T lines: a -> b. Ns = product of the radices already done.
// T is a power of two (Tsh = log2 T); magic = ceil(2^32 / Ns).
template <typename R, int RAD, bool INV>
__device__ __forceinline__ void mix_stage(const typename Cx<R>::T* a, typename Cx<R>::T* b, int n, int Ns,
                                          unsigned magic, int rs, int ls, int Tsh,
                                          const typename Cx<R>::T* __restrict__ tw, int tws, int tid, int nth) {
  using C = typename Cx<R>::T;
  const int m = n / RAD;
  const int tstep = n / (Ns * RAD) * tws;  // tw has n * tws entries (tws = 2: table of the real length)
  const int items = m << Tsh;
  for (int it = tid; it < items; it += nth) {
    const int l = it & ((1 << Tsh) - 1), j = it >> Tsh;
    const int jhi = (Ns > 1) ? (int)fast_div((unsigned)j, magic) : j;
    const int k = j - jhi * Ns;
    const C* pa = a + j * rs + l * ls;
    C v[RAD];
#pragma unroll
    for (int nn = 0; nn < RAD; ++nn) {
      C x = pa[nn * m * rs];
      if (nn > 0 && Ns > 1) {
        const C w = tw[nn * k * tstep];  // e^{-2 pi i nn k / (Ns RAD)}
        x = INV ? cmulc(x, w) : cmul(x, w);
      }
      v[nn] = x;
    }
    small_dft<RAD, INV>(v);
    C* pb = b + (jhi * Ns * RAD + k) * rs + l * ls;
#pragma unroll
    for (int q = 0; q < RAD; ++q) pb[q * Ns * rs] = v[q];
  }
}

// generic prime radix p: one thread per output point, p terms each
template <typename R, bool INV>
__device__ __forceinline__ void mix_stage_prime(const typename Cx<R>::T* a, typename Cx<R>::T* b, int n, int p, int Ns,
                                                unsigned magic, unsigned m_magic, int rs, int ls, int Tsh,
                                                const typename Cx<R>::T* __restrict__ tw, int tws, int tid, int nth) {
  using C = typename Cx<R>::T;
  const int m = n / p;
  const int tstep = n / (Ns * p), rstep = n / p;
  const int items = n << Tsh;
  for (int it = tid; it < items; it += nth) {
    const int l = it & ((1 << Tsh) - 1), o = it >> Tsh;  // o = output point index within the line: (j, q)
    const int q = (m > 1) ? (int)fast_div((unsigned)o, m_magic) : o;
    const int j = o - q * m;
    const int jhi = (Ns > 1) ? (int)fast_div((unsigned)j, magic) : j;
    const int k = j - jhi * Ns;
    const C* pa = a + j * rs + l * ls;
    C acc = pa[0];
    int ti = 0, ri = 0;  // (nn k tstep) and (nn q mod p) rstep, both < n
    for (int nn = 1; nn < p; ++nn) {
      ti += k * tstep;
      ri += q * rstep;
      if (ri >= n) ri -= n;
      int idx = ti + ri;
      if (idx >= n) idx -= n;
      const C w = tw[idx * tws];
      const C x = pa[nn * m * rs];
      const C t = INV ? cmulc(x, w) : cmul(x, w);
      acc.x += t.x;
      acc.y += t.y;
    }
    b[(jhi * Ns * p + k + q * Ns) * rs + l * ls] = acc;
  }
}

// full transform of T lines; returns the buffer that holds the result
template <typename R, bool INV>
__device__ __forceinline__ typename Cx<R>::T* mix_fft(typename Cx<R>::T* a, typename Cx<R>::T* b, const MixPlan& pl,
                                                      int rs, int ls, int Tsh,
                                                      const typename Cx<R>::T* __restrict__ tw, int tid, int nth,
                                                      int tws = 1) {
  int Ns = 1;
  const int n = pl.n;
  for (int s = 0; s < pl.nst; ++s) {
    const int r = pl.rad[s];
    const unsigned mg = pl.ns_magic[s];
    switch (r) {
      case 2: mix_stage<R, 2, INV>(a, b, n, Ns, mg, rs, ls, Tsh, tw, tws, tid, nth); break;
      case 3: mix_stage<R, 3, INV>(a, b, n, Ns, mg, rs, ls, Tsh, tw, tws, tid, nth); break;
      case 4: mix_stage<R, 4, INV>(a, b, n, Ns, mg, rs, ls, Tsh, tw, tws, tid, nth); break;
      case 5: mix_stage<R, 5, INV>(a, b, n, Ns, mg, rs, ls, Tsh, tw, tws, tid, nth); break;
      case 7: mix_stage<R, 7, INV>(a, b, n, Ns, mg, rs, ls, Tsh, tw, tws, tid, nth); break;
      default: mix_stage_prime<R, INV>(a, b, n, r, Ns, mg, pl.m_magic[s], rs, ls, Tsh, tw, tws, tid, nth); break;
    }
    __syncthreads();
    typename Cx<R>::T* t = a;
    a = b;
    b = t;
    Ns *= r;
  }
  return a;
}

constexpr int kMixThreads = 256;

// complex lines along an axis of length n whose points are `st` words apart (st = number of
// neighbouring lines): element (outer, q, inner) at (outer*n + q)*st + inner. In place.
template <typename R, bool INV>
__global__ void __launch_bounds__(kMixThreads)
mix_c2c_kernel(typename Cx<R>::T* __restrict__ data, long long st, MixPlan pl, int Tsh,
               const typename Cx<R>::T* __restrict__ tw) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char mix_smem[];
  const int n = pl.n, T = 1 << Tsh;
  C* a = reinterpret_cast<C*>(mix_smem);
  C* b = a + (size_t)T * n;
  const long long inner0 = (long long)blockIdx.x * T;
  const int lvalid = (int)((st - inner0 < T) ? (st - inner0) : T);
  C* g = data + (long long)blockIdx.y * n * st + inner0;
  const int tid = threadIdx.x;
  for (int i = tid; i < T * n; i += kMixThreads) {
    const int l = i & (T - 1), q = i >> Tsh;
    C v;
    v.x = v.y = R(0);
    if (l < lvalid) v = g[(long long)q * st + l];
    a[i] = v;
  }
  __syncthreads();
  C* res = mix_fft<R, INV>(a, b, pl, T, 1, Tsh, tw, tid, kMixThreads);
  for (int i = tid; i < T * n; i += kMixThreads) {
    const int l = i & (T - 1), q = i >> Tsh;
    if (l < lvalid) g[(long long)q * st + l] = res[i];
  }
}

// real lines (rows x n) -> half spectra (rows x nc), scaled
template <typename R>
__global__ void __launch_bounds__(kMixThreads)
mix_r2c_kernel(typename Cx<R>::T* __restrict__ out, const R* __restrict__ in, long long rows, MixPlan pl, int Tsh,
               const typename Cx<R>::T* __restrict__ tw, R scale) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char mix_smem[];
  const int n = pl.n, nc = n / 2 + 1, pitch = n | 1, T = 1 << Tsh;
  C* a = reinterpret_cast<C*>(mix_smem);
  C* b = a + (size_t)T * pitch;
  const long long row0 = (long long)blockIdx.x * T;
  const int lvalid = (int)((rows - row0 < T) ? (rows - row0) : T);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;  // a warp owns whole lines in the fill / drain loops
  for (int l = warp; l < T; l += kMixThreads / 32)
    for (int q = lane; q < n; q += 32) {
      C v;
      v.x = (l < lvalid) ? in[(row0 + l) * n + q] : R(0);
      v.y = R(0);
      a[l * pitch + q] = v;
    }
  __syncthreads();
  C* res = mix_fft<R, false>(a, b, pl, 1, pitch, Tsh, tw, tid, kMixThreads);
  for (int l = warp; l < lvalid; l += kMixThreads / 32)
    for (int q = lane; q < nc; q += 32) {
      C v = res[l * pitch + q];
      v.x *= scale;
      v.y *= scale;
      out[(row0 + l) * nc + q] = v;
    }
}

// half spectra (rows x nc) -> real lines (rows x n), scaled: Hermitian extension + complex inverse
template <typename R>
__global__ void __launch_bounds__(kMixThreads)
mix_c2r_kernel(R* __restrict__ out, const typename Cx<R>::T* __restrict__ in, long long rows, MixPlan pl, int Tsh,
               const typename Cx<R>::T* __restrict__ tw, R scale) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char mix_smem[];
  const int n = pl.n, nc = n / 2 + 1, pitch = n | 1, T = 1 << Tsh;
  C* a = reinterpret_cast<C*>(mix_smem);
  C* b = a + (size_t)T * pitch;
  const long long row0 = (long long)blockIdx.x * T;
  const int lvalid = (int)((rows - row0 < T) ? (rows - row0) : T);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  for (int l = warp; l < T; l += kMixThreads / 32)
    for (int q = lane; q < nc; q += 32) {
      C v;
      v.x = v.y = R(0);
      if (l < lvalid) v = in[(row0 + l) * nc + q];
      a[l * pitch + q] = v;
      if (q > 0 && q < n - q) {  // X[n-q] = conj(X[q])
        v.y = -v.y;
        a[l * pitch + (n - q)] = v;
      }
    }
  __syncthreads();
  C* res = mix_fft<R, true>(a, b, pl, 1, pitch, Tsh, tw, tid, kMixThreads);
  for (int l = warp; l < lvalid; l += kMixThreads / 32)
    for (int q = lane; q < n; q += 32) out[(row0 + l) * n + q] = res[l * pitch + q].x * scale;
}

// Even last-axis length n = 2M: the real line is transformed as M complex points z[j] = x[2j] + i x[2j+1]
// (half the shared memory and half the butterflies of the full-length transform) and split:
//   X[k] = E[k] + w^k O[k],  E = (Z[k] + conj Z[M-k]) / 2,  O = (Z[k] - conj Z[M-k]) / 2i,  w = e^{-2 pi i / n}
// pl is the plan of M, tw the table of n entries. A warp owns whole lines in the fill / drain loops.
template <typename R>
__global__ void __launch_bounds__(kMixThreads)
mix_r2c_even_kernel(typename Cx<R>::T* __restrict__ out, const R* __restrict__ in, long long rows, MixPlan pl,
                    int Tsh, const typename Cx<R>::T* __restrict__ tw, R scale) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char mix_smem[];
  const int M = pl.n, n = 2 * M, nc = M + 1, pitch = M | 1, T = 1 << Tsh;
  C* a = reinterpret_cast<C*>(mix_smem);
  C* b = a + (size_t)T * pitch;
  const long long row0 = (long long)blockIdx.x * T;
  const int lvalid = (int)((rows - row0 < T) ? (rows - row0) : T);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int l = warp; l < T; l += kMixThreads / 32) {
    const C* src = reinterpret_cast<const C*>(in + (row0 + l) * n);
    for (int q = lane; q < M; q += 32) {
      C v;
      v.x = v.y = R(0);
      if (l < lvalid) v = src[q];
      a[l * pitch + q] = v;
    }
  }
  __syncthreads();
  const C* Z = mix_fft<R, false>(a, b, pl, 1, pitch, Tsh, tw, tid, kMixThreads, 2);
  const R hs = R(0.5) * scale;
  for (int l = warp; l < lvalid; l += kMixThreads / 32) {
    const C* z = Z + l * pitch;
    C* dst = out + (row0 + l) * nc;
    for (int k = lane; k <= M; k += 32) {
      const C zk = z[k == M ? 0 : k], zm = z[k == 0 ? 0 : M - k];
      C E, O, w, o;
      E.x = zk.x + zm.x;            // 2E = Z[k] + conj Z[M-k]
      E.y = zk.y - zm.y;
      O.x = zk.y + zm.y;            // 2O = (Z[k] - conj Z[M-k]) / i
      O.y = zm.x - zk.x;
      if (k == M) { w.x = R(-1); w.y = R(0); } else { w = tw[k]; }
      o.x = (E.x + (w.x * O.x - w.y * O.y)) * hs;
      o.y = (E.y + (w.x * O.y + w.y * O.x)) * hs;
      dst[k] = o;
    }
  }
}

// inverse of the above: Z'[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) conj(w^k), z = IDFT_M(Z'),
// x[2j] = Re z[j], x[2j+1] = Im z[j]; the imaginary parts of X[0] and X[M] are ignored (C2R semantics).
template <typename R>
__global__ void __launch_bounds__(kMixThreads)
mix_c2r_even_kernel(R* __restrict__ out, const typename Cx<R>::T* __restrict__ in, long long rows, MixPlan pl,
                    int Tsh, const typename Cx<R>::T* __restrict__ tw, R scale) {
  using C = typename Cx<R>::T;
  extern __shared__ __align__(16) unsigned char mix_smem[];
  const int M = pl.n, n = 2 * M, nc = M + 1, pitch = M | 1, T = 1 << Tsh;
  C* a = reinterpret_cast<C*>(mix_smem);
  C* b = a + (size_t)T * pitch;
  const long long row0 = (long long)blockIdx.x * T;
  const int lvalid = (int)((rows - row0 < T) ? (rows - row0) : T);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int l = warp; l < T; l += kMixThreads / 32) {
    const C* src = in + (row0 + l) * nc;
    for (int k = lane; k < M; k += 32) {
      C v;
      v.x = v.y = R(0);
      if (l < lvalid) {
        C xk = src[k], xm = src[M - k];
        if (k == 0) { xk.y = R(0); xm.y = R(0); }
        C s2, d2;
        s2.x = xk.x + xm.x;          // X[k] + conj X[M-k]
        s2.y = xk.y - xm.y;
        d2.x = xk.x - xm.x;          // X[k] - conj X[M-k]
        d2.y = xk.y + xm.y;
        const C w = tw[k];           // conj(w^k) = (w.x, -w.y); i * d2 * conj(w)
        const R tx = d2.x * w.x + d2.y * w.y, ty = d2.y * w.x - d2.x * w.y;
        v.x = s2.x - ty;
        v.y = s2.y + tx;
      }
      a[l * pitch + k] = v;
    }
  }
  __syncthreads();
  const C* z = mix_fft<R, true>(a, b, pl, 1, pitch, Tsh, tw, tid, kMixThreads, 2);
  for (int l = warp; l < lvalid; l += kMixThreads / 32) {
    C* dst = reinterpret_cast<C*>(out + (row0 + l) * n);
    for (int q = lane; q < M; q += 32) {
      C v = z[l * pitch + q];
      v.x *= scale;
      v.y *= scale;
      dst[q] = v;
    }
  }
}

// The first-axis pass with the Fourier multiplier inside: forward transform along X, `mult` on every
// (frequency row, line) of the NCH channels held together, inverse transform, in place. `mult(q, l, v)`
// gets the X frequency q, the line l of the tile and the NCH complex values.
template <typename R, int NCH, typename F>
__device__ __forceinline__ void mix_xmid(typename Cx<R>::T* __restrict__ data, long long chs, long long st, int lvalid,
                                         const MixPlan& pl, int Tsh, const typename Cx<R>::T* __restrict__ tw,
                                         typename Cx<R>::T* smem, F mult) {
  using C = typename Cx<R>::T;
  const int n = pl.n, T = 1 << Tsh, tid = threadIdx.x;
  const int tile = T * n;
  C* a = smem;
  C* b = smem + (size_t)NCH * tile;
  for (int c = 0; c < NCH; ++c)
    for (int i = tid; i < tile; i += kMixThreads) {
      const int l = i & (T - 1), q = i >> Tsh;
      C v;
      v.x = v.y = R(0);
      if (l < lvalid) v = data[c * chs + (long long)q * st + l];
      a[c * tile + i] = v;
    }
  __syncthreads();
  C* res = a;
  for (int c = 0; c < NCH; ++c) {
    C* r = mix_fft<R, false>(a + c * tile, b + c * tile, pl, T, 1, Tsh, tw, tid, kMixThreads);
    res = r - c * tile;
  }
  C* oth = (res == a) ? b : a;
  for (int i = tid; i < tile; i += kMixThreads) {
    C v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) v[c] = res[c * tile + i];
    mult(i >> Tsh, i & (T - 1), v);
#pragma unroll
    for (int c = 0; c < NCH; ++c) res[c * tile + i] = v[c];
  }
  __syncthreads();
  C* fin = res;
  for (int c = 0; c < NCH; ++c) {
    C* r = mix_fft<R, true>(res + c * tile, oth + c * tile, pl, T, 1, Tsh, tw, tid, kMixThreads);
    fin = r - c * tile;
  }
  for (int c = 0; c < NCH; ++c)
    for (int i = tid; i < tile; i += kMixThreads) {
      const int l = i & (T - 1), q = i >> Tsh;
      if (l < lvalid) data[c * chs + (long long)q * st + l] = fin[c * tile + i];
    }
}

}  // namespace lgm
