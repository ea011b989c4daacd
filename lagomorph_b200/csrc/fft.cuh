// fft.cuh -- in-shared-memory batched "column" FFTs for the FluidMetric passes.
//
// A tile holds L independent lines of N complex points: element (r, l) lives at
// tile[r*rs + l*ls]. Work items are (line, sub-butterfly) pairs with the LINE
// index fastest across lanes, so for ls == 1 every shared-memory access of a
// warp is a run of consecutive complex words (bank-conflict free), and for the
// transposed view (rs == 1, ls odd) the odd line pitch does the same job.
//
// N = r_1 * ... * r_k with radices in {2,4,8,16}; each stage is an r-point DFT
// done entirely in registers (radix-2 DIF network, compile-time twiddles),
// followed by the inter-stage twiddle, stored back IN PLACE. The forward
// transform therefore leaves the spectrum in mixed-radix digit-reversed order
// (fft_pos() gives the storage row of a frequency); the inverse consumes that
// order and produces natural order. No reordering pass exists anywhere: the
// Fourier multiplier looks its LUTs up in storage order.
#pragma once
#include "common.cuh"

namespace lgm {

template <typename R> struct Cx;
template <> struct Cx<float> { using T = float2; };
template <> struct Cx<double> { using T = double2; };

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// fp32 complex add/sub as ONE packed instruction (Blackwell FADD2: two IEEE fp32 adds per issue
// slot, same rounding as the scalar pair). The butterflies are mostly complex adds, so this
// removes ~40 % of the FFT passes' arithmetic issue slots.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
// v -> fl(fl(post * v) + 0): what `ds*u + dt*interp(zeros)` of the compose stage leaves (gather3.cu)
__device__ __forceinline__ float post_scale(float post, float v) { return __fadd_rn(__fmul_rn(post, v), 0.f); }
__device__ __forceinline__ double post_scale(double post, double v) { return __dadd_rn(__dmul_rn(post, v), 0.0); }

template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {  // a * conj(b)
  C r;
  r.x = a.x * b.x + a.y * b.y;
  r.y = a.y * b.x - a.x * b.y;
  return r;
}

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }
// bits of stage i when 2^b points are split into ceil(b/4) stages as evenly as possible
__host__ __device__ constexpr int stage_bits(int b, int i) {
  return b / ((b + 3) / 4) + (i < b % ((b + 3) / 4) ? 1 : 0);
}
__host__ __device__ constexpr int bitrev(int i, int bits) {
  int r = 0;
  for (int k = 0; k < bits; ++k) r |= ((i >> k) & 1) << (bits - 1 - k);
  return r;
}

// storage row of frequency k after the in-place forward transform of 2^b points
template <int N>
__host__ __device__ __forceinline__ int fft_pos(int k) {
  constexpr int b = ilog2(N);
  constexpr int ns = (b + 3) / 4;
  int pos = 0, size = N;
#pragma unroll
  for (int s = 0; s < ns; ++s) {
    const int r = 1 << stage_bits(b, s);
    size /= r;
    pos += (k % r) * size;
    k /= r;
  }
  return pos;
}
// runtime-N twin for the host-side LUT builder
inline int fft_pos_rt(int N, int k) {
  const int b = ilog2(N), ns = (b + 3) / 4;
  int pos = 0, size = N;
  for (int s = 0; s < ns; ++s) {
    const int r = 1 << stage_bits(b, s);
    size /= r;
    pos += (k % r) * size;
    k /= r;
  }
  return pos;
}

// multiply by e^{-+ 2 pi i j/16} (forward: minus), j compile-time after unrolling
template <bool INV, typename C>
__device__ __forceinline__ C mul_w16(C t, int j) {
  using R = decltype(t.x);
  const R c1 = R(0.92387953251128675613), s1 = R(0.38268343236508977173), c2 = R(0.70710678118654752440);
  C r;
  R wr, wi;  // w = wr - i*wi (forward), wr + i*wi (inverse)
  switch (j) {
    case 0: return t;
    case 4:  // -+ i
      if (!INV) { r.x = t.y; r.y = -t.x; } else { r.x = -t.y; r.y = t.x; }
      return r;
    case 2:
      if (!INV) { r.x = c2 * (t.x + t.y); r.y = c2 * (t.y - t.x); }
      else { r.x = c2 * (t.x - t.y); r.y = c2 * (t.y + t.x); }
      return r;
    case 6:
      if (!INV) { r.x = c2 * (t.y - t.x); r.y = -c2 * (t.x + t.y); }
      else { r.x = -c2 * (t.x + t.y); r.y = c2 * (t.x - t.y); }
      return r;
    case 1: wr = c1; wi = s1; break;
    case 3: wr = s1; wi = c1; break;
    case 5: wr = -s1; wi = c1; break;
    default: wr = -c1; wi = s1; break;  // 7
  }
  if (!INV) { r.x = t.x * wr + t.y * wi; r.y = t.y * wr - t.x * wi; }
  else { r.x = t.x * wr - t.y * wi; r.y = t.y * wr + t.x * wi; }
  return r;
}

// multiply by e^{-+ 2 pi i j/32}, j in 0..15 compile-time after unrolling (even j: the mul_w16 cases,
// so radices <= 16 are unchanged); used by the 32-point register DFT of the quarter-slab X pass
template <bool INV, typename C>
__device__ __forceinline__ C mul_w32(C t, int j) {
  using R = decltype(t.x);
  if ((j & 1) == 0) return mul_w16<INV>(t, j >> 1);
  constexpr double cs[8] = {0.98078528040323044913, 0.83146961230254523708, 0.55557023301960222474,
                            0.19509032201612826785, -0.19509032201612826785, -0.55557023301960222474,
                            -0.83146961230254523708, -0.98078528040323044913};   // cos(2 pi j/32), j = 1,3,..,15
  constexpr double sn[8] = {0.19509032201612826785, 0.55557023301960222474, 0.83146961230254523708,
                            0.98078528040323044913, 0.98078528040323044913, 0.83146961230254523708,
                            0.55557023301960222474, 0.19509032201612826785};    // sin(2 pi j/32)
  const R wr = R(cs[(j >> 1) & 7]), wi = R(sn[(j >> 1) & 7]);
  C r;
  if (!INV) { r.x = t.x * wr + t.y * wi; r.y = t.y * wr - t.x * wi; }
  else { r.x = t.x * wr - t.y * wi; r.y = t.y * wr + t.x * wi; }
  return r;
}

// RAD-point DFT in registers (radix-2 DIF). On return x[i] holds output bitrev(i).
template <int RAD, bool INV, typename C>
__device__ __forceinline__ void reg_fft(C (&x)[RAD]) {
  static_assert(RAD <= 32, "register DFT of at most 32 points");
#pragma unroll
  for (int half = RAD / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int base = 0; base < RAD; base += 2 * half) {
#pragma unroll
      for (int i = 0; i < half; ++i) {
        C a = x[base + i], b = x[base + i + half];
        x[base + i] = cadd(a, b);
        if constexpr (RAD == 32) x[base + i + half] = mul_w32<INV>(csub(a, b), i * (16 / half));
        else x[base + i + half] = mul_w16<INV>(csub(a, b), i * (8 / half));
      }
    }
  }
}

// One in-place stage over blocks of BLOCK consecutive rows (BLOCK | N).
// tw: table of N entries, tw[j] = e^{-2 pi i j / N}.
// Global-memory side of a stage: element (row r, line l) lives at g[r*grs + l] (lines are
// consecutive words, so lanes running over l coalesce); lines l >= lvalid are padding.
template <typename C>
struct GSide {
  C* g;
  long long grs;
  int lvalid;
  int lsplit = 1 << 30;  // lines l >= lsplit sit lskip words further on (the Nyquist line of a
  int lskip = 0;         // cluster slab CTA); fft_stage only
};

// GSRC: the stage's inputs come from global memory instead of the tile; GDST: its outputs go to
// global memory instead of the tile (used for the first / last stage of a pass, so the tile is
// never filled or drained by a separate copy loop).
// JSH / jump: tile row r lives (r >> JSH) * jump words further on than r*rs (a tile stored as panels
// of 2^JSH rows, see the cluster slab kernels in fluid.cu); JSH = 0 means a plain tile.
template <typename R, int N, int BLOCK, int RAD, int L, bool INV, bool GSRC = false, bool GDST = false, int JSH = 0>
__device__ __forceinline__ void fft_stage(typename Cx<R>::T* tile, int rs, int ls,
                                          const typename Cx<R>::T* __restrict__ tw, int tid, int nth,
                                          GSide<typename Cx<R>::T> gs = GSide<typename Cx<R>::T>(),
                                          int jump = 0) {
  using C = typename Cx<R>::T;
  constexpr int SUB = BLOCK / RAD;
  constexpr int ITEMS = L * (N / RAD);
  constexpr int BITS = ilog2(RAD);
#ifndef LGM_GSRC_UNROLL
#define LGM_GSRC_UNROLL 1
#endif
  constexpr int kUnroll = GSRC ? LGM_GSRC_UNROLL : 1;  // items whose global loads are in flight together
#pragma unroll kUnroll
  for (int it = tid; it < ITEMS; it += nth) {
    const int l = it % L, q = it / L;
    const int blk = q / SUB, rest = q % SUB;
    const int row0 = blk * BLOCK + rest;
    C* p = tile + row0 * rs + l * ls;
    // panel offset of tile row (row0 + n*SUB): compile-time per n in the outermost stage
    auto jo = [&](int n) -> int { return JSH > 0 ? ((row0 + n * SUB) >> JSH) * jump : 0; };
    C* gp = (GSRC || GDST) ? gs.g + (long long)row0 * gs.grs + l + (l >= gs.lsplit ? gs.lskip : 0) : nullptr;
    const bool ok = !(GSRC || GDST) || l < gs.lvalid;
    C x[RAD];
    if (!INV) {
#pragma unroll
      for (int n = 0; n < RAD; ++n) {
        if (GSRC) {
          C v;
          v.x = v.y = R(0);
          if (ok) v = gp[(long long)n * SUB * gs.grs];
          x[n] = v;
        } else {
          x[n] = p[n * SUB * rs + jo(n)];
        }
      }
      reg_fft<RAD, false>(x);
#pragma unroll
      for (int i = 0; i < RAD; ++i) {
        const int k = bitrev(i, BITS);
        C v = x[i];
        if (SUB > 1 && k != 0) v = cmul(v, tw[rest * k * (N / BLOCK)]);
        if (GDST) {
          if (ok) gp[(long long)k * SUB * gs.grs] = v;
        } else {
          p[k * SUB * rs + jo(k)] = v;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < RAD; ++k) {
        C v;
        if (GSRC) {
          v.x = v.y = R(0);
          if (ok) v = gp[(long long)k * SUB * gs.grs];
        } else {
          v = p[k * SUB * rs + jo(k)];
        }
        if (SUB > 1 && k != 0) v = cmulc(v, tw[rest * k * (N / BLOCK)]);
        x[k] = v;
      }
      reg_fft<RAD, true>(x);
#pragma unroll
      for (int i = 0; i < RAD; ++i) {
        if (GDST) {
          if (ok) gp[(long long)bitrev(i, BITS) * SUB * gs.grs] = x[i];
        } else {
          p[bitrev(i, BITS) * SUB * rs + jo(bitrev(i, BITS))] = x[i];
        }
      }
    }
  }
}

// Innermost forward stage, a per-element multiplier and the innermost inverse stage fused in
// registers. The last forward stage (BLOCK == RAD) transforms RAD consecutive tile rows of one line;
// the first executed inverse stage reads exactly those rows again, so a Fourier multiplier that sits
// between the two transforms (the X pass of the fluid operator) needs no shared-memory round trip
// and no barrier on either side of it: rows in -> DFT -> mult(row, v) -> inverse DFT -> rows out.
// Arithmetic and its order are those of the three separate steps (bit-identical results).
template <typename R, int N, int RAD, int L, typename F>
__device__ __forceinline__ void fft_mid_stage(typename Cx<R>::T* tile, int rs, int ls, int tid, int nth, F mult) {
  using C = typename Cx<R>::T;
  constexpr int ITEMS = L * (N / RAD);
  constexpr int BITS = ilog2(RAD);
  for (int it = tid; it < ITEMS; it += nth) {
    const int l = it % L, blk = it / L;
    C* p = tile + blk * RAD * rs + l * ls;
    C x[RAD], y[RAD];
#pragma unroll
    for (int n = 0; n < RAD; ++n) x[n] = p[n * rs];
    reg_fft<RAD, false>(x);  // x[i] = frequency row blk*RAD + bitrev(i)
#pragma unroll
    for (int k = 0; k < RAD; ++k) y[k] = mult(blk * RAD + k, l, x[bitrev(k, BITS)]);
    reg_fft<RAD, true>(y);
#pragma unroll
    for (int i = 0; i < RAD; ++i) p[bitrev(i, BITS) * rs] = y[i];
  }
}

// fft_mid_stage for NCH coupled channels (tiles `chs` words apart): the multiplier sees the NCH values
// of one (row, line) together (the matrix-valued Fourier symbol of the fluid operator with beta != 0).
template <typename R, int N, int RAD, int L, int NCH, typename F>
__device__ __forceinline__ void fft_mid_stage_multi(typename Cx<R>::T* tile, int chs, int rs, int ls, int tid,
                                                    int nth, F mult) {
  using C = typename Cx<R>::T;
  constexpr int ITEMS = L * (N / RAD);
  constexpr int BITS = ilog2(RAD);
  for (int it = tid; it < ITEMS; it += nth) {
    const int l = it % L, blk = it / L;
    C* p = tile + blk * RAD * rs + l * ls;
    C x[NCH][RAD];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int n = 0; n < RAD; ++n) x[c][n] = p[c * chs + n * rs];
      reg_fft<RAD, false>(x[c]);
    }
    C y[NCH][RAD];
#pragma unroll
    for (int k = 0; k < RAD; ++k) {
      C v[NCH];
#pragma unroll
      for (int c = 0; c < NCH; ++c) v[c] = x[c][bitrev(k, BITS)];
      mult(blk * RAD + k, l, v);
#pragma unroll
      for (int c = 0; c < NCH; ++c) y[c][k] = v[c];
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      reg_fft<RAD, true>(y[c]);
#pragma unroll
      for (int i = 0; i < RAD; ++i) p[c * chs + bitrev(i, BITS) * rs] = y[c][i];
    }
  }
}

// Outermost stage (BLOCK == N) with one side in GLOBAL memory: the forward transform's first
// stage reads its RAD inputs straight from global memory (row stride grs, lanes along l are
// consecutive words => coalesced) and stores to the shared tile; the inverse transform's last
// stage reads the tile and stores straight to global. Columns l >= lvalid are padding.
template <typename R, int N, int RAD, int L, bool INV>
__device__ __forceinline__ void fft_stage_edge(typename Cx<R>::T* __restrict__ g, long long grs,
                                               typename Cx<R>::T* tile, int rs, int ls,
                                               const typename Cx<R>::T* __restrict__ tw, int lvalid,
                                               int tid, int nth) {
  using C = typename Cx<R>::T;
  constexpr int SUB = N / RAD;
  constexpr int ITEMS = L * SUB;
  constexpr int BITS = ilog2(RAD);
  for (int it = tid; it < ITEMS; it += nth) {
    const int l = it % L, rest = it / L;
    C* p = tile + rest * rs + l * ls;
    C* gp = g + (long long)rest * grs + l;
    const bool ok = l < lvalid;
    C x[RAD];
    if (!INV) {
#pragma unroll
      for (int n = 0; n < RAD; ++n) {
        C v;
        v.x = v.y = R(0);
        if (ok) v = gp[(long long)n * SUB * grs];
        x[n] = v;
      }
      reg_fft<RAD, false>(x);
#pragma unroll
      for (int i = 0; i < RAD; ++i) {
        const int k = bitrev(i, BITS);
        C v = x[i];
        if (SUB > 1 && k != 0) v = cmul(v, tw[rest * k]);
        p[k * SUB * rs] = v;
      }
    } else {
#pragma unroll
      for (int k = 0; k < RAD; ++k) {
        C v = p[k * SUB * rs];
        if (SUB > 1 && k != 0) v = cmulc(v, tw[rest * k]);
        x[k] = v;
      }
      reg_fft<RAD, true>(x);
      if (ok) {
#pragma unroll
        for (int i = 0; i < RAD; ++i) gp[(long long)bitrev(i, BITS) * SUB * grs] = x[i];
      }
    }
  }
}

// forward: global -> (first stage) -> tile -> remaining stages in the tile
template <typename R, int N, int L>
__device__ __forceinline__ void col_fft_fwd_from_global(typename Cx<R>::T* g, long long grs,
                                                        typename Cx<R>::T* tile, int rs, int ls,
                                                        const typename Cx<R>::T* tw, int lvalid,
                                                        int tid, int nth);
// inverse: tile -> inner stages -> (last stage) -> global
template <typename R, int N, int L>
__device__ __forceinline__ void col_fft_inv_to_global(typename Cx<R>::T* g, long long grs,
                                                      typename Cx<R>::T* tile, int rs, int ls,
                                                      const typename Cx<R>::T* tw, int lvalid,
                                                      int tid, int nth);

template <typename R, int N, int BLOCK, int SI, int L>
struct ColFFT {
  using C = typename Cx<R>::T;
  static constexpr int RAD = 1 << stage_bits(ilog2(N), SI);
  static constexpr bool LASTSTAGE = (BLOCK / RAD == 1);
  static __device__ __forceinline__ void fwd(C* tile, int rs, int ls, const C* tw, int tid, int nth) {
    fft_stage<R, N, BLOCK, RAD, L, false>(tile, rs, ls, tw, tid, nth);
    if constexpr (BLOCK / RAD > 1) {
      __syncthreads();
      ColFFT<R, N, BLOCK / RAD, SI + 1, L>::fwd(tile, rs, ls, tw, tid, nth);
    }
  }
  static __device__ __forceinline__ void inv(C* tile, int rs, int ls, const C* tw, int tid, int nth) {
    if constexpr (BLOCK / RAD > 1) {
      ColFFT<R, N, BLOCK / RAD, SI + 1, L>::inv(tile, rs, ls, tw, tid, nth);
      __syncthreads();
    }
    fft_stage<R, N, BLOCK, RAD, L, true>(tile, rs, ls, tw, tid, nth);
  }
  // all stages except the innermost one (which real_edge_stage() fuses with the real split)
  static __device__ __forceinline__ void fwd_nolast(C* tile, int rs, int ls, const C* tw, int tid, int nth) {
    if constexpr (!LASTSTAGE) {
      fft_stage<R, N, BLOCK, RAD, L, false>(tile, rs, ls, tw, tid, nth);
      if constexpr (!ColFFT<R, N, BLOCK / RAD, SI + 1, L>::LASTSTAGE) {
        __syncthreads();
        ColFFT<R, N, BLOCK / RAD, SI + 1, L>::fwd_nolast(tile, rs, ls, tw, tid, nth);
      }
    }
  }
  static __device__ __forceinline__ void inv_nolast(C* tile, int rs, int ls, const C* tw, int tid, int nth) {
    if constexpr (!LASTSTAGE) {
      if constexpr (!ColFFT<R, N, BLOCK / RAD, SI + 1, L>::LASTSTAGE) {
        ColFFT<R, N, BLOCK / RAD, SI + 1, L>::inv_nolast(tile, rs, ls, tw, tid, nth);
        __syncthreads();
      }
      fft_stage<R, N, BLOCK, RAD, L, true>(tile, rs, ls, tw, tid, nth);
    }
  }
  // variants whose first executed stage reads global memory (GIN) and/or whose last executed
  // stage writes global memory (GOUT); gin / gout describe those arrays.
  template <bool GIN, bool GOUT>
  static __device__ __forceinline__ void fwd_g(C* tile, int rs, int ls, const C* tw, int tid, int nth,
                                               GSide<C> gin, GSide<C> gout) {
    constexpr bool first = (SI == 0);
    if constexpr (first && GIN && LASTSTAGE && GOUT) {
      // single-stage transform: global -> registers -> global (gin and gout may be the same array)
      fft_stage<R, N, BLOCK, RAD, L, false, true, false>(tile, rs, ls, tw, tid, nth, gin);
      __syncthreads();
      for (int it = tid; it < L * N; it += nth) {
        const int l = it % L, r = it / L;
        if (l < gout.lvalid) gout.g[(long long)r * gout.grs + l] = tile[r * rs + l * ls];
      }
    } else {
      fft_stage<R, N, BLOCK, RAD, L, false, first && GIN, LASTSTAGE && GOUT>(
          tile, rs, ls, tw, tid, nth, (first && GIN) ? gin : gout);
      if constexpr (!LASTSTAGE) {
        __syncthreads();
        ColFFT<R, N, BLOCK / RAD, SI + 1, L>::template fwd_g<GIN, GOUT>(tile, rs, ls, tw, tid, nth, gin, gout);
      }
    }
  }
  template <bool GIN, bool GOUT>
  static __device__ __forceinline__ void inv_g(C* tile, int rs, int ls, const C* tw, int tid, int nth,
                                               GSide<C> gin, GSide<C> gout) {
    constexpr bool last = (SI == 0);  // executed last in the inverse
    if constexpr (last && GOUT && LASTSTAGE && GIN) {
      for (int it = tid; it < L * N; it += nth) {
        const int l = it % L, r = it / L;
        C v;
        v.x = v.y = R(0);
        if (l < gin.lvalid) v = gin.g[(long long)r * gin.grs + l];
        tile[r * rs + l * ls] = v;
      }
      __syncthreads();
      fft_stage<R, N, BLOCK, RAD, L, true, false, true>(tile, rs, ls, tw, tid, nth, gout);
    } else {
      if constexpr (!LASTSTAGE) {
        ColFFT<R, N, BLOCK / RAD, SI + 1, L>::template inv_g<GIN, GOUT>(tile, rs, ls, tw, tid, nth, gin, gout);
        __syncthreads();
      }
      fft_stage<R, N, BLOCK, RAD, L, true, LASTSTAGE && GIN, last && GOUT>(
          tile, rs, ls, tw, tid, nth, (LASTSTAGE && GIN) ? gin : gout);
    }
  }
};

// Forward / inverse (unnormalised) FFT of L lines of N points. Callers
// __syncthreads() before (tile + tw visible) and after.
template <typename R, int N, int L>
__device__ __forceinline__ void col_fft_fwd(typename Cx<R>::T* tile, int rs, int ls,
                                            const typename Cx<R>::T* tw, int tid, int nth) {
  if constexpr (N > 1) ColFFT<R, N, N, 0, L>::fwd(tile, rs, ls, tw, tid, nth);
}
template <typename R, int N, int L>
__device__ __forceinline__ void col_fft_inv(typename Cx<R>::T* tile, int rs, int ls,
                                            const typename Cx<R>::T* tw, int tid, int nth) {
  if constexpr (N > 1) ColFFT<R, N, N, 0, L>::inv(tile, rs, ls, tw, tid, nth);
}

}  // namespace lgm

namespace lgm {

template <typename R, int N, int L>
__device__ __forceinline__ void col_fft_fwd_from_global(typename Cx<R>::T* g, long long grs,
                                                        typename Cx<R>::T* tile, int rs, int ls,
                                                        const typename Cx<R>::T* tw, int lvalid,
                                                        int tid, int nth) {
  constexpr int RAD = 1 << stage_bits(ilog2(N), 0);
  fft_stage_edge<R, N, RAD, L, false>(g, grs, tile, rs, ls, tw, lvalid, tid, nth);
  if constexpr (N / RAD > 1) {
    __syncthreads();
    ColFFT<R, N, N / RAD, 1, L>::fwd(tile, rs, ls, tw, tid, nth);
  }
}

// ---------------------------------------------------------------------------------------------
// Real transform of 2M points through a complex FFT of M points on (x[2j], x[2j+1]) pairs, with the
// split / unsplit folded into the LAST forward / FIRST inverse radix stage.
//
// In the last forward stage a block b (RAD consecutive rows) holds the frequencies
// k = klow(b) + NB*k2, k2 < RAD, NB = M/RAD. The split pairs k with M-k, whose block has
// klow' = (NB - klow) mod NB and k2' = RAD-1-k2 (k2' = RAD-k2 when klow == 0). A work item therefore
// transforms the two partner blocks together and has every (k, M-k) pair in registers:
//   X[k]   = 1/2[(Zk + conj Zm) - i W^k (Zk - conj Zm)],  Zm = Z[M-k], W = e^{-2 pi i / 2M}
// so the separate split pass (one more shared-memory round trip plus a barrier) disappears.
// The tile needs M+1 rows: row M receives the Nyquist word X[M].
template <int N>
__host__ __device__ __forceinline__ int fft_freq(int pos) {  // inverse of fft_pos
  constexpr int b = ilog2(N);
  constexpr int ns = (b + 3) / 4;
  int k = 0, mult = 1, size = N;
#pragma unroll
  for (int s = 0; s < ns; ++s) {
    const int r = 1 << stage_bits(b, s);
    size /= r;
    const int d = pos / size;
    pos -= d * size;
    k += d * mult;
    mult *= r;
  }
  return k;
}

template <typename C>
__device__ __forceinline__ void split_pair(C a, C b, C w, C& xk, C& xm) {
  using R = decltype(a.x);
  C s, d, t, w2, d2, t2;
  s.x = a.x + b.x; s.y = a.y - b.y;
  d.x = a.x - b.x; d.y = a.y + b.y;
  t = cmul(w, d);
  xk.x = R(0.5) * (s.x + t.y);
  xk.y = R(0.5) * (s.y - t.x);
  w2.x = -w.x; w2.y = w.y;
  d2.x = -d.x; d2.y = d.y;
  t2 = cmul(w2, d2);
  xm.x = R(0.5) * (s.x + t2.y);
  xm.y = R(0.5) * (-s.y - t2.x);
}
template <typename C>
__device__ __forceinline__ void unsplit_pair(C a, C b, C w, C& zk, C& zm) {
  C s, d, t, d2, t2;
  s.x = a.x + b.x; s.y = a.y - b.y;
  d.x = a.x - b.x; d.y = a.y + b.y;
  t = cmulc(d, w);
  zk.x = s.x - t.y;
  zk.y = s.y + t.x;
  d2.x = -d.x; d2.y = d.y;
  t2 = cmul(d2, w);
  zm.x = s.x + t2.y;
  zm.y = -s.y - t2.x;
}

// twz: table of 2M entries e^{-2 pi i j / 2M}. Rows of the tile: element (r, l) at tile[r*rs + l*ls].
// JSH / jump as in fft_stage; nyq = offset of the Nyquist row (row M) in the tile, -1 = M*rs.
template <typename R, int M, int L, bool INV, int JSH = 0>
__device__ __forceinline__ void real_edge_stage(typename Cx<R>::T* tile, int rs, int ls,
                                                const typename Cx<R>::T* __restrict__ twz, int tid, int nth,
                                                int jump = 0, int nyq = -1) {
  using C = typename Cx<R>::T;
  constexpr int bM = ilog2(M), ns = (bM + 3) / 4;
  constexpr int RAD = 1 << stage_bits(bM, ns - 1);
  constexpr int NB = M / RAD;
  constexpr int NP = NB / 2 + 1;
  constexpr int BITS = ilog2(RAD);
  for (int it = tid; it < L * NP; it += nth) {
    const int l = it % L, kl = it / L;
    const int klp = (NB - kl) % NB;
    const int b0 = fft_pos<NB>(kl), b1 = fft_pos<NB>(klp);
    static_assert(JSH == 0 || (1 << JSH) % RAD == 0, "a radix block must not straddle panels");
    C* p0 = tile + (b0 * RAD) * rs + l * ls + (JSH > 0 ? ((b0 * RAD) >> JSH) * jump : 0);
    C* p1 = tile + (b1 * RAD) * rs + l * ls + (JSH > 0 ? ((b1 * RAD) >> JSH) * jump : 0);
    C* pn = tile + (nyq < 0 ? M * rs : nyq) + l * ls;  // Nyquist word of this line
    C x[RAD], y[RAD];
    if (!INV) {
#pragma unroll
      for (int n = 0; n < RAD; ++n) x[n] = p0[n * rs];
      reg_fft<RAD, false>(x);  // x[bitrev(k2)] = Z[kl + NB*k2]
      if (kl == 0) {
        C a = x[0], x0, xM;
        x0.x = a.x + a.y; x0.y = R(0);
        xM.x = a.x - a.y; xM.y = R(0);
        p0[0] = x0;
        *pn = xM;
#pragma unroll
        for (int k2 = 1; k2 <= RAD / 2; ++k2) {
          C xk, xm;
          split_pair(x[bitrev(k2, BITS)], x[bitrev(RAD - k2, BITS)], twz[NB * k2], xk, xm);
          p0[k2 * rs] = xk;
          if (k2 != RAD - k2) p0[(RAD - k2) * rs] = xm;
        }
      } else if (klp == kl) {
#pragma unroll
        for (int k2 = 0; k2 < RAD / 2; ++k2) {
          C xk, xm;
          split_pair(x[bitrev(k2, BITS)], x[bitrev(RAD - 1 - k2, BITS)], twz[kl + NB * k2], xk, xm);
          p0[k2 * rs] = xk;
          p0[(RAD - 1 - k2) * rs] = xm;
        }
      } else {
#pragma unroll
        for (int n = 0; n < RAD; ++n) y[n] = p1[n * rs];
        reg_fft<RAD, false>(y);
#pragma unroll
        for (int k2 = 0; k2 < RAD; ++k2) {
          C xk, xm;
          split_pair(x[bitrev(k2, BITS)], y[bitrev(RAD - 1 - k2, BITS)], twz[kl + NB * k2], xk, xm);
          p0[k2 * rs] = xk;
          p1[(RAD - 1 - k2) * rs] = xm;
        }
      }
    } else {
      // unsplit Z'[k] = (Xk + conj Xm) + i conj(W^k)(Xk - conj Xm), then the inverse RAD-point stage
      if (kl == 0) {
        const R r0 = p0[0].x, rM = pn->x;  // imaginary parts of DC / Nyquist ignored (C2R)
        x[0].x = r0 + rM;
        x[0].y = r0 - rM;
#pragma unroll
        for (int k2 = 1; k2 <= RAD / 2; ++k2) {
          C zk, zm;
          unsplit_pair(p0[k2 * rs], p0[(RAD - k2) * rs], twz[NB * k2], zk, zm);
          x[k2] = zk;
          if (k2 != RAD - k2) x[RAD - k2] = zm;
        }
        reg_fft<RAD, true>(x);
#pragma unroll
        for (int i = 0; i < RAD; ++i) p0[bitrev(i, BITS) * rs] = x[i];
      } else if (klp == kl) {
#pragma unroll
        for (int k2 = 0; k2 < RAD / 2; ++k2) {
          C zk, zm;
          unsplit_pair(p0[k2 * rs], p0[(RAD - 1 - k2) * rs], twz[kl + NB * k2], zk, zm);
          x[k2] = zk;
          x[RAD - 1 - k2] = zm;
        }
        reg_fft<RAD, true>(x);
#pragma unroll
        for (int i = 0; i < RAD; ++i) p0[bitrev(i, BITS) * rs] = x[i];
      } else {
#pragma unroll
        for (int k2 = 0; k2 < RAD; ++k2) {
          C zk, zm;
          unsplit_pair(p0[k2 * rs], p1[(RAD - 1 - k2) * rs], twz[kl + NB * k2], zk, zm);
          x[k2] = zk;
          y[RAD - 1 - k2] = zm;
        }
        reg_fft<RAD, true>(x);
        reg_fft<RAD, true>(y);
#pragma unroll
        for (int i = 0; i < RAD; ++i) {
          p0[bitrev(i, BITS) * rs] = x[i];
          p1[bitrev(i, BITS) * rs] = y[i];
        }
      }
    }
  }
}

// Outermost stage of the Z (contiguous-axis) transform with its global side along the LINE: the L
// lines of M complex words are rows of a row-major global array (line stride = M words), the tile is
// transposed (element (r, l) at tile[r*rs + l], l = line). Lanes run over `rest` first (consecutive
// words of one line: SUB*8-byte runs, every fetched sector fully used) and then over lines spaced SUB
// apart, which with a pitch rs = 1 (mod 16) makes the 64-bit shared-memory accesses of every
// half-warp hit 16 distinct bank pairs. Forward: global -> registers -> tile; inverse: tile ->
// registers -> global. Removes the separate fill / drain loop of the Z passes (one shared-memory
// write + read of the whole tile). Needs L % 32 == 0. JSH / jump as in fft_stage.
// POST (inverse only): every stored real is scaled by `post` on its way out (post_scale).
template <typename R, int M, int RAD, int L, bool INV, int JSH = 0, bool POST = false>
__device__ __forceinline__ void zedge_stage(typename Cx<R>::T* __restrict__ g, typename Cx<R>::T* tile, int rs,
                                            const typename Cx<R>::T* __restrict__ tw, int tid, int nth,
                                            int jump = 0, int lstride = M, R post = R(1)) {  // lstride: words between lines
  using C = typename Cx<R>::T;
  constexpr int SUB = M / RAD;
  constexpr int W = SUB < 32 ? 32 / SUB : 1;   // lines per warp
  constexpr int BITS = ilog2(RAD);
  static_assert(L % 32 == 0 && (SUB >= 32 || 32 % SUB == 0), "zedge_stage: L must be a multiple of 32");
#ifndef LGM_ZEDGE_UNROLL
#define LGM_ZEDGE_UNROLL 1
#endif
  constexpr int kUnroll = INV ? 1 : LGM_ZEDGE_UNROLL;  // forward: items whose global loads are in flight together
#pragma unroll kUnroll
  for (int it = tid; it < L * SUB; it += nth) {
    const int rest = it % SUB, slot = it / SUB;
    const int l = (SUB < 32) ? (slot / 32) * 32 + (slot % W) * SUB + (slot % 32) / W : slot;
    C* gp = g + (size_t)l * lstride + rest;
    C* p = tile + rest * rs + l;
    auto jo = [&](int n) -> int { return JSH > 0 ? ((rest + n * SUB) >> JSH) * jump : 0; };
    C x[RAD];
    if (!INV) {
#pragma unroll
      for (int n = 0; n < RAD; ++n) x[n] = gp[n * SUB];
      reg_fft<RAD, false>(x);
#pragma unroll
      for (int i = 0; i < RAD; ++i) {
        const int k = bitrev(i, BITS);
        C v = x[i];
        if (k != 0) v = cmul(v, tw[rest * k]);
        p[k * SUB * rs + jo(k)] = v;
      }
    } else {
#pragma unroll
      for (int k = 0; k < RAD; ++k) {
        C v = p[k * SUB * rs + jo(k)];
        if (k != 0) v = cmulc(v, tw[rest * k]);
        x[k] = v;
      }
      reg_fft<RAD, true>(x);
#pragma unroll
      for (int i = 0; i < RAD; ++i) {
        C v = x[i];
        if (POST) {
          v.x = post_scale(post, v.x);
          v.y = post_scale(post, v.y);
        }
        gp[bitrev(i, BITS) * SUB] = v;
      }
    }
  }
}

// Real transforms with the outermost stage on global memory (see zedge_stage): `g` points to the L
// real lines (2M reals = M complex words each). Forward leaves the half spectrum in the tile rows
// 0..M; the inverse consumes it and writes the real lines. Requires >= 2 radix stages.
template <typename R, int M, int L>
__device__ __forceinline__ void real_fft_fwd_g(const R* g, typename Cx<R>::T* tile, int rs,
                                               const typename Cx<R>::T* twM, const typename Cx<R>::T* twz,
                                               int tid, int nth, int lstride = M) {
  using C = typename Cx<R>::T;
  constexpr int RAD0 = 1 << stage_bits(ilog2(M), 0);
  static_assert((ilog2(M) + 3) / 4 >= 2, "real_fft_fwd_g needs two radix stages");
  zedge_stage<R, M, RAD0, L, false>(reinterpret_cast<C*>(const_cast<R*>(g)), tile, rs, twM, tid, nth, 0, lstride);
  __syncthreads();
  if constexpr (!ColFFT<R, M, M / RAD0, 1, L>::LASTSTAGE) {
    ColFFT<R, M, M / RAD0, 1, L>::fwd_nolast(tile, rs, 1, twM, tid, nth);
    __syncthreads();
  }
  real_edge_stage<R, M, L, false>(tile, rs, 1, twz, tid, nth);
}
template <typename R, int M, int L, bool POST = false>
__device__ __forceinline__ void real_fft_inv_g(R* g, typename Cx<R>::T* tile, int rs,
                                               const typename Cx<R>::T* twM, const typename Cx<R>::T* twz,
                                               int tid, int nth, int lstride = M, R post = R(1)) {
  using C = typename Cx<R>::T;
  constexpr int RAD0 = 1 << stage_bits(ilog2(M), 0);
  real_edge_stage<R, M, L, true>(tile, rs, 1, twz, tid, nth);
  __syncthreads();
  if constexpr (!ColFFT<R, M, M / RAD0, 1, L>::LASTSTAGE) {
    ColFFT<R, M, M / RAD0, 1, L>::inv_nolast(tile, rs, 1, twM, tid, nth);
    __syncthreads();
  }
  zedge_stage<R, M, RAD0, L, true, 0, POST>(reinterpret_cast<C*>(g), tile, rs, twM, tid, nth, 0, lstride, post);
}

// real forward: tile rows 0..M-1 hold z[j] = (x[2j], x[2j+1]); on return rows 0..M hold the half
// spectrum in storage order (row M = Nyquist). twM: e^{-2 pi i j / M} (M entries).
template <typename R, int M, int L>
__device__ __forceinline__ void real_fft_fwd(typename Cx<R>::T* tile, int rs, int ls,
                                             const typename Cx<R>::T* twM, const typename Cx<R>::T* twz,
                                             int tid, int nth) {
  constexpr int ns = (ilog2(M) + 3) / 4;
  if constexpr (ns > 1) {
    ColFFT<R, M, M, 0, L>::fwd_nolast(tile, rs, ls, twM, tid, nth);
    __syncthreads();
  }
  real_edge_stage<R, M, L, false>(tile, rs, ls, twz, tid, nth);
}
// real inverse: rows 0..M hold the half spectrum; on return rows 0..M-1 hold the (unnormalised)
// time samples as (x[2j], x[2j+1]) pairs.
template <typename R, int M, int L>
__device__ __forceinline__ void real_fft_inv(typename Cx<R>::T* tile, int rs, int ls,
                                             const typename Cx<R>::T* twM, const typename Cx<R>::T* twz,
                                             int tid, int nth) {
  constexpr int ns = (ilog2(M) + 3) / 4;
  real_edge_stage<R, M, L, true>(tile, rs, ls, twz, tid, nth);
  if constexpr (ns > 1) {
    __syncthreads();
    ColFFT<R, M, M, 0, L>::inv_nolast(tile, rs, ls, twM, tid, nth);
  }
}

template <typename R, int N, int L>
__device__ __forceinline__ void col_fft_inv_to_global(typename Cx<R>::T* g, long long grs,
                                                      typename Cx<R>::T* tile, int rs, int ls,
                                                      const typename Cx<R>::T* tw, int lvalid,
                                                      int tid, int nth) {
  constexpr int RAD = 1 << stage_bits(ilog2(N), 0);
  if constexpr (N / RAD > 1) {
    ColFFT<R, N, N / RAD, 1, L>::inv(tile, rs, ls, tw, tid, nth);
    __syncthreads();
  }
  fft_stage_edge<R, N, RAD, L, true>(g, grs, tile, rs, ls, tw, lvalid, tid, nth);
}

}  // namespace lgm
