// affine.cu -- affine_interp forward / backward (BASELINE config 4).
//
// Replaces the reference's cuda/affine.cu K19/K20. The reference's backward runs
// one 512-thread CTA per (subject, channel) that walks the whole volume
// (grid (1,N,C), cuda/affine.cu:556-557); here the volume is spread over the
// grid, each thread owns one voxel (all channels), partial d_A/d_T are reduced
// with warp shuffles + one shared-memory step and flushed with one atomic per
// CTA and matrix entry.
#include "common.cuh"

namespace lgm {

constexpr int kThreads = 256;

template <typename R, int D>
__device__ __forceinline__ void affine_coords(const R* __restrict__ An, const R* __restrict__ Tn,
                                              const int (&pos)[D], const Geom<D>& g, R (&f)[D],
                                              R (&h)[D]) {
  R o[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    o[a] = (R)(.5 * (double)(R)(g.n[a] - 1));  // cuda/affine.cu:42-43, :81-83
    f[a] = (R)pos[a] - o[a];
  }
#pragma unroll
  for (int r = 0; r < D; ++r) {
    // hx = An[0]*fi + An[1]*fj (+ An[2]*fk) + Tn[0] + ox   (affine.cu:51-52, :98-100)
    R s = An[r * D] * f[0] + An[r * D + 1] * f[1];
    if constexpr (D == 3) s = s + An[r * D + 2] * f[2];
    h[r] = s + Tn[r] + o[r];
  }
}

template <typename R, int D>
__global__ void __launch_bounds__(kThreads)
affine_fwd_kernel(R* __restrict__ out, const R* __restrict__ I, const R* __restrict__ A,
                  const R* __restrict__ T, Geom<D> g, int C, long long I_batch_stride) {
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  R f[D], h[D];
  affine_coords<R, D>(A + n * D * D, T + n * D, pos, g, f, h);
  Axis<R> ax[D];
#pragma unroll
  for (int a = 0; a < D; ++a) ax[a] = axis_setup(h[a], g.n[a]);
  const R* In = I + n * I_batch_stride;
  R* on = out + n * C * g.V + vid;
  for (int c = 0; c < C; ++c) {
    if constexpr (D == 2) {
      on[c * g.V] = lerp2<R>(In + c * g.V, ax[0], ax[1], g.st[0]);
    } else {
      Corners3<R> k = gather3<R>(In + c * g.V, ax[0], ax[1], ax[2], g.st[0], g.st[1]);
      on[c * g.V] = lerp3_eval<R>(k, ax[0].t, ax[1].t, ax[2].t);
    }
  }
}

template <typename R>
__device__ __forceinline__ void flip_seq(R x, int xi, R (&w)[4]) {
  w[0] = R(1) - (x - (R)xi);
  w[1] = R(1) - w[0];
  w[2] = R(1) - w[1];
  w[3] = R(1) - w[2];
}

template <typename R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <typename R, int D, bool NEED_I, bool NEED_AT>
__global__ void __launch_bounds__(kThreads)
affine_bwd_kernel(R* __restrict__ d_I, R* __restrict__ d_A, R* __restrict__ d_T,
                  const R* __restrict__ go, const R* __restrict__ I, const R* __restrict__ A,
                  const R* __restrict__ T, Geom<D> g, int C, long long I_batch_stride) {
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  const long long n = blockIdx.y;
  constexpr int NP = D * D + D;
  R part[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) part[q] = R(0);
  if (vid < g.V) {
    int pos[D];
    decode<D>(vid, g, pos);
    R f[D], h[D];
    affine_coords<R, D>(A + n * D * D, T + n * D, pos, g, f, h);
    Axis<R> ax[D];
    int fl[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      ax[a] = axis_setup(h[a], g.n[a]);
      fl[a] = (sizeof(R) == 4) ? __float2int_rd((float)h[a]) : __double2int_rd((double)h[a]);
    }
    R wx[4], wy[4], wz[4];
    if (NEED_I) {
      flip_seq<R>(h[0], fl[0], wx);
      flip_seq<R>(h[1], fl[1], wy);
      if constexpr (D == 3) flip_seq<R>(h[2], fl[2], wz);
    }
    const R* In = I + n * I_batch_stride;
    R* dIn = d_I + n * I_batch_stride;
    const R* gon = go + n * C * g.V + vid;
    for (int c = 0; c < C; ++c) {
      R diff = gon[c * g.V];
      if (NEED_I) {
        R* dI = dIn + c * g.V;
        if constexpr (D == 2) {
          const long long r0 = ax[0].i0 * g.st[0], r1 = ax[0].i1 * g.st[0];
          red_add(dI + r0 + ax[1].i0, (wx[0] * wy[0]) * diff);
          red_add(dI + r0 + ax[1].i1, (wx[0] * wy[1]) * diff);
          red_add(dI + r1 + ax[1].i0, (wx[1] * wy[2]) * diff);
          red_add(dI + r1 + ax[1].i1, (wx[1] * wy[3]) * diff);
        } else {
          R* p00 = dI + ax[0].i0 * g.st[0] + ax[1].i0 * g.st[1];
          R* p01 = dI + ax[0].i0 * g.st[0] + ax[1].i1 * g.st[1];
          R* p10 = dI + ax[0].i1 * g.st[0] + ax[1].i0 * g.st[1];
          R* p11 = dI + ax[0].i1 * g.st[0] + ax[1].i1 * g.st[1];
          red_add(p00 + ax[2].i0, (wx[0] * wy[0] * wz[0]) * diff);
          red_add(p00 + ax[2].i1, (wx[0] * wy[0] * wz[1]) * diff);
          red_add(p01 + ax[2].i0, (wx[0] * wy[1] * wz[2]) * diff);
          red_add(p01 + ax[2].i1, (wx[0] * wy[1] * wz[3]) * diff);
          red_add(p10 + ax[2].i0, (wx[1] * wy[2] * wz[2]) * diff);
          red_add(p10 + ax[2].i1, (wx[1] * wy[2] * wz[3]) * diff);
          red_add(p11 + ax[2].i0, (wx[1] * wy[3] * wz[2]) * diff);
          red_add(p11 + ax[2].i1, (wx[1] * wy[3] * wz[3]) * diff);
        }
      }
      if (NEED_AT) {
        R gr[D];
        if constexpr (D == 2) {
          const R* r0 = In + c * g.V + ax[0].i0 * g.st[0];
          const R* r1 = In + c * g.V + ax[0].i1 * g.st[0];
          R v0 = __ldg(r0 + ax[1].i0), v3 = __ldg(r0 + ax[1].i1);
          R v1 = __ldg(r1 + ax[1].i0), v2 = __ldg(r1 + ax[1].i1);
          gr[0] = v1 - v0 + ax[1].t * (v2 - v3 - v1 + v0);
          gr[1] = v3 - v0 + ax[0].t * (v2 - v1 - v3 + v0);
        } else {
          Corners3<R> k = gather3<R>(In + c * g.V, ax[0], ax[1], ax[2], g.st[0], g.st[1]);
          lerp3_grad<R>(k, ax[0].t, ax[1].t, ax[2].t, gr[0], gr[1], gr[2]);
        }
#pragma unroll
        for (int r = 0; r < D; ++r) {
          R gd = gr[r] * diff;  // "gx *= diff": affine.cu:273-274, :421-423
#pragma unroll
          for (int s = 0; s < D; ++s) part[r * D + s] += gd * f[s];
          part[D * D + r] += gd;
        }
      }
    }
  }
  if (NEED_AT) {
    __shared__ R red[kThreads / 32][NP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      R v = warp_sum<R>(part[q]);
      if (lane == 0) red[wid][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < NP) {
      R v = R(0);
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) v += red[w][threadIdx.x];
      if (threadIdx.x < D * D) {
        if (d_A) red_add(d_A + n * D * D + threadIdx.x, v);
      } else {
        if (d_T) red_add(d_T + n * D + (threadIdx.x - D * D), v);
      }
    }
  }
}

// fp32 3-D fast paths (affine3.cu); LGM_EUNSUP = not applicable
int affine3_fwd_f32(void* out, const void* I, const void* A, const void* T, int64_t N, int64_t NI, int64_t C,
                    const int64_t* sh, cudaStream_t s);
int affine3_bwd_f32(void* d_I, void* d_A, void* d_T, const void* go, const void* I, const void* A, const void* T,
                    int64_t N, int64_t NI, int64_t C, const int64_t* sh, cudaStream_t s);

template <typename R, int D>
static int affine_fwd_t(void* out, const void* I, const void* A, const void* T, int64_t N,
                        int64_t NI, int64_t C, const int64_t* shape, cudaStream_t s) {
  Geom<D> g = make_geom<D>(shape);
  if (g.V == 0 || N == 0 || C == 0) return LGM_OK;
  if constexpr (sizeof(R) == 4 && D == 3) {
    int rc = affine3_fwd_f32(out, I, A, T, N, NI, C, shape, s);
    if (rc != LGM_EUNSUP) return rc;
  }
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  const long long ibs = (NI == 1 && N > 1) ? 0 : C * g.V;
  affine_fwd_kernel<R, D><<<grid, kThreads, 0, s>>>((R*)out, (const R*)I, (const R*)A, (const R*)T, g, (int)C, ibs);
  count_launch("affine_fwd", s);
  return finish(s, "lgm_affine_interp_fwd");
}

template <typename R, int D>
static int affine_bwd_t(void* d_I, void* d_A, void* d_T, const void* go, const void* I,
                        const void* A, const void* T, int64_t N, int64_t NI, int64_t C,
                        const int64_t* shape, cudaStream_t s) {
  Geom<D> g = make_geom<D>(shape);
  cudaError_t e = cudaSuccess;
  if (d_I) e = cudaMemsetAsync(d_I, 0, (size_t)(NI * C * g.V) * sizeof(R), s);
  if (e == cudaSuccess && d_A) e = cudaMemsetAsync(d_A, 0, (size_t)(N * D * D) * sizeof(R), s);
  if (e == cudaSuccess && d_T) e = cudaMemsetAsync(d_T, 0, (size_t)(N * D) * sizeof(R), s);
  if (e != cudaSuccess) return set_error((int)e, "lgm_affine_interp_bwd: memset: %s", cudaGetErrorString(e));
  if (g.V == 0 || N == 0 || C == 0 || (!d_I && !d_A && !d_T)) return LGM_OK;
  if constexpr (sizeof(R) == 4 && D == 3) {
    int rc = affine3_bwd_f32(d_I, d_A, d_T, go, I, A, T, N, NI, C, shape, s);
    if (rc != LGM_EUNSUP) return rc;
  }
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  const long long ibs = (NI == 1 && N > 1) ? 0 : C * g.V;
  const bool need_at = d_A || d_T;
#define L(NI_, NAT_) affine_bwd_kernel<R, D, NI_, NAT_><<<grid, kThreads, 0, s>>>((R*)d_I, (R*)d_A, (R*)d_T, (const R*)go, (const R*)I, (const R*)A, (const R*)T, g, (int)C, ibs)
  if (d_I && need_at) L(true, true);
  else if (d_I) L(true, false);
  else L(false, true);
#undef L
  count_launch("affine_bwd", s);
  return finish(s, "lgm_affine_interp_bwd");
}

}  // namespace lgm

using namespace lgm;

#define DISPATCH_RD(dtype, dim, FN, ...)                                           \
  do {                                                                             \
    if ((dtype) == LGM_F32 && (dim) == 2) return FN<float, 2>(__VA_ARGS__);        \
    if ((dtype) == LGM_F32 && (dim) == 3) return FN<float, 3>(__VA_ARGS__);        \
    if ((dtype) == LGM_F64 && (dim) == 2) return FN<double, 2>(__VA_ARGS__);       \
    if ((dtype) == LGM_F64 && (dim) == 3) return FN<double, 3>(__VA_ARGS__);       \
    return set_error(LGM_EINVAL, "unsupported dtype %d / dim %d", (dtype), (dim)); \
  } while (0)

extern "C" int lgm_affine_interp_fwd(int dtype, void* out, const void* I, const void* A,
                                     const void* T, int64_t N, int64_t NI, int64_t C, int dim,
                                     const int64_t* shape, void* stream) {
  LGM_REQUIRE(dim == 2 || dim == 3, "Only two- and three-dimensional affine interpolation is supported");
  LGM_REQUIRE(N >= 0 && N <= 65535 && (NI == N || NI == 1), "lgm_affine_interp_fwd: bad batch sizes");
  LGM_REQUIRE(dim == 2 ? geom_fits<2>(shape) : geom_fits<3>(shape), "lgm_affine_interp_fwd: volume too large");
  DISPATCH_RD(dtype, dim, affine_fwd_t, out, I, A, T, N, NI, C, shape, (cudaStream_t)stream);
}
extern "C" int lgm_affine_interp_bwd(int dtype, void* d_I, void* d_A, void* d_T, const void* gout,
                                     const void* I, const void* A, const void* T, int64_t N,
                                     int64_t NI, int64_t C, int dim, const int64_t* shape,
                                     void* stream) {
  LGM_REQUIRE(dim == 2 || dim == 3, "Only two- and three-dimensional affine interpolation is supported");
  LGM_REQUIRE(N >= 0 && N <= 65535 && (NI == N || NI == 1), "lgm_affine_interp_bwd: bad batch sizes");
  LGM_REQUIRE(dim == 2 ? geom_fits<2>(shape) : geom_fits<3>(shape), "lgm_affine_interp_bwd: volume too large");
  DISPATCH_RD(dtype, dim, affine_bwd_t, d_I, d_A, d_T, gout, I, A, T, N, NI, C, shape, (cudaStream_t)stream);
}
