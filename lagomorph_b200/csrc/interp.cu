// interp.cu -- free-form interpolation (gather), its adjoint (splat) and regrid.
//
// Replaces the reference's cuda/interp.cu (K1-K4) and the regrid part of
// cuda/affine.cu (K17/K18). Not a port: the reference maps threads to (x,y) and
// loops serially over batch, channel and the contiguous z axis (uncoalesced);
// here one thread owns one voxel (consecutive lanes = consecutive addresses on
// the fastest axis), the batch is a grid dimension, and interpolation indices
// and weights are computed once per voxel and reused for every channel.
#include "common.cuh"

namespace lgm {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ forward
template <typename R, int D>
__global__ void __launch_bounds__(kThreads)
interp_fwd_kernel(R* __restrict__ out, const R* __restrict__ I, const R* __restrict__ u,
                  Geom<D> g, int C, long long I_batch_stride, double dt) {
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* un = u + n * D * g.V + vid;
  Axis<R> ax[D];
#pragma unroll
  for (int a = 0; a < D; ++a) ax[a] = axis_setup(coord<R>(pos[a], dt, un[a * g.V]), g.n[a]);
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

// ------------------------------------------------------------------ backward
// Splat weights exactly as the reference's alternating "d = 1 - d" sequence
// (include/interp.h:413-425, :437-453): w[0], w[1] are the weights used for the
// first visit of (lower, upper), w[2], w[3] for every later visit.
template <typename R>
__device__ __forceinline__ void flip_seq(R x, int xi, R (&w)[4]) {
  w[0] = R(1) - (x - (R)xi);
  w[1] = R(1) - w[0];
  w[2] = R(1) - w[1];
  w[3] = R(1) - w[2];
}

template <typename R, int D, bool NEED_I, bool NEED_U>
__global__ void __launch_bounds__(kThreads)
interp_bwd_kernel(R* __restrict__ d_I, R* __restrict__ d_u, const R* __restrict__ go,
                  const R* __restrict__ I, const R* __restrict__ u, Geom<D> g, int C,
                  long long I_batch_stride, double dt) {
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* un = u + n * D * g.V + vid;
  R h[D];
  Axis<R> ax[D];
  int fl[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    h[a] = coord<R>(pos[a], dt, un[a * g.V]);
    ax[a] = axis_setup(h[a], g.n[a]);
    fl[a] = (sizeof(R) == 4) ? __float2int_rd((float)h[a]) : __double2int_rd((double)h[a]);
  }
  const R* In = I + n * I_batch_stride;
  R* dIn = d_I + n * I_batch_stride;
  const R* gon = go + n * C * g.V + vid;
  R acc[D];
#pragma unroll
  for (int a = 0; a < D; ++a) acc[a] = R(0);

  R wx[4], wy[4], wz[4];
  if (NEED_I) {
    flip_seq<R>(h[0], fl[0], wx);
    flip_seq<R>(h[1], fl[1], wy);
    if constexpr (D == 3) flip_seq<R>(h[2], fl[2], wz);
  }

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
        // visit order x-major, then y, then z; y weights: first x uses wy[0],wy[1],
        // second x uses wy[2],wy[3]; z weights: first (x,y) uses wz[0],wz[1], later wz[2],wz[3]
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
    if (NEED_U) {
      R gd = (R)((double)diff * dt);  // "diff *= dt": cuda/interp.cu:169, :230
      if constexpr (D == 2) {
        // include/interp.h:202-203
        const R* r0 = In + c * g.V + ax[0].i0 * g.st[0];
        const R* r1 = In + c * g.V + ax[0].i1 * g.st[0];
        R v0 = __ldg(r0 + ax[1].i0), v3 = __ldg(r0 + ax[1].i1);
        R v1 = __ldg(r1 + ax[1].i0), v2 = __ldg(r1 + ax[1].i1);
        R gx = v1 - v0 + ax[1].t * (v2 - v3 - v1 + v0);
        R gy = v3 - v0 + ax[0].t * (v2 - v1 - v3 + v0);
        acc[0] = acc[0] + gx * gd;
        acc[1] = acc[1] + gy * gd;
      } else {
        Corners3<R> k = gather3<R>(In + c * g.V, ax[0], ax[1], ax[2], g.st[0], g.st[1]);
        R gx, gy, gz;
        lerp3_grad<R>(k, ax[0].t, ax[1].t, ax[2].t, gx, gy, gz);
        acc[0] = acc[0] + gx * gd;
        acc[1] = acc[1] + gy * gd;
        acc[2] = acc[2] + gz * gd;
      }
    }
  }
  if (NEED_U) {
    R* dun = d_u + n * D * g.V + vid;
#pragma unroll
    for (int a = 0; a < D; ++a) dun[a * g.V] = acc[a];
  }
}

// ------------------------------------------------------------------ regrid
template <typename R, int D, bool ADJOINT>
__global__ void __launch_bounds__(kThreads)
regrid_kernel(R* __restrict__ dst, const R* __restrict__ src, Geom<D> gi, Geom<D> go, int NC,
              R O0, R O1, R O2, R S0, R S1, R S2) {
  // forward : dst = out (NC, go), src = I (NC, gi)
  // adjoint : dst = d_I (NC, gi, pre-zeroed), src = gout (NC, go)
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= go.V) return;
  int pos[D];
  decode<D>(vid, go, pos);
  const R O[3] = {O0, O1, O2}, S[3] = {S0, S1, S2};
  R h[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    R o = (R)(.5 * (double)(R)(go.n[a] - 1));  // cuda/affine.cu:627-628, :661-663
    if (D == 3 && a == 2 && !ADJOINT) {
      // the reference's 3-D forward accumulates hz += Sz along k (affine.cu:669-675);
      // reproduce the rounding of that running sum.
      R hz = O[a] - o * S[a];
      for (int k = 0; k < pos[a]; ++k) hz += S[a];
      h[a] = hz;
    } else {
      h[a] = ((R)pos[a] - o) * S[a] + O[a];  // affine.cu:629-630, :792
    }
  }
  Axis<R> ax[D];
  int fl[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    ax[a] = axis_setup(h[a], gi.n[a]);
    fl[a] = (sizeof(R) == 4) ? __float2int_rd((float)h[a]) : __double2int_rd((double)h[a]);
  }
  if (!ADJOINT) {
    for (int c = 0; c < NC; ++c) {
      const R* In = src + c * gi.V;
      if constexpr (D == 2) {
        dst[c * go.V + vid] = lerp2<R>(In, ax[0], ax[1], gi.st[0]);
      } else {
        Corners3<R> k = gather3<R>(In, ax[0], ax[1], ax[2], gi.st[0], gi.st[1]);
        dst[c * go.V + vid] = lerp3_eval<R>(k, ax[0].t, ax[1].t, ax[2].t);
      }
    }
  } else {
    R wx[4], wy[4], wz[4];
    flip_seq<R>(h[0], fl[0], wx);
    flip_seq<R>(h[1], fl[1], wy);
    if constexpr (D == 3) flip_seq<R>(h[2], fl[2], wz);
    for (int c = 0; c < NC; ++c) {
      R diff = src[c * go.V + vid];
      R* dI = dst + c * gi.V;
      if constexpr (D == 2) {
        const long long r0 = ax[0].i0 * gi.st[0], r1 = ax[0].i1 * gi.st[0];
        red_add(dI + r0 + ax[1].i0, (wx[0] * wy[0]) * diff);
        red_add(dI + r0 + ax[1].i1, (wx[0] * wy[1]) * diff);
        red_add(dI + r1 + ax[1].i0, (wx[1] * wy[2]) * diff);
        red_add(dI + r1 + ax[1].i1, (wx[1] * wy[3]) * diff);
      } else {
        R* p00 = dI + ax[0].i0 * gi.st[0] + ax[1].i0 * gi.st[1];
        R* p01 = dI + ax[0].i0 * gi.st[0] + ax[1].i1 * gi.st[1];
        R* p10 = dI + ax[0].i1 * gi.st[0] + ax[1].i0 * gi.st[1];
        R* p11 = dI + ax[0].i1 * gi.st[0] + ax[1].i1 * gi.st[1];
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
  }
}

// ------------------------------------------------------------------ host side
template <typename R, int D>
static int interp_fwd_t(void* out, const void* I, const void* u, int64_t N, int64_t NI, int64_t C,
                        const int64_t* shape, double dt, cudaStream_t s) {
  Geom<D> g = make_geom<D>(shape);
  if (g.V == 0 || N == 0 || C == 0) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  const long long ibs = (NI < N) ? 0 : C * g.V;
  interp_fwd_kernel<R, D><<<grid, kThreads, 0, s>>>((R*)out, (const R*)I, (const R*)u, g, (int)C, ibs, dt);
  count_launch("interp_fwd", s);
  return finish(s, "lgm_interp_fwd");
}

template <typename R, int D>
static int interp_bwd_t(void* d_I, void* d_u, const void* go, const void* I, const void* u,
                        int64_t N, int64_t NI, int64_t C, const int64_t* shape, double dt,
                        cudaStream_t s) {
  Geom<D> g = make_geom<D>(shape);
  if (d_I) {
    cudaError_t e = cudaMemsetAsync(d_I, 0, (size_t)(NI * C * g.V) * sizeof(R), s);
    if (e != cudaSuccess) return set_error((int)e, "lgm_interp_bwd: memset: %s", cudaGetErrorString(e));
  }
  if (g.V == 0 || N == 0 || C == 0 || (!d_I && !d_u)) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  const long long ibs = (NI < N) ? 0 : C * g.V;
#define LAUNCH(NI_, NU_)                                                                         \
  interp_bwd_kernel<R, D, NI_, NU_><<<grid, kThreads, 0, s>>>((R*)d_I, (R*)d_u, (const R*)go,   \
                                                              (const R*)I, (const R*)u, g, (int)C, ibs, dt)
  if (d_I && d_u) LAUNCH(true, true);
  else if (d_I) LAUNCH(true, false);
  else LAUNCH(false, true);
#undef LAUNCH
  count_launch("interp_bwd", s);
  return finish(s, "lgm_interp_bwd");
}

template <typename R, int D, bool ADJ>
static int regrid_t(void* dst, const void* src, int64_t N, int64_t C, const int64_t* shape,
                    const int64_t* oshape, const double* origin, const double* spacing,
                    cudaStream_t s) {
  Geom<D> gi = make_geom<D>(shape), go = make_geom<D>(oshape);
  if (ADJ) {
    cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)(N * C * gi.V) * sizeof(R), s);
    if (e != cudaSuccess) return set_error((int)e, "lgm_regrid_bwd: memset: %s", cudaGetErrorString(e));
  }
  if (gi.V == 0 || go.V == 0 || N * C == 0) return LGM_OK;
  dim3 grid((unsigned)cdiv(go.V, kThreads));
  R O2 = D == 3 ? (R)origin[2] : R(0), S2 = D == 3 ? (R)spacing[2] : R(0);
  regrid_kernel<R, D, ADJ><<<grid, kThreads, 0, s>>>((R*)dst, (const R*)src, gi, go, (int)(N * C),
                                                     (R)origin[0], (R)origin[1], O2, (R)spacing[0],
                                                     (R)spacing[1], S2);
  count_launch("regrid", s);
  return finish(s, ADJ ? "lgm_regrid_bwd" : "lgm_regrid_fwd");
}

}  // namespace lgm

namespace lgm {  // gather3.cu; LGM_EUNSUP = fast path not applicable
int interp3_f32(void* out, const void* I, const void* u, int64_t N, int64_t NI, int64_t C, const int64_t* sh,
                double dt, cudaStream_t s);
int splat3_f32(void* d_I, const void* go, const void* u, int64_t N, int64_t NI, int64_t C, const int64_t* sh,
               double dt, cudaStream_t s);
int interp_du3_f32(void* d_u, const void* go, const void* I, const void* u, int64_t N, int64_t NI, int64_t C,
                   const int64_t* sh, double dt, cudaStream_t s);
}  // namespace lgm

using namespace lgm;

#define DISPATCH_RD(dtype, dim, FN, ...)                                 \
  do {                                                                   \
    if ((dtype) == LGM_F32 && (dim) == 2) return FN<float, 2>(__VA_ARGS__);  \
    if ((dtype) == LGM_F32 && (dim) == 3) return FN<float, 3>(__VA_ARGS__);  \
    if ((dtype) == LGM_F64 && (dim) == 2) return FN<double, 2>(__VA_ARGS__); \
    if ((dtype) == LGM_F64 && (dim) == 3) return FN<double, 3>(__VA_ARGS__); \
    return set_error(LGM_EINVAL, "unsupported dtype %d / dim %d", (dtype), (dim)); \
  } while (0)

static bool shape_ok(int dim, const int64_t* shape) {
  if (dim != 2 && dim != 3) return false;
  long long v = 1;
  for (int a = 0; a < dim; ++a) {
    if (shape[a] < 0 || shape[a] > 0x7fffffff) return false;
    v *= shape[a];
    if (v > 0x7fffffffLL) return false;  // 32-bit in-volume indexing
  }
  return v >= 0;
}

extern "C" int lgm_interp_fwd(int dtype, void* out, const void* I, const void* u, int64_t N,
                              int64_t NI, int64_t C, int dim, const int64_t* shape, double dt,
                              void* stream) {
  LGM_REQUIRE(shape_ok(dim, shape), "lgm_interp_fwd: Only two- and three-dimensional interpolation is supported");
  LGM_REQUIRE(N >= 0 && C >= 0 && NI >= 0 && N <= 65535, "lgm_interp_fwd: bad batch/channel count");
  LGM_REQUIRE(NI == N || (NI == 1 && N >= 1), "lgm_interp_fwd: image batch must equal the displacement batch or be 1");
  if (dtype == LGM_F32 && dim == 3 && N > 0 && C > 0) {  // fp32 3-D fast path (gather3.cu)
    int rc = interp3_f32(out, I, u, N, NI, C, shape, dt, (cudaStream_t)stream);
    if (rc != LGM_EUNSUP) return rc;
  }
  DISPATCH_RD(dtype, dim, interp_fwd_t, out, I, u, N, NI, C, shape, dt, (cudaStream_t)stream);
}

extern "C" int lgm_interp_bwd(int dtype, void* d_I, void* d_u, const void* gout, const void* I,
                              const void* u, int64_t N, int64_t NI, int64_t C, int dim,
                              const int64_t* shape, double dt, void* stream) {
  LGM_REQUIRE(shape_ok(dim, shape), "lgm_interp_bwd: Only two- and three-dimensional interpolation is supported");
  LGM_REQUIRE(N >= 0 && C >= 0 && NI >= 0 && N <= 65535, "lgm_interp_bwd: bad batch/channel count");
  LGM_REQUIRE(NI == N || (NI == 1 && N >= 1), "lgm_interp_bwd: image batch must equal the displacement batch or be 1");
  if (dtype == LGM_F32 && dim == 3 && N > 0 && C > 0) {
    // fp32 3-D fast paths (gather3.cu): warp-aggregated splat for d_I, gather kernel for d_u;
    // whatever they decline falls through to the generic kernel
    cudaStream_t s = (cudaStream_t)stream;
    bool done_I = (d_I == nullptr), done_u = (d_u == nullptr);
    if (d_I) {
      long long V = shape[0] * shape[1] * shape[2];
      cudaError_t e = cudaMemsetAsync(d_I, 0, (size_t)(NI * C * V) * sizeof(float), s);
      if (e != cudaSuccess) return set_error((int)e, "lgm_interp_bwd: memset: %s", cudaGetErrorString(e));
      int rc = splat3_f32(d_I, gout, u, N, NI, C, shape, dt, s);
      if (rc == LGM_OK) done_I = true;
      else if (rc != LGM_EUNSUP) return rc;
    }
    if (d_u) {
      int rc = interp_du3_f32(d_u, gout, I, u, N, NI, C, shape, dt, s);
      if (rc == LGM_OK) done_u = true;
      else if (rc != LGM_EUNSUP) return rc;
    }
    if (done_I && done_u) return LGM_OK;
    return interp_bwd_t<float, 3>(done_I ? nullptr : d_I, done_u ? nullptr : d_u, gout, I, u, N, NI, C, shape, dt, s);
  }
  DISPATCH_RD(dtype, dim, interp_bwd_t, d_I, d_u, gout, I, u, N, NI, C, shape, dt, (cudaStream_t)stream);
}

template <typename R, int D>
static int regrid_fwd_t(void* a, const void* b, int64_t N, int64_t C, const int64_t* sh,
                        const int64_t* osh, const double* o, const double* sp, cudaStream_t s) {
  return regrid_t<R, D, false>(a, b, N, C, sh, osh, o, sp, s);
}
template <typename R, int D>
static int regrid_bwd_t(void* a, const void* b, int64_t N, int64_t C, const int64_t* sh,
                        const int64_t* osh, const double* o, const double* sp, cudaStream_t s) {
  return regrid_t<R, D, true>(a, b, N, C, sh, osh, o, sp, s);
}

extern "C" int lgm_regrid_fwd(int dtype, void* out, const void* I, int64_t N, int64_t C, int dim,
                              const int64_t* shape, const int64_t* out_shape, const double* origin,
                              const double* spacing, void* stream) {
  LGM_REQUIRE(shape_ok(dim, shape) && shape_ok(dim, out_shape), "lgm_regrid_fwd: Only two- and three-dimensional regridding is supported");
  DISPATCH_RD(dtype, dim, regrid_fwd_t, out, I, N, C, shape, out_shape, origin, spacing, (cudaStream_t)stream);
}
extern "C" int lgm_regrid_bwd(int dtype, void* d_I, const void* gout, int64_t N, int64_t C, int dim,
                              const int64_t* shape, const int64_t* out_shape, const double* origin,
                              const double* spacing, void* stream) {
  LGM_REQUIRE(shape_ok(dim, shape) && shape_ok(dim, out_shape), "lgm_regrid_bwd: Only two- and three-dimensional regridding is supported");
  DISPATCH_RD(dtype, dim, regrid_bwd_t, d_I, gout, N, C, shape, out_shape, origin, spacing, (cudaStream_t)stream);
}
