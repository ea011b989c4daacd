// diff.cu -- clamped central-difference Jacobian operators and the fused
// adjoint-representation kernels built from them.
//
// Replaces the reference's cuda/diff.cu (K6-K13) and fuses the Python-level
// chains of lagomorph/adjrep.py (ad, ad_star, Ad_star) into single kernels.
// One thread owns one voxel; lanes run along the fastest axis so every stencil
// leg is a coalesced, L1-served load; the batch is a grid dimension.
#include "common.cuh"

namespace lgm {

constexpr int kThreads = 256;

// D_a f at the centre voxel: 0.5*(f[+1] - f[-1]) with clamped indices, i.e. the
// boundary value is half the one-sided difference (include/diff.h:6-52 with
// include/extrap.h:121-125).
template <typename R>
__device__ __forceinline__ R cdiff(const R* __restrict__ f, int pos, int n, long long st) {
  R hi = __ldg(f + (pos < n - 1 ? st : 0));
  R lo = __ldg(f - (pos > 0 ? st : 0));
  return R(0.5) * (hi - lo);
}

template <typename R, int D>
__device__ __forceinline__ void grad(const R* __restrict__ f, const int (&pos)[D], const Geom<D>& g,
                                     R (&out)[D]) {
#pragma unroll
  for (int a = 0; a < D; ++a) out[a] = cdiff<R>(f, pos[a], g.n[a], g.st[a]);
}

// (D_a^T (p*q)) at the centre voxel: exact transpose of cdiff including its
// boundary rows (cuda/diff.cu:432-460). p, q point at the centre voxel.
template <typename R>
__device__ __forceinline__ R cdiffT(const R* __restrict__ p, const R* __restrict__ q, int pos, int n,
                                    long long st) {
  if (pos == 0) return R(-0.5) * (__ldg(p) * __ldg(q) + __ldg(p + st) * __ldg(q + st));
  if (pos == n - 1) return R(0.5) * (__ldg(p) * __ldg(q) + __ldg(p - st) * __ldg(q - st));
  return R(-0.5) * (__ldg(p + st) * __ldg(q + st) - __ldg(p - st) * __ldg(q - st));
}

// ---------------------------------------------------------------- jtvf forward
// CC: compile-time channel count (0 = use the runtime C); the common C == D case is unrolled so the
// loads of w and of the stencil legs are hoisted and batched.
template <typename R, int D, bool DISP, bool TRANS, int CC = 0>
__global__ void __launch_bounds__(kThreads)
jtvf_fwd_kernel(R* __restrict__ out, const R* __restrict__ v, const R* __restrict__ w, Geom<D> g,
                int C_rt) {
  const int C = CC ? CC : C_rt;
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* vn = v + n * C * g.V + vid;
  const R* wn = w + n * D * g.V + vid;
  R* on = out + n * C * g.V + vid;
  R wv[D];
#pragma unroll
  for (int a = 0; a < D; ++a) wv[a] = wn[a * g.V];
  if (TRANS) {  // cuda/diff.cu:34-44, :81-102
    R acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      R gr[D];
      grad<R, D>(vn + c * g.V, pos, g, gr);
      if (DISP) gr[c] += R(1);
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = (c == 0) ? gr[d] * wv[c] : acc[d] + gr[d] * wv[c];
    }
#pragma unroll
    for (int d = 0; d < D; ++d) on[d * g.V] = acc[d];
  } else {  // cuda/diff.cu:46-56, :104-122
#pragma unroll
    for (int c = 0; c < C; ++c) {
      R gr[D];
      grad<R, D>(vn + c * g.V, pos, g, gr);
      if (DISP) {
#pragma unroll
        for (int d = 0; d < D; ++d)
          if (c == d) gr[d] += R(1);
      }
      R s = gr[0] * wv[0] + gr[1] * wv[1];
      if constexpr (D == 3) s = s + gr[2] * wv[2];
      on[c * g.V] = s;
    }
  }
}

// ---------------------------------------------------------------- jtvf backward
template <typename R, int D, bool DISP, bool TRANS, bool NEED_V, bool NEED_W, int CC = 0>
__global__ void __launch_bounds__(kThreads)
jtvf_bwd_kernel(R* __restrict__ d_v, R* __restrict__ d_w, const R* __restrict__ go,
                const R* __restrict__ v, const R* __restrict__ w, Geom<D> g, int C_rt) {
  const int C = CC ? CC : C_rt;
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* vn = v + n * C * g.V + vid;
  const R* wn = w + n * D * g.V + vid;
  const R* gon = go + n * C * g.V + vid;
  if (TRANS) {
    if (NEED_W) {  // d_w[c] = sum_d (D_d v_c + delta_cd) gout_d   (diff.cu:311-335)
      R gv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) gv[d] = gon[d * g.V];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        R gr[D];
        grad<R, D>(vn + c * g.V, pos, g, gr);
        if (DISP) gr[c] += R(1);
        R s = gr[0] * gv[0] + gr[1] * gv[1];
        if constexpr (D == 3) s = s + gr[2] * gv[2];
        d_w[(n * D + c) * g.V + vid] = s;
      }
    }
    if (NEED_V) {  // d_v[c] = sum_d D_d^T (w_c gout_d)   (diff.cu:336-407)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        R acc = R(0);
#pragma unroll
        for (int d = 0; d < D; ++d)
          acc += cdiffT<R>(wn + c * g.V, gon + d * g.V, pos[d], g.n[d], g.st[d]);
        d_v[(n * C + c) * g.V + vid] = acc;
      }
    }
  } else {
    R dw[D];
#pragma unroll
    for (int d = 0; d < D; ++d) dw[d] = R(0);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (NEED_W) {  // d_w[d] += (D_d v_c + delta) gout_c   (diff.cu:417-431)
        R gr[D];
        grad<R, D>(vn + c * g.V, pos, g, gr);
        if (DISP) {
#pragma unroll
          for (int d = 0; d < D; ++d)
            if (c == d) gr[d] += R(1);
        }
        R gc = gon[c * g.V];
#pragma unroll
        for (int d = 0; d < D; ++d) dw[d] += gr[d] * gc;
      }
      if (NEED_V) {  // d_v[c] = sum_d D_d^T (w_d gout_c)   (diff.cu:432-460)
        R acc = R(0);
#pragma unroll
        for (int d = 0; d < D; ++d)
          acc += cdiffT<R>(wn + d * g.V, gon + c * g.V, pos[d], g.n[d], g.st[d]);
        d_v[(n * C + c) * g.V + vid] = acc;
      }
    }
    if (NEED_W) {
#pragma unroll
      for (int d = 0; d < D; ++d) d_w[(n * D + d) * g.V + vid] = dw[d];
    }
  }
}

// ---------------------------------------------------------------- adjoint fwd / bwd
template <typename R, int D, int CC = 0>
__global__ void __launch_bounds__(kThreads)
jtvf_adj_fwd_kernel(R* __restrict__ out, const R* __restrict__ z, const R* __restrict__ w,
                    Geom<D> g, int C_rt) {
  const int C = CC ? CC : C_rt;
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* wn = w + n * D * g.V + vid;
#pragma unroll
  for (int c = 0; c < C; ++c) {  // out_c = sum_d D_d^T (w_d z_c)   (diff.cu:546-632)
    const R* zc = z + (n * C + c) * g.V + vid;
    R acc = R(0);
#pragma unroll
    for (int d = 0; d < D; ++d) acc += cdiffT<R>(wn + d * g.V, zc, pos[d], g.n[d], g.st[d]);
    out[(n * C + c) * g.V + vid] = acc;
  }
}

template <typename R, int D, bool NEED_Z, bool NEED_W>
__global__ void __launch_bounds__(kThreads)
jtvf_adj_bwd_kernel(R* __restrict__ d_z, R* __restrict__ d_w, const R* __restrict__ go,
                    const R* __restrict__ z, const R* __restrict__ w, Geom<D> g) {
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* zn = z + n * D * g.V + vid;
  const R* wn = w + n * D * g.V + vid;
  const R* gon = go + n * D * g.V + vid;
  R wv[D], dw[D];
#pragma unroll
  for (int d = 0; d < D; ++d) wv[d] = wn[d * g.V];
#pragma unroll
  for (int c = 0; c < D; ++c) {  // cuda/diff.cu:722-780
    R gr[D];
    grad<R, D>(gon + c * g.V, pos, g, gr);
    R zc = zn[c * g.V];
#pragma unroll
    for (int d = 0; d < D; ++d) dw[d] = (c == 0) ? gr[d] * zc : dw[d] + gr[d] * zc;
    if (NEED_Z) {
      R s = gr[0] * wv[0] + gr[1] * wv[1];
      if constexpr (D == 3) s = s + gr[2] * wv[2];
      d_z[(n * D + c) * g.V + vid] = s;
    }
  }
  if (NEED_W) {
#pragma unroll
    for (int d = 0; d < D; ++d) d_w[(n * D + d) * g.V + vid] = dw[d];
  }
}

// ---------------------------------------------------------------- fused adjoint representation
// ad_star(v,m)_c = sum_d (D_c v_d) m_d - sum_d D_d^T (v_d m_c)      (adjrep.py:69-83)
// ad(v,w)_c      = sum_d (D_d v_c) w_d - sum_d (D_d w_c) v_d        (adjrep.py:37-47)
template <typename R, int D, bool STAR>
__global__ void __launch_bounds__(kThreads)
ad_kernel(R* __restrict__ out, const R* __restrict__ v, const R* __restrict__ m, Geom<D> g) {
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* vn = v + n * D * g.V + vid;
  const R* mn = m + n * D * g.V + vid;
  R* on = out + n * D * g.V + vid;
  R vv[D], mv[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    vv[a] = vn[a * g.V];
    mv[a] = mn[a * g.V];
  }
  if (STAR) {
    R A[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {  // transposed jtvf, same accumulation order as jtvf_fwd_kernel
      R gr[D];
      grad<R, D>(vn + c * g.V, pos, g, gr);
#pragma unroll
      for (int d = 0; d < D; ++d) A[d] = (c == 0) ? gr[d] * mv[c] : A[d] + gr[d] * mv[c];
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
      R B = R(0);
#pragma unroll
      for (int d = 0; d < D; ++d) B += cdiffT<R>(vn + d * g.V, mn + c * g.V, pos[d], g.n[d], g.st[d]);
      on[c * g.V] = A[c] - B;
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      R gv[D], gw[D];
      grad<R, D>(vn + c * g.V, pos, g, gv);
      grad<R, D>(mn + c * g.V, pos, g, gw);
      R a = gv[0] * mv[0] + gv[1] * mv[1];
      R b = gw[0] * vv[0] + gw[1] * vv[1];
      if constexpr (D == 3) {
        a = a + gv[2] * mv[2];
        b = b + gw[2] * vv[2];
      }
      on[c * g.V] = a - b;
    }
  }
}

// Ad_star(phiinv, m)_c = sum_d (D_d phiinv_c + delta_cd) * m_d(x + phiinv(x))   (adjrep.py:86-97)
template <typename R, int D>
__global__ void __launch_bounds__(kThreads)
Ad_star_kernel(R* __restrict__ out, const R* __restrict__ phi, const R* __restrict__ m, Geom<D> g) {
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* pn = phi + n * D * g.V + vid;
  const R* mn = m + n * D * g.V;
  Axis<R> ax[D];
#pragma unroll
  for (int a = 0; a < D; ++a) ax[a] = axis_setup(coord<R>(pos[a], 1.0, pn[a * g.V]), g.n[a]);
  R mphi[D];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    if constexpr (D == 2) {
      mphi[c] = lerp2<R>(mn + c * g.V, ax[0], ax[1], g.st[0]);
    } else {
      Corners3<R> k = gather3<R>(mn + c * g.V, ax[0], ax[1], ax[2], g.st[0], g.st[1]);
      mphi[c] = lerp3_eval<R>(k, ax[0].t, ax[1].t, ax[2].t);
    }
  }
  R* on = out + n * D * g.V + vid;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    R gr[D];
    grad<R, D>(pn + c * g.V, pos, g, gr);
    gr[c] += R(1);
    R s = gr[0] * mphi[0] + gr[1] * mphi[1];
    if constexpr (D == 3) s = s + gr[2] * mphi[2];
    on[c * g.V] = s;
  }
}

// compose(u, v, ds, dt) = ds*u(x) + dt*v(x + ds*u(x))   (deform.py:53-55). The two
// scalings and the add are rounded separately, like the reference's three ATen ops.
template <typename R, int D>
__global__ void __launch_bounds__(kThreads)
compose_kernel(R* __restrict__ out, const R* __restrict__ u, const R* __restrict__ v, Geom<D> g,
               double ds, double dt) {
  const long long vid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vid >= g.V) return;
  const long long n = blockIdx.y;
  int pos[D];
  decode<D>(vid, g, pos);
  const R* un = u + n * D * g.V + vid;
  const R* vn = v + n * D * g.V;
  R uv[D];
  Axis<R> ax[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    uv[a] = un[a * g.V];
    ax[a] = axis_setup(coord<R>(pos[a], ds, uv[a]), g.n[a]);
  }
  const R dsr = (R)ds, dtr = (R)dt;
  R* on = out + n * D * g.V + vid;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    R val;
    if constexpr (D == 2) {
      val = lerp2<R>(vn + c * g.V, ax[0], ax[1], g.st[0]);
    } else {
      Corners3<R> k = gather3<R>(vn + c * g.V, ax[0], ax[1], ax[2], g.st[0], g.st[1]);
      val = lerp3_eval<R>(k, ax[0].t, ax[1].t, ax[2].t);
    }
    if constexpr (sizeof(R) == 4) {
      on[c * g.V] = __fadd_rn(__fmul_rn((float)dsr, (float)uv[c]), __fmul_rn((float)dtr, (float)val));
    } else {
      on[c * g.V] = __dadd_rn(__dmul_rn((double)dsr, (double)uv[c]), __dmul_rn((double)dtr, (double)val));
    }
  }
}

// ---------------------------------------------------------------- host side
template <int D>
static bool thin(const int64_t* shape) {
  for (int a = 0; a < D; ++a)
    if (shape[a] < 2) return true;
  return false;
}

template <typename R, int D>
static int jtvf_fwd_t(void* out, const void* v, const void* w, int64_t N, int64_t C,
                      const int64_t* shape, int disp, int trans, cudaStream_t s) {
  LGM_REQUIRE(!thin<D>(shape), "Jacobian times vectorfield not implemented for 'thin' dimensions");
  LGM_REQUIRE(!disp || C == D, "Displacement mode only defined for vector fields");
  LGM_REQUIRE(!trans || C == D, "Jacobian transpose only implemented for vector fields");
  Geom<D> g = make_geom<D>(shape);
  if (N == 0 || C == 0) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
#define L(DI, TR)                                                                                              \
  do {                                                                                                         \
    if (C == D) jtvf_fwd_kernel<R, D, DI, TR, D><<<grid, kThreads, 0, s>>>((R*)out, (const R*)v, (const R*)w, g, (int)C); \
    else jtvf_fwd_kernel<R, D, DI, TR, 0><<<grid, kThreads, 0, s>>>((R*)out, (const R*)v, (const R*)w, g, (int)C);        \
  } while (0)
  if (disp && trans) L(true, true);
  else if (disp) L(true, false);
  else if (trans) L(false, true);
  else L(false, false);
#undef L
  count_launch("jtvf_fwd", s);
  return finish(s, "lgm_jtvf_fwd");
}

template <typename R, int D, bool DI, bool TR>
static void jtvf_bwd_launch(dim3 grid, cudaStream_t s, void* d_v, void* d_w, const void* go,
                            const void* v, const void* w, const Geom<D>& g, int C) {
#define L(NV, NW)                                                                                              \
  do {                                                                                                         \
    if (C == D) jtvf_bwd_kernel<R, D, DI, TR, NV, NW, D><<<grid, kThreads, 0, s>>>((R*)d_v, (R*)d_w, (const R*)go, (const R*)v, (const R*)w, g, C); \
    else jtvf_bwd_kernel<R, D, DI, TR, NV, NW, 0><<<grid, kThreads, 0, s>>>((R*)d_v, (R*)d_w, (const R*)go, (const R*)v, (const R*)w, g, C);        \
  } while (0)
  if (d_v && d_w) L(true, true);
  else if (d_v) L(true, false);
  else L(false, true);
#undef L
}

template <typename R, int D>
static int jtvf_bwd_t(void* d_v, void* d_w, const void* go, const void* v, const void* w, int64_t N,
                      int64_t C, const int64_t* shape, int disp, int trans, cudaStream_t s) {
  LGM_REQUIRE(!thin<D>(shape), "Jacobian times vectorfield not implemented for 'thin' dimensions");
  LGM_REQUIRE(!disp || C == D, "Displacement mode only defined for vector fields");
  LGM_REQUIRE(!trans || C == D, "Jacobian transpose only implemented for vector fields");
  Geom<D> g = make_geom<D>(shape);
  if (N == 0 || C == 0 || (!d_v && !d_w)) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  if (disp && trans) jtvf_bwd_launch<R, D, true, true>(grid, s, d_v, d_w, go, v, w, g, (int)C);
  else if (disp) jtvf_bwd_launch<R, D, true, false>(grid, s, d_v, d_w, go, v, w, g, (int)C);
  else if (trans) jtvf_bwd_launch<R, D, false, true>(grid, s, d_v, d_w, go, v, w, g, (int)C);
  else jtvf_bwd_launch<R, D, false, false>(grid, s, d_v, d_w, go, v, w, g, (int)C);
  count_launch("jtvf_bwd", s);
  return finish(s, "lgm_jtvf_bwd");
}

template <typename R, int D>
static int jtvf_adj_fwd_t(void* out, const void* z, const void* w, int64_t N, int64_t C,
                          const int64_t* shape, cudaStream_t s) {
  LGM_REQUIRE(!thin<D>(shape), "Jacobian times vectorfield not implemented for 'thin' dimensions");
  Geom<D> g = make_geom<D>(shape);
  if (N == 0 || C == 0) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  if (C == D) jtvf_adj_fwd_kernel<R, D, D><<<grid, kThreads, 0, s>>>((R*)out, (const R*)z, (const R*)w, g, (int)C);
  else jtvf_adj_fwd_kernel<R, D, 0><<<grid, kThreads, 0, s>>>((R*)out, (const R*)z, (const R*)w, g, (int)C);
  count_launch("jtvf_adj_fwd", s);
  return finish(s, "lgm_jtvf_adj_fwd");
}

template <typename R, int D>
static int jtvf_adj_bwd_t(void* d_z, void* d_w, const void* go, const void* z, const void* w,
                          int64_t N, int64_t C, const int64_t* shape, cudaStream_t s) {
  LGM_REQUIRE(!thin<D>(shape), "Jacobian times vectorfield not implemented for 'thin' dimensions");
  LGM_REQUIRE(C == D, "vector field is of wrong dimension");
  Geom<D> g = make_geom<D>(shape);
  if (N == 0 || (!d_z && !d_w)) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
#define L(NZ, NW) jtvf_adj_bwd_kernel<R, D, NZ, NW><<<grid, kThreads, 0, s>>>((R*)d_z, (R*)d_w, (const R*)go, (const R*)z, (const R*)w, g)
  if (d_z && d_w) L(true, true);
  else if (d_z) L(true, false);
  else L(false, true);
#undef L
  count_launch("jtvf_adj_bwd", s);
  return finish(s, "lgm_jtvf_adj_bwd");
}

template <typename R, int D, bool STAR>
static int ad_t(void* out, const void* v, const void* m, int64_t N, const int64_t* shape,
                cudaStream_t s) {
  LGM_REQUIRE(!thin<D>(shape), "Jacobian times vectorfield not implemented for 'thin' dimensions");
  Geom<D> g = make_geom<D>(shape);
  if (N == 0) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  ad_kernel<R, D, STAR><<<grid, kThreads, 0, s>>>((R*)out, (const R*)v, (const R*)m, g);
  count_launch("ad", s);
  return finish(s, STAR ? "lgm_ad_star_fwd" : "lgm_ad_fwd");
}
template <typename R, int D>
static int ad_star_t(void* o, const void* v, const void* m, int64_t N, const int64_t* sh, cudaStream_t s) {
  return ad_t<R, D, true>(o, v, m, N, sh, s);
}
template <typename R, int D>
static int ad_plain_t(void* o, const void* v, const void* m, int64_t N, const int64_t* sh, cudaStream_t s) {
  return ad_t<R, D, false>(o, v, m, N, sh, s);
}

template <typename R, int D>
int Ad_star_generic(void* out, const void* phi, const void* m, int64_t N, const int64_t* shape,
                    cudaStream_t s) {
  LGM_REQUIRE(!thin<D>(shape), "Jacobian times vectorfield not implemented for 'thin' dimensions");
  Geom<D> g = make_geom<D>(shape);
  if (N == 0) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  Ad_star_kernel<R, D><<<grid, kThreads, 0, s>>>((R*)out, (const R*)phi, (const R*)m, g);
  count_launch("Ad_star", s);
  return finish(s, "lgm_Ad_star_fwd");
}

template <typename R, int D>
int compose_generic(void* out, const void* u, const void* v, int64_t N, const int64_t* shape,
                    double ds, double dt, cudaStream_t s) {
  Geom<D> g = make_geom<D>(shape);
  if (N == 0 || g.V == 0) return LGM_OK;
  dim3 grid((unsigned)cdiv(g.V, kThreads), (unsigned)N);
  compose_kernel<R, D><<<grid, kThreads, 0, s>>>((R*)out, (const R*)u, (const R*)v, g, ds, dt);
  count_launch("compose", s);
  return finish(s, "lgm_compose_fwd");
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

#define CHECK_N(N)                                                                                     \
  LGM_REQUIRE((N) >= 0 && (N) <= 65535, "batch size out of range");                                    \
  LGM_REQUIRE((dim == 2 && geom_fits<2>(shape)) || (dim == 3 && geom_fits<3>(shape)),                  \
              "Only two- and three-dimensional fields of fewer than 2^31 voxels are supported")

namespace lgm {  // fp32 3-D fast paths (stencil3.cu); LGM_EUNSUP = not applicable
int ad_star3_f32(void* out, const void* v, const void* m, int64_t N, const int64_t* sh, cudaStream_t s);
int jtvf_adj3_f32(void* out, const void* z, const void* w, int64_t N, const int64_t* sh, cudaStream_t s);
}  // namespace lgm

extern "C" int lgm_jtvf_fwd(int dtype, void* out, const void* v, const void* w, int64_t N,
                            int64_t C, int dim, const int64_t* shape, int displacement,
                            int transpose, void* stream) {
  CHECK_N(N);
  DISPATCH_RD(dtype, dim, jtvf_fwd_t, out, v, w, N, C, shape, displacement, transpose, (cudaStream_t)stream);
}
extern "C" int lgm_jtvf_bwd(int dtype, void* d_v, void* d_w, const void* gout, const void* v,
                            const void* w, int64_t N, int64_t C, int dim, const int64_t* shape,
                            int displacement, int transpose, void* stream) {
  CHECK_N(N);
  if (dtype == LGM_F32 && dim == 3 && C == 3 && N > 0 && !thin<3>(shape)) {
    // fp32 3-D vector fields: both gradients are operators that have tuned kernels of their own
    //   d_w = jtvf(v, gout, displacement, !transpose)                  (diff.cu:311-335, :417-431)
    //   d_v = sum_d D_d^T (w_d gout_c)  resp.  sum_d D_d^T (gout_d w_c)  (diff.cu:432-460, :336-407)
    // (same arithmetic and accumulation order as jtvf_bwd_kernel, about half its time)
    cudaStream_t s = (cudaStream_t)stream;
    int rc = LGM_OK;
    if (d_v) {
      rc = transpose ? jtvf_adj3_f32(d_v, w, gout, N, shape, s) : jtvf_adj3_f32(d_v, gout, w, N, shape, s);
      if (rc == LGM_EUNSUP) goto generic;
      if (rc) return rc;
    }
    if (d_w) rc = jtvf_fwd_t<float, 3>(d_w, v, gout, N, C, shape, displacement, !transpose, s);
    return rc;
  }
generic:
  DISPATCH_RD(dtype, dim, jtvf_bwd_t, d_v, d_w, gout, v, w, N, C, shape, displacement, transpose, (cudaStream_t)stream);
}
extern "C" int lgm_jtvf_adj_fwd(int dtype, void* out, const void* z, const void* w, int64_t N,
                                int64_t C, int dim, const int64_t* shape, void* stream) {
  CHECK_N(N);
  if (dtype == LGM_F32 && dim == 3 && C == 3 && N > 0) {
    int rc = jtvf_adj3_f32(out, z, w, N, shape, (cudaStream_t)stream);
    if (rc != LGM_EUNSUP) return rc;
  }
  DISPATCH_RD(dtype, dim, jtvf_adj_fwd_t, out, z, w, N, C, shape, (cudaStream_t)stream);
}
extern "C" int lgm_jtvf_adj_bwd(int dtype, void* d_z, void* d_w, const void* gout, const void* z,
                                const void* w, int64_t N, int64_t C, int dim, const int64_t* shape,
                                void* stream) {
  CHECK_N(N);
  DISPATCH_RD(dtype, dim, jtvf_adj_bwd_t, d_z, d_w, gout, z, w, N, C, shape, (cudaStream_t)stream);
}
extern "C" int lgm_ad_star_fwd(int dtype, void* out, const void* v, const void* m, int64_t N,
                               int dim, const int64_t* shape, void* stream) {
  CHECK_N(N);
  if (dtype == LGM_F32 && dim == 3 && N > 0) {
    int rc = ad_star3_f32(out, v, m, N, shape, (cudaStream_t)stream);
    if (rc != LGM_EUNSUP) return rc;
  }
  DISPATCH_RD(dtype, dim, ad_star_t, out, v, m, N, shape, (cudaStream_t)stream);
}
extern "C" int lgm_ad_fwd(int dtype, void* out, const void* v, const void* w, int64_t N, int dim,
                          const int64_t* shape, void* stream) {
  CHECK_N(N);
  DISPATCH_RD(dtype, dim, ad_plain_t, out, v, w, N, shape, (cudaStream_t)stream);
}
namespace lgm {  // fp32 3-D fast paths (gather3.cu); LGM_EUNSUP = not applicable
int Ad_star3_f32(void* out, const void* phi, const void* m, int64_t N, const int64_t* sh, int rev, cudaStream_t s);
int compose3_f32(void* out, const void* u, const void* v, int64_t N, const int64_t* sh, double ds, double dt,
                 int rev, cudaStream_t s);
}  // namespace lgm

extern "C" int lgm_Ad_star_fwd(int dtype, void* out, const void* phiinv, const void* m, int64_t N,
                               int dim, const int64_t* shape, void* stream) {
  CHECK_N(N);
  if (dtype == LGM_F32 && dim == 3 && N > 0) {
    int rc = Ad_star3_f32(out, phiinv, m, N, shape, 0, (cudaStream_t)stream);
    if (rc != LGM_EUNSUP) return rc;
  }
  DISPATCH_RD(dtype, dim, Ad_star_generic, out, phiinv, m, N, shape, (cudaStream_t)stream);
}
extern "C" int lgm_compose_fwd(int dtype, void* out, const void* u, const void* v, int64_t N,
                               int dim, const int64_t* shape, double ds, double dt, void* stream) {
  CHECK_N(N);
  if (dtype == LGM_F32 && dim == 3 && N > 0) {
    int rc = compose3_f32(out, u, v, N, shape, ds, dt, 0, (cudaStream_t)stream);
    if (rc != LGM_EUNSUP) return rc;
  }
  DISPATCH_RD(dtype, dim, compose_generic, out, u, v, N, shape, ds, dt, (cudaStream_t)stream);
}
