// epdiff.cu -- one forward EPDiff step (lagomorph/lddmm.py:39-44) as a fixed
// sequence of launches on the caller's stream, capturable in a CUDA graph:
//   m   = Ad_star(phiinv, m0)            fused gather + Jacobian        (diff.cu)
//   v   = sharp(m)                       FFT passes + fused multiplier (fluid.cu)
//   out = -dt*v + phiinv(x - dt*v)       fused gather + axpy           (diff.cu)
#include <cstdlib>
#include "common.cuh"

namespace lgm {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// fp32 3-D fast paths (gather3.cu) and sharp with an explicit traversal direction (fluid.cu)
int Ad_star3_f32(void* out, const void* phi, const void* m, int64_t N, const int64_t* sh, int rev, cudaStream_t s);
int compose3_f32(void* out, const void* u, const void* v, int64_t N, const int64_t* sh, double ds, double dt,
                 int rev, cudaStream_t s);
int fluid_apply_dir(int dtype, void* out, const void* in, int64_t N, int dim, const int64_t* shape, int inverse,
                    double alpha, double beta, double gamma, void* ws, int64_t ws_bytes, int rev, cudaStream_t s);

template <typename R>
__global__ void mul_mask_kernel(R* __restrict__ m, const R* __restrict__ mask, long long total,
                                long long mask_total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) m[i] *= mask[i % mask_total];
}

}  // namespace lgm

using namespace lgm;

extern "C" int64_t lgm_epdiff_scratch_bytes(int dtype, int64_t N, int dim, const int64_t* shape) {
  if ((dim != 2 && dim != 3) || (dtype != LGM_F32 && dtype != LGM_F64)) return -1;
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  const size_t esz = dtype == LGM_F32 ? 4 : 8;
  const size_t field = align_up((size_t)(N * dim * V) * esz, 256);
  return (int64_t)(field + (size_t)lgm_fluid_workspace_bytes(dtype, N, dim, shape));
}

extern "C" int lgm_epdiff_step_fwd(int dtype, void* phiinv_out, const void* phiinv, const void* m0,
                                   const void* mommask, int64_t N, int dim, const int64_t* shape,
                                   double dt, double alpha, double beta, double gamma, void* scratch,
                                   int64_t scratch_bytes, void* stream) {
  LGM_REQUIRE(dim == 2 || dim == 3, "lgm_epdiff_step_fwd: dim must be 2 or 3");
  LGM_REQUIRE(phiinv_out != phiinv, "lgm_epdiff_step_fwd: phiinv_out must not alias phiinv");
  const int64_t need = lgm_epdiff_scratch_bytes(dtype, N, dim, shape);
  if (need < 0 || scratch_bytes < need)
    return set_error(LGM_ENOSPC, "lgm_epdiff_step_fwd: scratch too small (%lld < %lld bytes)",
                     (long long)scratch_bytes, (long long)need);
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  const size_t esz = dtype == LGM_F32 ? 4 : 8;
  const size_t field = align_up((size_t)(N * dim * V) * esz, 256);
  // Chunked over subjects: a chunk's momentum/velocity scratch and its spectrum stay resident in L2
  // between the five launches, so per voxel-step only phiinv, the m0 gather and the result touch HBM.
  void* m = scratch;
  void* ws = (char*)scratch + field;
  const size_t spec_per_subject = (size_t)dim * (V / shape[dim - 1]) * (shape[dim - 1] / 2 + 1) * 2 * esz;
  static long long budget = -1;
  if (budget < 0) {
    const char* e = getenv("LGM_FLUID_CHUNK_MB");
    budget = (e && atoll(e) > 0) ? atoll(e) << 20 : 1LL << 50;  // default: no chunking (see fluid.cu)
  }
  long long G = budget / (long long)spec_per_subject;
  if (G < 1) G = 1;
  if (G > N) G = N;
  const size_t sub = (size_t)dim * V * esz;  // bytes of one subject's vector field
  for (long long n0 = 0; n0 < N; n0 += G) {
    const long long g = (N - n0 < G) ? (N - n0) : G;
    const char* phi_g = (const char*)phiinv + n0 * sub;
    const char* m0_g = (const char*)m0 + n0 * sub;
    char* out_g = (char*)phiinv_out + n0 * sub;
    // traversal directions (explicit arguments, no state across calls): Ad_star and compose ascending,
    // the slab passes of sharp descending (its X pass ascending again), so that each kernel starts on
    // the data its predecessor wrote last (L2). lgm_expmap_fwd additionally flips all of them from
    // step to step.
    static const bool alternate = getenv("LGM_NO_ALTERNATE") == nullptr;
    int rc = LGM_EUNSUP;
    if (dtype == LGM_F32 && dim == 3) rc = Ad_star3_f32(m, phi_g, m0_g, g, shape, 0, (cudaStream_t)stream);
    if (rc == LGM_EUNSUP) rc = lgm_Ad_star_fwd(dtype, m, phi_g, m0_g, g, dim, shape, stream);
    if (rc) return rc;
    if (mommask) {  // full-shape mask (N,dim,...), applied like `m = m * mommask` (lddmm.py:41-42)
      const long long total = g * dim * V;
      const char* mk = (const char*)mommask + n0 * sub;
      if (dtype == LGM_F32)
        mul_mask_kernel<float><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((float*)m, (const float*)mk, total, total);
      else
        mul_mask_kernel<double><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((double*)m, (const double*)mk, total, total);
      count_launch("mul_mask", (cudaStream_t)stream);
    }
    rc = fluid_apply_dir(dtype, m, m, g, dim, shape, 1, alpha, beta, gamma, ws, scratch_bytes - (int64_t)field,
                         alternate ? 1 : 0, (cudaStream_t)stream);
    if (rc) return rc;
    // compose_disp_vel(phiinv, v, -dt) = compose(v, phiinv, ds=-dt, dt=1)  (deform.py:58-62)
    rc = LGM_EUNSUP;
    if (dtype == LGM_F32 && dim == 3) rc = compose3_f32(out_g, m, phi_g, g, shape, -dt, 1.0, 0, (cudaStream_t)stream);
    if (rc == LGM_EUNSUP) rc = lgm_compose_fwd(dtype, out_g, m, phi_g, g, dim, shape, -dt, 1.0, stream);
    if (rc) return rc;
  }
  return LGM_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of one step (fp32 3-D; see epdiff_bwd.cu for the kernels).
// ------------------------------------------------------------------------------------------------
namespace lgm {
bool epdiff_bwd3_ok(int64_t N, const int64_t* sh);
int compose_bwd3_f32(void* dv, void* S, const void* G, const void* phi, const void* v, int64_t N,
                     const int64_t* sh, double ds, bool need_phi, cudaStream_t s);
int adstar_bwd3_f32(void* mi, void* d_m0, void* S, const void* phi, const void* dm, const void* m0, int64_t N,
                    const int64_t* sh, bool need_m0, bool need_phi, cudaStream_t s);
int stencil_bwd3_f32(void* G, void* S, const void* mi, const void* dm, int64_t N, const int64_t* sh,
                     cudaStream_t s);
}  // namespace lgm

extern "C" int64_t lgm_epdiff_bwd_scratch_bytes(int dtype, int64_t N, int dim, const int64_t* shape) {
  if (dtype != LGM_F32 || dim != 3 || N < 1 || !epdiff_bwd3_ok(N, shape)) return -1;
  const size_t field = align_up((size_t)(N * 3 * shape[0] * shape[1] * shape[2]) * 4, 256);
  return (int64_t)(2 * field + (size_t)lgm_fluid_workspace_bytes(dtype, N, dim, shape));
}

extern "C" int lgm_epdiff_step_bwd(int dtype, void* g_phi, void* d_m0, void* splat_acc, const void* phiinv,
                                   const void* v, const void* m0, const void* mommask, int64_t N, int dim,
                                   const int64_t* shape, double dt, double alpha, double beta, double gamma,
                                   void* scratch, int64_t scratch_bytes, int need_phi, int need_m0,
                                   void* stream) {
  const int64_t need = lgm_epdiff_bwd_scratch_bytes(dtype, N, dim, shape);
  if (need < 0) return set_error(LGM_EUNSUP, "lgm_epdiff_step_bwd: only fp32 3-D volumes with Z %% 32 == 0");
  if (scratch_bytes < need)
    return set_error(LGM_ENOSPC, "lgm_epdiff_step_bwd: scratch too small (%lld < %lld bytes)",
                     (long long)scratch_bytes, (long long)need);
  LGM_REQUIRE(g_phi && phiinv && v && m0 && scratch, "lgm_epdiff_step_bwd: null pointer");
  LGM_REQUIRE(!need_phi || splat_acc, "lgm_epdiff_step_bwd: splat_acc required when need_phi");
  LGM_REQUIRE(!need_m0 || d_m0, "lgm_epdiff_step_bwd: d_m0 required when need_m0");
  cudaStream_t s = (cudaStream_t)stream;
  const long long V = shape[0] * shape[1] * shape[2];
  const size_t field = align_up((size_t)(N * 3 * V) * 4, 256);
  void* W = scratch;                        // d_v, then d_m in place
  void* MI = (char*)scratch + field;        // m0(x + phiinv), for the stencil pass
  void* ws = (char*)scratch + 2 * field;
  // the gradient w.r.t. v is needed for both outputs (d_m0 and d_phiinv depend on d_m = sharp(d_v))
  int rc = compose_bwd3_f32(W, splat_acc, g_phi, phiinv, v, N, shape, -dt, need_phi != 0, s);
  if (rc) return rc;
  rc = lgm_fluid_apply(dtype, W, W, N, dim, shape, 1, alpha, beta, gamma, ws, scratch_bytes - (int64_t)(2 * field), stream);
  if (rc) return rc;
  if (mommask) {
    const long long total = N * 3 * V;
    mul_mask_kernel<float><<<(unsigned)cdiv(total, 256), 256, 0, s>>>((float*)W, (const float*)mommask, total, total);
    count_launch("mul_mask", s);
  }
  if (!need_phi && !need_m0) return LGM_OK;
  rc = adstar_bwd3_f32(MI, d_m0, splat_acc, phiinv, W, m0, N, shape, need_m0 != 0, need_phi != 0, s);
  if (rc) return rc;
  if (need_phi) rc = stencil_bwd3_f32(g_phi, splat_acc, MI, W, N, shape, s);
  return rc;
}
