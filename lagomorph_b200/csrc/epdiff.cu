// epdiff.cu -- one forward EPDiff step (lagomorph/lddmm.py:39-44) as a fixed
// sequence of launches on the caller's stream, capturable in a CUDA graph:
//   m   = Ad_star(phiinv, m0)            fused gather + Jacobian        (diff.cu)
//   v   = sharp(m)                       FFT passes + fused multiplier (fluid.cu)
//   out = -dt*v + phiinv(x - dt*v)       fused gather + axpy           (diff.cu)
#include <cstdlib>
#include "common.cuh"

namespace lgm {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

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
    int rc = lgm_Ad_star_fwd(dtype, m, phi_g, m0_g, g, dim, shape, stream);
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
    rc = lgm_fluid_apply(dtype, m, m, g, dim, shape, 1, alpha, beta, gamma, ws, scratch_bytes - (int64_t)field, stream);
    if (rc) return rc;
    // compose_disp_vel(phiinv, v, -dt) = compose(v, phiinv, ds=-dt, dt=1)  (deform.py:58-62)
    rc = lgm_compose_fwd(dtype, out_g, m, phi_g, g, dim, shape, -dt, 1.0, stream);
    if (rc) return rc;
  }
  return LGM_OK;
}
