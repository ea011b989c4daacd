// shoot3.cu -- the whole forward EPDiff shoot (lagomorph/lddmm.py:73-91, non-checkpointed branch) as
// ONE library call: num_steps x (Ad_star, sharp, compose) enqueued back to back on the caller's stream
// with explicit traversal directions, the displacement ping-ponging between the output and one
// scratch field. Same kernels and results as calling lgm_epdiff_step_fwd num_steps times.
//
// Measured and rejected here (round 2, profiles/r2_exp1_packed_gathers.log): keeping m0 and phiinv
// channel-packed (one float4 per voxel) between the steps so that every trilinear corner / stencil
// neighbour is ONE 128-bit load (Ad_star 48 -> 15 load instructions, compose 27 -> 11). Slower on
// B200: Ad_star 0.434 -> 0.483 ms, compose 0.326 -> 0.338 ms at C2 (+4 B/voxel of traffic each). A
// warp-wide LDG.128 is 4-5 L1 wavefronts, so the gathers move the same number of wavefronts through
// the L1 data pipe either way: that pipe (not the count of load instructions) bounds these kernels.
#include <cstdlib>
#include "common.cuh"

namespace lgm {

int Ad_star3_f32(void* out, const void* phi, const void* m, int64_t N, const int64_t* sh, int rev, cudaStream_t s);
int compose3_f32(void* out, const void* u, const void* v, int64_t N, const int64_t* sh, double ds, double dt,
                 int rev, cudaStream_t s);
int fluid_apply_dir(int dtype, void* out, const void* in, int64_t N, int dim, const int64_t* shape, int inverse,
                    double alpha, double beta, double gamma, void* ws, int64_t ws_bytes, int rev, cudaStream_t s);
int fluid_apply_post(int dtype, void* out, const void* in, int64_t N, int dim, const int64_t* shape, int inverse,
                     double alpha, double beta, double gamma, void* ws, int64_t ws_bytes, int rev, const double* post,
                     cudaStream_t s);

template <typename R>
__global__ void mul_mask2_kernel(R* __restrict__ m, const R* __restrict__ mask, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) m[i] *= mask[i];
}

static size_t align_up256(size_t v) { return (v + 255) / 256 * 256; }

// First step from the identity (phiinv_in == NULL, no mask): Ad_star(0, m0) = m0 and
// compose_disp_vel(0, v, -dt) = -dt*v, so the step is ONE sharp whose last kernel scales its output
// by -dt (fl(fl(-dt * v) + 0), the compose stage's own rounding): 3 launches instead of 5, no memset,
// 24 B per voxel instead of 96. Same bits as the full step except the sign of exact zeros (the full
// step turns -0 into +0). LGM_NO_FIRST_STEP_SHORTCUT=1: the full step (kernel experiments, tests).
static bool first_step_shortcut() {
  static const bool on = getenv("LGM_NO_FIRST_STEP_SHORTCUT") == nullptr;
  return on;
}

static bool alternate_enabled() {
  static const bool on = getenv("LGM_NO_ALTERNATE") == nullptr;  // kernel experiments
  return on;
}

}  // namespace lgm

using namespace lgm;

extern "C" int64_t lgm_expmap_scratch_bytes(int dtype, int64_t N, int dim, const int64_t* shape) {
  const int64_t step = lgm_epdiff_scratch_bytes(dtype, N, dim, shape);
  if (step < 0) return -1;
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  const size_t esz = dtype == LGM_F32 ? 4 : 8;
  // the step's scratch (momentum / velocity field + FFT workspace) + one displacement field
  return (int64_t)(align_up256((size_t)step) + align_up256((size_t)(N * dim * V) * esz));
}

extern "C" int lgm_expmap_fwd(int dtype, void* phiinv_out, const void* phiinv_in, const void* m0,
                              const void* mommask, int64_t N, int dim, const int64_t* shape, double dt,
                              int num_steps, double alpha, double beta, double gamma, void* scratch,
                              int64_t scratch_bytes, void* stream) {
  LGM_REQUIRE(dim == 2 || dim == 3, "lgm_expmap_fwd: dim must be 2 or 3");
  LGM_REQUIRE(dtype == LGM_F32 || dtype == LGM_F64, "lgm_expmap_fwd: unsupported dtype %d", dtype);
  LGM_REQUIRE(num_steps >= 1, "lgm_expmap_fwd: num_steps must be >= 1");
  LGM_REQUIRE(phiinv_out && m0 && scratch, "lgm_expmap_fwd: null pointer");
  LGM_REQUIRE(phiinv_out != phiinv_in, "lgm_expmap_fwd: phiinv_out must not alias phiinv_in");
  LGM_REQUIRE(N >= 0 && N <= 21845, "lgm_expmap_fwd: batch size out of range");
  const int64_t need = lgm_expmap_scratch_bytes(dtype, N, dim, shape);
  if (need < 0 || scratch_bytes < need)
    return set_error(LGM_ENOSPC, "lgm_expmap_fwd: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes,
                     (long long)need);
  cudaStream_t s = (cudaStream_t)stream;
  long long V = 1;
  for (int a = 0; a < dim; ++a) V *= shape[a];
  if (N == 0 || V == 0) return LGM_OK;
  const size_t esz = dtype == LGM_F32 ? 4 : 8;
  const size_t field_bytes = (size_t)(N * dim * V) * esz;
  const size_t field = align_up256(field_bytes);
  const int64_t step_bytes = lgm_epdiff_scratch_bytes(dtype, N, dim, shape);
  void* m = scratch;                   // momentum, then velocity in place
  void* ws = (char*)scratch + field;   // FFT workspace
  const int64_t ws_bytes = step_bytes - (int64_t)field;
  void* tmp = (char*)scratch + align_up256((size_t)step_bytes);

  // the steps ping-pong between phiinv_out and tmp so that the last one lands in phiinv_out
  const void* cur = phiinv_in;
  const bool shortcut = !phiinv_in && !mommask && first_step_shortcut();
  if (!cur && !shortcut) {  // phiinv = zeros (lddmm.py:84-85)
    void* z = (num_steps & 1) ? tmp : phiinv_out;
    cudaError_t e = cudaMemsetAsync(z, 0, field_bytes, s);
    if (e != cudaSuccess) return set_error((int)e, "lgm_expmap_fwd: memset: %s", cudaGetErrorString(e));
    cur = z;
  }
  const bool alt = alternate_enabled();
  const bool fast3 = (dtype == LGM_F32 && dim == 3);
  for (int k = 0; k < num_steps; ++k) {
    void* dst = ((num_steps - 1 - k) & 1) ? tmp : phiinv_out;
    // traversal directions: Ad_star and compose walk in direction p, the slab passes of sharp in !p
    // (its X pass in p): every kernel starts on the data its predecessor wrote last (L2); p flips
    // every step because compose(p) leaves the far end of phiinv for the next step's Ad_star.
    const int p = alt ? (k & 1) : 0;
    int rc = LGM_EUNSUP;
    if (k == 0 && shortcut) {
      // slab passes ascending: the next step's Ad_star (direction 1) starts on what was written last
      const double post = -dt;
      rc = fluid_apply_post(dtype, dst, m0, N, dim, shape, 1, alpha, beta, gamma, ws, ws_bytes, 0, &post, s);
      if (rc) return rc;
      cur = dst;
      continue;
    }
    if (fast3) rc = Ad_star3_f32(m, cur, m0, N, shape, p, s);
    if (rc == LGM_EUNSUP) rc = lgm_Ad_star_fwd(dtype, m, cur, m0, N, dim, shape, stream);
    if (rc) return rc;
    if (mommask) {  // `m = m * mommask` (lddmm.py:41-42), full-shape mask
      const long long total = N * dim * V;
      if (dtype == LGM_F32)
        mul_mask2_kernel<float><<<(unsigned)cdiv(total, 256), 256, 0, s>>>((float*)m, (const float*)mommask, total);
      else
        mul_mask2_kernel<double><<<(unsigned)cdiv(total, 256), 256, 0, s>>>((double*)m, (const double*)mommask, total);
      count_launch("mul_mask", s);
    }
    rc = fluid_apply_dir(dtype, m, m, N, dim, shape, 1, alpha, beta, gamma, ws, ws_bytes, alt ? !p : 0, s);
    if (rc) return rc;
    // compose_disp_vel(phiinv, v, -dt) = compose(v, phiinv, ds=-dt, dt=1)  (deform.py:58-62)
    rc = LGM_EUNSUP;
    if (fast3) rc = compose3_f32(dst, m, cur, N, shape, -dt, 1.0, p, s);
    if (rc == LGM_EUNSUP) rc = lgm_compose_fwd(dtype, dst, m, cur, N, dim, shape, -dt, 1.0, stream);
    if (rc) return rc;
    cur = dst;
  }
  return LGM_OK;
}
