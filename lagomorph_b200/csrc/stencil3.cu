// stencil3.cu -- fp32 3-D fast paths of the two pure-stencil operators that contain the transposed
// central difference D_d^T:
//   ad_star(v, m)_c   = sum_d (D_c v_d) m_d - sum_d D_d^T (v_d m_c)        (adjrep.py:69-83)
//   jtvf_adjoint(z, w)_c = sum_d D_d^T (w_d z_c)                            (cuda/diff.cu:546-632)
// Same work layout as gather3.cu: lane = z, a warp walks one z row in NV chunks of 32, a CTA covers 8
// y rows of one x slab, every address is base + 4*index with a 32-bit in-volume index (one IMAD.WIDE).
//
// D_d^T is evaluated without branches. With P = p*q, clamped neighbour offsets and a sign per side
// (cuda/diff.cu:432-460, the exact transpose of the clamped central difference incl. its boundary rows):
//   interior : 0.5*( P[i-1] - P[i+1])
//   i == 0   : 0.5*(-P[0]   - P[1])        lower neighbour clamps onto the voxel itself, sign -1
//   i == n-1 : 0.5*( P[n-2] + P[n-1])      upper neighbour clamps onto the voxel itself, sign -1
// i.e. 0.5*(s_lo*P[lo] - s_hi*P[hi]) with lo = max(i-1,0), hi = min(i+1,n-1): bit-identical to the
// three-way branch of the generic kernel (diff.cu cdiffT). The plain difference D_d uses the same
// clamped neighbours without signs.
#include "gather_common.cuh"

namespace lgm {

// MODE 0: ad_star (a = v, b = m); MODE 1: jtvf_adjoint with C == 3 (a = w, b = z).
template <int MODE, int NV>
__global__ void __launch_bounds__(256, 4)
stencil3_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b, int X, int Y,
                int Z) {
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (j >= Y) return;
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const float* ab[3];
  const float* bb[3];
  float* ob[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ab[c] = a + ((size_t)n * 3 + c) * V;
    bb[c] = b + ((size_t)n * 3 + c) * V;
    ob[c] = out + ((size_t)n * 3 + c) * V;
  }
  const unsigned four = opaque_four();
  const int row = i * sx + j * sy;
  // clamped neighbour offsets and the D^T signs of the x and y directions (uniform per warp)
  int off[6];
  float sg[6];
  off[0] = (i > 0) ? -sx : 0;      sg[0] = (i > 0) ? 1.f : -1.f;
  off[1] = (i < X - 1) ? sx : 0;   sg[1] = (i < X - 1) ? 1.f : -1.f;
  off[2] = (j > 0) ? -sy : 0;      sg[2] = (j > 0) ? 1.f : -1.f;
  off[3] = (j < Y - 1) ? sy : 0;   sg[3] = (j < Y - 1) ? 1.f : -1.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = (blockIdx.x * NV + v) * 32 + threadIdx.x;
    if (k >= Z) break;
    const int c0 = row + k;
    off[4] = (k > 0) ? -1 : 0;       sg[4] = (k > 0) ? 1.f : -1.f;
    off[5] = (k < Z - 1) ? 1 : 0;    sg[5] = (k < Z - 1) ? 1.f : -1.f;
    float bc[3];  // b at the centre (ad_star: m_d)
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) bc[c] = __ldg(at4(bb[c], c0, four));
    }
    float A[3], B[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const unsigned lo = c0 + off[2 * d], hi = c0 + off[2 * d + 1];
      float blo[3], bhi[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        blo[c] = __ldg(at4(bb[c], lo, four));
        bhi[c] = __ldg(at4(bb[c], hi, four));
      }
      float alo_d, ahi_d;
      if (MODE == 0) {
        float alo[3], ahi[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          alo[c] = __ldg(at4(ab[c], lo, four));
          ahi[c] = __ldg(at4(ab[c], hi, four));
        }
        // A_d = sum_c (D_d v_c) m_c, accumulated over c as jtvf(transpose) does (diff.cu:34-44)
        A[d] = (0.5f * (ahi[0] - alo[0])) * bc[0];
        A[d] = A[d] + (0.5f * (ahi[1] - alo[1])) * bc[1];
        A[d] = A[d] + (0.5f * (ahi[2] - alo[2])) * bc[2];
        alo_d = alo[d];
        ahi_d = ahi[d];
      } else {
        alo_d = __ldg(at4(ab[d], lo, four));
        ahi_d = __ldg(at4(ab[d], hi, four));
      }
      const float slo = sg[2 * d], shi = sg[2 * d + 1];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        B[c] += 0.5f * (slo * (alo_d * blo[c]) - shi * (ahi_d * bhi[c]));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) *at4(ob[c], c0, four) = (MODE == 0) ? A[c] - B[c] : B[c];
  }
}

static bool stencil3_ok(int64_t N, const int64_t* sh) {
  if (sh[0] < 2 || sh[1] < 2 || sh[2] < 2) return false;
  if (sh[0] * sh[1] * sh[2] >= (1LL << 31) / 4) return false;
  if (N * sh[0] > 65535 || sh[1] > 65535LL * 8) return false;
  return true;
}

// LGM_EUNSUP when the fast path does not apply (the caller falls back to the generic kernel)
int ad_star3_f32(void* out, const void* v, const void* m, int64_t N, const int64_t* sh, cudaStream_t s) {
  if (!stencil3_ok(N, sh)) return LGM_EUNSUP;
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0])), block(32, 8);
  stencil3_kernel<0, 4><<<grid, block, 0, s>>>((float*)out, (const float*)v, (const float*)m, (int)sh[0], (int)sh[1],
                                               (int)sh[2]);
  count_launch("ad", s);
  return finish(s, "lgm_ad_star_fwd");
}

int jtvf_adj3_f32(void* out, const void* z, const void* w, int64_t N, const int64_t* sh, cudaStream_t s) {
  if (!stencil3_ok(N, sh)) return LGM_EUNSUP;
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0])), block(32, 8);
  stencil3_kernel<1, 4><<<grid, block, 0, s>>>((float*)out, (const float*)w, (const float*)z, (int)sh[0], (int)sh[1],
                                               (int)sh[2]);
  count_launch("jtvf_adj_fwd", s);
  return finish(s, "lgm_jtvf_adj_fwd");
}

}  // namespace lgm
