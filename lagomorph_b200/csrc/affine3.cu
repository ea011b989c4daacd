// affine3.cu -- fp32 3-D fast paths of affine_interp forward / backward (BASELINE config 4), in the work
// layout of gather3.cu: lane = z, a warp walks one z row in NV chunks of 32, a CTA covers 8 y rows of
// one x slab, 32-bit in-volume indices. Coordinates are formed exactly as the generic kernel does
// (affine.cu affine_coords: cuda/affine.cu:42-52, :81-100); the gather uses the corner-pair form of
// gather_common.cuh, the adjoint splat merges the upper-z share of lane L into the lower-z share of lane
// L+1 where both hit the same voxel (as splat3_kernel), and the d_A / d_T partial sums are carried in
// registers over a thread's voxels before ONE warp reduction and one atomic per CTA and matrix entry.
#include "gather_common.cuh"

namespace lgm {

namespace {

// h = A (x - o) + T + o, o = (n-1)/2, same expression order as affine.cu affine_coords<float, 3>
__device__ __forceinline__ void coords3(const float (&An)[9], const float (&Tn)[3], const float (&o)[3], float f0,
                                        float f1, float f2, float (&h)[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float s = An[r * 3] * f0 + An[r * 3 + 1] * f1;
    s = s + An[r * 3 + 2] * f2;
    h[r] = s + Tn[r] + o[r];
  }
}

template <int NV, int CC>
__global__ void __launch_bounds__(256)
affine3_fwd_kernel(float* __restrict__ out, const float* __restrict__ I, const float* __restrict__ A,
                   const float* __restrict__ T, int X, int Y, int Z, int C_rt, size_t I_batch_stride) {
  const int C = CC ? CC : C_rt;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (j >= Y) return;
  const int i = blockIdx.z % X;
  const int n = blockIdx.z / X;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  float An[9], Tn[3];
#pragma unroll
  for (int q = 0; q < 9; ++q) An[q] = __ldg(A + (size_t)n * 9 + q);
#pragma unroll
  for (int q = 0; q < 3; ++q) Tn[q] = __ldg(T + (size_t)n * 3 + q);
  const float o[3] = {(float)(.5 * (double)(float)(X - 1)), (float)(.5 * (double)(float)(Y - 1)),
                      (float)(.5 * (double)(float)(Z - 1))};
  const float* In = I + (size_t)n * I_batch_stride;
  float* on = out + (size_t)n * C * V;
  const unsigned four = opaque_four();
  const float hiX = (float)X - 0.5f, hiY = (float)Y - 0.5f, hiZ = (float)Z - 0.5f;
  const int row = i * sx + j * sy;
  const float f0 = (float)i - o[0], f1 = (float)j - o[1];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = (blockIdx.x * NV + v) * 32 + threadIdx.x;
    if (k >= Z) break;
    const int c0 = row + k;
    float h[3];
    coords3(An, Tn, o, f0, f1, (float)k - o[2], h);
    const Ax3 ax = axis_fwd(h[0], X, hiX), ay = axis_fwd(h[1], Y, hiY), az = axis_fwd(h[2], Z, hiZ);
    int zs;
    float wv;
    z_pair(az, Z, zs, wv);
    const unsigned rx0 = ax.i0 * sx + zs, rx1 = ax.i1 * sx + zs;
    const unsigned ry0 = ay.i0 * sy, ry1 = ay.i1 * sy;
    const unsigned i00 = rx0 + ry0, i01 = rx0 + ry1, i10 = rx1 + ry0, i11 = rx1 + ry1;
    const float omt = 1.f - ax.t, omu = 1.f - ay.t, omv = 1.f - wv;
    if constexpr (CC > 0) {
      float r[CC];
#pragma unroll
      for (int c = 0; c < CC; ++c) r[c] = trilerp(In + c * V, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv, four);
#pragma unroll
      for (int c = 0; c < CC; ++c) on[c0 + c * V] = r[c];
    } else {
      for (int c = 0; c < C; ++c)
        on[c0 + (size_t)c * V] = trilerp(In + (size_t)c * V, i00, i01, i10, i11, ax.t, ay.t, wv, omt, omu, omv, four);
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Z % 32 == 0 (whole warps per chunk: the splat's lane exchange)
// XS: x slabs walked by one CTA. The d_A / d_T partial sums live in registers across all of them, so the
// block reduction at the end (12 warp sums = 60 shuffles per warp, 12 atomics per CTA) is paid once
// per XS * 128 voxels of a warp instead of once per 128.
template <int NV, bool NEED_I, bool NEED_AT, int XS>
__global__ void __launch_bounds__(256)
affine3_bwd_kernel(float* __restrict__ d_I, float* __restrict__ d_A, float* __restrict__ d_T,
                   const float* __restrict__ go, const float* __restrict__ I, const float* __restrict__ A,
                   const float* __restrict__ T, int X, int Y, int Z, int C, size_t I_batch_stride) {
  const int j = blockIdx.y * 8 + threadIdx.y;
  const bool rowok = j < Y;  // whole warps (a warp is one row); no early return: the CTA reduces at the end
  const int XB = (X + XS - 1) / XS;
  const int i0 = (blockIdx.z % XB) * XS;
  const int n = blockIdx.z / XB;
  const int sy = Z, sx = Y * Z;
  const int V = X * sx;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x;
  float part[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) part[q] = 0.f;
  if (rowok) {
    float An[9], Tn[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) An[q] = __ldg(A + (size_t)n * 9 + q);
#pragma unroll
    for (int q = 0; q < 3; ++q) Tn[q] = __ldg(T + (size_t)n * 3 + q);
    const float o[3] = {(float)(.5 * (double)(float)(X - 1)), (float)(.5 * (double)(float)(Y - 1)),
                        (float)(.5 * (double)(float)(Z - 1))};
    const float* In = I + (size_t)n * I_batch_stride;
    float* dIn = d_I + (size_t)n * I_batch_stride;
    const float* gn = go + (size_t)n * C * V;
    const float f1 = (float)j - o[1];
#pragma unroll 1
    for (int iv = 0; iv < XS * NV; ++iv) {
      const int i = i0 + iv / NV, v = iv % NV;
      if (i >= X) break;
      const int row = i * sx + j * sy;
      const float f0 = (float)i - o[0];
      const int kb = (blockIdx.x * NV + v) * 32;
      if (kb >= Z) continue;
      const int k = kb + lane;
      const int c0 = row + k;
      const float f2 = (float)k - o[2];
      float h[3];
      coords3(An, Tn, o, f0, f1, f2, h);
      const Ax3 ax = axis_fast(h[0], X), ay = axis_fast(h[1], Y), az = axis_fast(h[2], Z);
      const unsigned rb[4] = {(unsigned)(ax.i0 * sx + ay.i0 * sy), (unsigned)(ax.i0 * sx + ay.i1 * sy),
                              (unsigned)(ax.i1 * sx + ay.i0 * sy), (unsigned)(ax.i1 * sx + ay.i1 * sy)};
      float wlo[4], whi[4];
      bool give[4], took[4];
      if (NEED_I) {  // weight sequences "d = 1 - d" (include/interp.h:437-453), as splat3_kernel
        const float wx0 = 1.f - ax.t, wx1 = 1.f - wx0;
        const float wy0 = 1.f - ay.t, wy1 = 1.f - wy0, wy2 = 1.f - wy1, wy3 = 1.f - wy2;
        const float wz0 = 1.f - az.t, wz1 = 1.f - wz0, wz2 = 1.f - wz1, wz3 = 1.f - wz2;
        const float wr[4] = {wx0 * wy0, wx0 * wy1, wx1 * wy2, wx1 * wy3};
        wlo[0] = wr[0] * wz0; wlo[1] = wr[1] * wz2; wlo[2] = wr[2] * wz2; wlo[3] = wr[3] * wz2;
        whi[0] = wr[0] * wz1; whi[1] = wr[1] * wz3; whi[2] = wr[2] * wz3; whi[3] = wr[3] * wz3;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const unsigned alo = rb[r] + az.i0, ahi = rb[r] + az.i1;
          const unsigned nxt = __shfl_down_sync(full, alo, 1);
          give[r] = (lane < 31) && (nxt == ahi) && (ahi != alo);
          took[r] = __shfl_up_sync(full, (int)give[r], 1) != 0 && lane > 0;
        }
      }
      const int dz = az.i1 - az.i0;
      const float t = ax.t, uu = ay.t, w = az.t;
      const float omt = 1.f - t, omu = 1.f - uu, omv = 1.f - w;
      for (int c = 0; c < C; ++c) {
        const float diff = __ldg(gn + (size_t)c * V + c0);
        if (NEED_I) {
          float* dc = dIn + (size_t)c * V;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float vlo = wlo[r] * diff;
            const float vhi = whi[r] * diff;
            const float recv = __shfl_up_sync(full, give[r] ? vhi : 0.f, 1);
            if (took[r]) vlo += recv;
            atomicAdd(dc + rb[r] + az.i0, vlo);
            if (!give[r]) atomicAdd(dc + rb[r] + az.i1, vhi);
          }
        }
        if (NEED_AT) {  // gradient of the interpolant (include/interp.h:315-326), g *= diff, outer product with f
          const float* Ic = In + (size_t)c * V;
          const float v0 = __ldg(Ic + rb[0] + az.i0), v4 = __ldg(Ic + rb[0] + az.i0 + dz);
          const float v3 = __ldg(Ic + rb[1] + az.i0), v7 = __ldg(Ic + rb[1] + az.i0 + dz);
          const float v1 = __ldg(Ic + rb[2] + az.i0), v5 = __ldg(Ic + rb[2] + az.i0 + dz);
          const float v2 = __ldg(Ic + rb[3] + az.i0), v6 = __ldg(Ic + rb[3] + az.i0 + dz);
          float gr[3];
          gr[0] = omv * (omu * (v1 - v0) + uu * (v2 - v3)) + w * (omu * (v5 - v4) + uu * (v6 - v7));
          gr[1] = omv * (omt * (v3 - v0) + t * (v2 - v1)) + w * (omt * (v7 - v4) + t * (v6 - v5));
          gr[2] = omu * (omt * (v4 - v0) + t * (v5 - v1)) + uu * (omt * (v7 - v3) + t * (v6 - v2));
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float gd = gr[r] * diff;  // "gx *= diff": cuda/affine.cu:273-274, :421-423
            part[r * 3 + 0] += gd * f0;
            part[r * 3 + 1] += gd * f1;
            part[r * 3 + 2] += gd * f2;
            part[9 + r] += gd;
          }
        }
      }
    }
  }
  if (NEED_AT) {
    __shared__ float red[8][12];
    const int wid = threadIdx.y;
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      const float s = warp_sum(part[q]);
      if (lane == 0) red[wid][q] = s;
    }
    __syncthreads();
    const int tid = threadIdx.y * 32 + lane;
    if (tid < 12) {
      float s = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) s += red[wq][tid];
      if (tid < 9) {
        if (d_A) atomicAdd(d_A + (size_t)n * 9 + tid, s);
      } else {
        if (d_T) atomicAdd(d_T + (size_t)n * 3 + (tid - 9), s);
      }
    }
  }
}

bool affine3_ok(int64_t N, int64_t C, const int64_t* sh) {
  if (sh[0] < 2 || sh[1] < 2 || sh[2] < 2 || C < 1) return false;
  if (sh[0] * sh[1] * sh[2] >= (1LL << 31) / 4 || C > 0x7fffffff / (sh[0] * sh[1] * sh[2])) return false;
  if (N * sh[0] > 65535 || sh[1] > 65535LL * 8) return false;
  return true;
}

}  // namespace

// LGM_EUNSUP when the fast path does not apply
int affine3_fwd_f32(void* out, const void* I, const void* A, const void* T, int64_t N, int64_t NI, int64_t C,
                    const int64_t* sh, cudaStream_t s) {
  if (!affine3_ok(N, C, sh)) return LGM_EUNSUP;
  const size_t ibs = (NI == 1 && N > 1) ? 0 : (size_t)C * sh[0] * sh[1] * sh[2];
  dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0])), block(32, 8);
#define LGM_AF(CC) \
  affine3_fwd_kernel<4, CC><<<grid, block, 0, s>>>((float*)out, (const float*)I, (const float*)A, (const float*)T, \
                                                   (int)sh[0], (int)sh[1], (int)sh[2], (int)C, ibs)
  if (C == 1) LGM_AF(1); else if (C == 3) LGM_AF(3); else LGM_AF(0);
#undef LGM_AF
  count_launch("affine_fwd", s);
  return finish(s, "lgm_affine_interp_fwd");
}

// d_I / d_A / d_T must be zero-filled by the caller (affine.cu does)
int affine3_bwd_f32(void* d_I, void* d_A, void* d_T, const void* go, const void* I, const void* A, const void* T,
                    int64_t N, int64_t NI, int64_t C, const int64_t* sh, cudaStream_t s) {
  if (!affine3_ok(N, C, sh) || sh[2] % 32 != 0) return LGM_EUNSUP;
  const size_t ibs = (NI == 1 && N > 1) ? 0 : (size_t)C * sh[0] * sh[1] * sh[2];
  const dim3 block(32, 8);
  const bool need_at = d_A || d_T;
  // The splat (d_I) and the pose gradients (d_A, d_T) run as TWO kernels: fused they need 104 registers
  // (2 CTAs / SM) and take longer than the pair (measured at 16 x 192^3: 2.9 ms fused, 1.2 + 1.1 apart).
  if (d_I) {
    dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * sh[0]));
    affine3_bwd_kernel<4, true, false, 1><<<grid, block, 0, s>>>((float*)d_I, nullptr, nullptr, (const float*)go,
        (const float*)I, (const float*)A, (const float*)T, (int)sh[0], (int)sh[1], (int)sh[2], (int)C, ibs);
    count_launch("affine_bwd_splat", s);
  }
  if (need_at) {
    constexpr int XS = 8;
    dim3 grid((unsigned)cdiv(sh[2], 128), (unsigned)cdiv(sh[1], 8), (unsigned)(N * cdiv(sh[0], XS)));
    affine3_bwd_kernel<4, false, true, XS><<<grid, block, 0, s>>>(nullptr, (float*)d_A, (float*)d_T, (const float*)go,
        (const float*)I, (const float*)A, (const float*)T, (int)sh[0], (int)sh[1], (int)sh[2], (int)C, ibs);
    count_launch("affine_bwd_pose", s);
  }
  return finish(s, "lgm_affine_interp_bwd");
}

}  // namespace lgm
