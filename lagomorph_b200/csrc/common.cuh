// common.cuh -- shared device helpers for liblagomorph_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lagomorph_b200.h"

namespace lgm {

// ---- host-side error plumbing (capi.cu) --------------------------------------
int set_error(int code, const char* fmt, ...);
int finish(cudaStream_t s, const char* what);  // checks launch error (+sync in debug mode)
void count_launch(const char* name, cudaStream_t s);  // counts; in profile mode also timestamps
bool debug_mode();

#define LGM_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return lgm::set_error(LGM_EINVAL, __VA_ARGS__); \
  } while (0)

// Spatial geometry of one tensor: extents and element strides, last axis fastest.
// One (sample, channel) volume is limited to 2^31-1 voxels so that all in-volume index
// arithmetic is 32-bit (batch / channel offsets stay 64-bit).
template <int D>
struct Geom {
  int n[D];
  int st[D];
  long long V;
};

template <int D>
inline Geom<D> make_geom(const int64_t* shape) {
  Geom<D> g;
  long long s = 1;
  for (int a = D - 1; a >= 0; --a) {
    g.n[a] = (int)shape[a];
    g.st[a] = (int)s;
    s *= shape[a];
  }
  g.V = s;
  return g;
}

template <int D>
inline bool geom_fits(const int64_t* shape) {
  long long s = 1;
  for (int a = 0; a < D; ++a) {
    if (shape[a] < 0 || shape[a] > 0x7fffffff) return false;
    s *= shape[a];
    if (s > 0x7fffffffLL) return false;
  }
  return true;
}

__host__ __device__ inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// ---- per-point interpolation -----------------------------------------------------
// One axis of a clamped linear interpolation: floor toward -inf, ceil = floor+1,
// weight from the UNclamped coordinate, both indices clamped independently into
// [0, n-1] (reference: include/interp.h:64-92, include/extrap.h:46-57).
template <typename R>
struct Axis {
  int i0, i1;  // clamped corner indices
  R t;         // weight of the upper corner
};

__device__ __forceinline__ int clampi(int r, int n) { return min(max(r, 0), n - 1); }

__device__ __forceinline__ Axis<float> axis_setup(float x, int n) {
  Axis<float> a;
  float fl = floorf(x);
  int f = __float2int_rd(x);  // saturating; equals the reference's trunc-then-decrement
  a.t = x - fl;
  a.i0 = clampi(f, n);
  a.i1 = clampi(f < 0x7fffffff ? f + 1 : f, n);
  return a;
}
__device__ __forceinline__ Axis<double> axis_setup(double x, int n) {
  Axis<double> a;
  double fl = floor(x);
  int f = __double2int_rd(x);
  a.t = x - fl;
  a.i0 = clampi(f, n);
  a.i1 = clampi(f < 0x7fffffff ? f + 1 : f, n);
  return a;
}

// Displaced sample coordinate: formed in double and rounded to Real exactly like
// the reference (dt is a double there: cuda/interp.cu:36-37, :68-73).
template <typename R>
__device__ __forceinline__ R coord(int i, double dt, R u) {
  return (R)((double)i + dt * (double)u);
}

// Corner gather + nested lerp in the reference's evaluation order.
template <typename R>
__device__ __forceinline__ R lerp2(const R* __restrict__ img, const Axis<R>& ax, const Axis<R>& ay,
                                   long long sx) {
  // include/interp.h:36-55: omt*(omu*v0 + u*v3) + t*(omu*v1 + u*v2)
  const R* r0 = img + ax.i0 * sx;
  const R* r1 = img + ax.i1 * sx;
  R v0 = __ldg(r0 + ay.i0), v3 = __ldg(r0 + ay.i1);
  R v1 = __ldg(r1 + ay.i0), v2 = __ldg(r1 + ay.i1);
  R omt = R(1) - ax.t, omu = R(1) - ay.t;
  return omt * (omu * v0 + ay.t * v3) + ax.t * (omu * v1 + ay.t * v2);
}

template <typename R>
struct Corners3 {
  R v0, v1, v2, v3, v4, v5, v6, v7;  // reference numbering, include/interp.h:91-98
};

template <typename R>
__device__ __forceinline__ Corners3<R> gather3(const R* __restrict__ img, const Axis<R>& ax,
                                               const Axis<R>& ay, const Axis<R>& az, long long sx,
                                               long long sy) {
  const R* p00 = img + ax.i0 * sx + ay.i0 * sy;
  const R* p10 = img + ax.i1 * sx + ay.i0 * sy;
  const R* p11 = img + ax.i1 * sx + ay.i1 * sy;
  const R* p01 = img + ax.i0 * sx + ay.i1 * sy;
  Corners3<R> c;
  c.v0 = __ldg(p00 + az.i0); c.v1 = __ldg(p10 + az.i0);
  c.v2 = __ldg(p11 + az.i0); c.v3 = __ldg(p01 + az.i0);
  c.v4 = __ldg(p00 + az.i1); c.v5 = __ldg(p10 + az.i1);
  c.v6 = __ldg(p11 + az.i1); c.v7 = __ldg(p01 + az.i1);
  return c;
}

template <typename R>
__device__ __forceinline__ R lerp3_eval(const Corners3<R>& c, R t, R u, R v) {
  // include/interp.h:115-122
  R omt = R(1) - t, omu = R(1) - u, omv = R(1) - v;
  return omv * (omu * (omt * c.v0 + t * c.v1) + u * (omt * c.v3 + t * c.v2)) +
         v * (omu * (omt * c.v4 + t * c.v5) + u * (omt * c.v7 + t * c.v6));
}

template <typename R>
__device__ __forceinline__ void lerp3_grad(const Corners3<R>& c, R t, R u, R v, R& gx, R& gy,
                                           R& gz) {
  // include/interp.h:315-326
  R omt = R(1) - t, omu = R(1) - u, omv = R(1) - v;
  gx = omv * (omu * (c.v1 - c.v0) + u * (c.v2 - c.v3)) + v * (omu * (c.v5 - c.v4) + u * (c.v6 - c.v7));
  gy = omv * (omt * (c.v3 - c.v0) + t * (c.v2 - c.v1)) + v * (omt * (c.v7 - c.v4) + t * (c.v6 - c.v5));
  gz = omu * (omt * (c.v4 - c.v0) + t * (c.v5 - c.v1)) + u * (omt * (c.v7 - c.v3) + t * (c.v6 - c.v2));
}

// float/double atomic add without return value (RED)
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }

// Decode a linear voxel id into per-axis positions.
template <int D>
__device__ __forceinline__ void decode(long long vid64, const Geom<D>& g, int (&pos)[D]) {
  unsigned vid = (unsigned)vid64;  // < 2^31 by construction (geom_fits)
#pragma unroll
  for (int a = D - 1; a > 0; --a) {
    unsigned q = vid / (unsigned)g.n[a];
    pos[a] = (int)(vid - q * (unsigned)g.n[a]);
    vid = q;
  }
  pos[0] = (int)vid;
}

}  // namespace lgm
