// ref_points.cpp -- TEST INFRASTRUCTURE ONLY (oracle pin).
//
// Instantiates the REFERENCE's own per-point template headers on the host
// (include/interp.h, include/extrap.h, include/diff.h -- they are plain C++
// once DEVICE/__device__ are empty, see include/defs.h:44-48) and exports them
// with a C ABI so tests can pin oracle/lgm_oracle.cpp against them point by
// point. Nothing from the reference is copied: the headers are #included from
// where they lie under /root/reference at build time (oracle/Makefile) and the
// resulting binary goes to oracle/_ref/ (git-ignored).
#include <cstddef>
#include <cstdio>
// host stand-in for CUDA's atomicAdd, same trick as cpu/affine.cpp:5 in the reference
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p += v; return o; }
#define __device__
#include "interp.h"
#include "diff.h"

bool lagomorph_debug_mode = false;
static const BackgroundStrategy BG = BACKGROUND_STRATEGY_CLAMP;

#define BOTH(NAME, R)                                                                          \
  R ref_bilerp_##NAME(const R* img, R x, R y, int nx, int ny) {                                \
    return biLerp<R, BG>(img, x, y, nx, ny); }                                                 \
  R ref_trilerp_##NAME(const R* img, R x, R y, R z, int nx, int ny, int nz) {                  \
    return triLerp<R, BG>(img, x, y, z, nx, ny, nz); }                                         \
  void ref_bilerp_grad_##NAME(R* o, const R* img, R x, R y, int nx, int ny) {                  \
    biLerp_grad<R, BG>(o[0], o[1], o[2], img, x, y, nx, ny); }                                 \
  void ref_trilerp_grad_##NAME(R* o, const R* img, R x, R y, R z, int nx, int ny, int nz) {    \
    triLerp_grad<R, BG>(o[0], o[1], o[2], o[3], img, x, y, z, nx, ny, nz); }                   \
  void ref_splat2_##NAME(R* d, R mass, R x, R y, int nx, int ny) {                             \
    atomicSplat<R, BG, false>(d, (R*)NULL, mass, x, y, nx, ny); }                              \
  void ref_splat3_##NAME(R* d, R mass, R x, R y, R z, int nx, int ny, int nz) {                \
    atomicSplat<R, BG, false>(d, (R*)NULL, mass, x, y, z, nx, ny, nz); }                       \
  void ref_grad2_##NAME(R* o, const R* a, int nx, int ny, int i, int j) {                      \
    grad_point<R, BG>(o[0], o[1], a, nx, ny, i, j); }                                          \
  void ref_grad3_##NAME(R* o, const R* a, int nx, int ny, int nz, int i, int j, int k) {       \
    grad_point<R, BG>(o[0], o[1], o[2], a, nx, ny, nz, i, j, k); }

extern "C" {
BOTH(f32, float)
BOTH(f64, double)
}
