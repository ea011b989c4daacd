// lgm_oracle.cpp -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the reference's (jacobhinkle/lagomorph) LDDMM hot-path
// kernels, written to follow the reference's arithmetic (evaluation order,
// where doubles are formed and rounded, clamp rules) so that the CUDA product
// in lagomorph_b200/csrc can be checked against it.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library. The product path never does.
//
// Parity pin: the per-point arithmetic below is cross-checked against the
// reference's own headers (include/interp.h, extrap.h, diff.h instantiated on
// the host by oracle/ref_points.cpp -> oracle/_ref/libref_points.so) and the
// whole-kernel results against the reference's own CUDA kernels compiled for
// sm_100a (oracle/ref_cuda -> oracle/_ref/libref_cuda.so) on the GPU box; the
// outputs of that run are committed as tests/golden/*.npz.
//
// All citations are relative to /root/reference/lagomorph/extension/.
// Layout everywhere: contiguous N C X Y (Z), channel d = component along axis d.
//
// Build: see oracle/Makefile (g++ -O2 -fopenmp -ffp-contract=off).

#include <cmath>
#include <vector>
#include <cstddef>
#include <cstdint>

namespace {

// ---- per-point helpers -----------------------------------------------------

// floor toward -inf exactly as include/interp.h:64-70
template <typename R>
inline int floor_ref(R x) {
  int f = (int)(x);
  if (x < 0 && x != (R)(int)(x)) --f;
  return f;
}

// include/extrap.h:41-44
inline int clampi(int r, int b) {
  if (r < 0) return 0;
  if (r >= b) return b - 1;
  return r;
}

// include/interp.h:9-56 (biLerp, CLAMP only)
template <typename R>
inline R bilerp(const R* img, R x, R y, int nx, int ny) {
  int fx = floor_ref(x), fy = floor_ref(y);
  int cx = fx + 1, cy = fy + 1;
  R t = x - fx, u = y - fy;
  R omt = R(1) - t, omu = R(1) - u;
  fx = clampi(fx, nx); cx = clampi(cx, nx);  // extrap.h:46-57 == independent clamps
  fy = clampi(fy, ny); cy = clampi(cy, ny);
  R v0 = img[fx * ny + fy];
  R v1 = img[cx * ny + fy];
  R v2 = img[cx * ny + cy];
  R v3 = img[fx * ny + cy];
  return omt * (omu * v0 + u * v3) + t * (omu * v1 + u * v2);
}

// include/interp.h:59-123 (triLerp)
template <typename R>
inline R trilerp(const R* img, R x, R y, R z, int nx, int ny, int nz) {
  int fx = floor_ref(x), fy = floor_ref(y), fz = floor_ref(z);
  int cx = fx + 1, cy = fy + 1, cz = fz + 1;
  R t = x - fx, u = y - fy, v = z - fz;
  R omt = R(1) - t, omu = R(1) - u, omv = R(1) - v;
  fx = clampi(fx, nx); cx = clampi(cx, nx);
  fy = clampi(fy, ny); cy = clampi(cy, ny);
  fz = clampi(fz, nz); cz = clampi(cz, nz);
#define AT3(a, b, c) img[((size_t)(a) * ny + (b)) * nz + (c)]
  R v0 = AT3(fx, fy, fz), v1 = AT3(cx, fy, fz), v2 = AT3(cx, cy, fz), v3 = AT3(fx, cy, fz);
  R v4 = AT3(fx, fy, cz), v5 = AT3(cx, fy, cz), v6 = AT3(cx, cy, cz), v7 = AT3(fx, cy, cz);
#undef AT3
  return omv * (omu * (omt * v0 + t * v1) + u * (omt * v3 + t * v2)) +
         v * (omu * (omt * v4 + t * v5) + u * (omt * v7 + t * v6));
}

// include/interp.h:128-204 (biLerp_grad, CLAMP => always "inside")
template <typename R>
inline void bilerp_grad(R& gx, R& gy, const R* img, R x, R y, int nx, int ny) {
  int fx = floor_ref(x), fy = floor_ref(y);
  int cx = fx + 1, cy = fy + 1;
  R t = x - fx, u = y - fy;
  fx = clampi(fx, nx); cx = clampi(cx, nx);
  fy = clampi(fy, ny); cy = clampi(cy, ny);
  R v0 = img[fx * ny + fy];
  R v1 = img[cx * ny + fy];
  R v2 = img[cx * ny + cy];
  R v3 = img[fx * ny + cy];
  gx = v1 - v0 + u * (v2 - v3 - v1 + v0);
  gy = v3 - v0 + t * (v2 - v1 - v3 + v0);
}

// include/interp.h:206-327 (triLerp_grad)
template <typename R>
inline void trilerp_grad(R& gx, R& gy, R& gz, const R* img, R x, R y, R z, int nx, int ny,
                         int nz) {
  int fx = floor_ref(x), fy = floor_ref(y), fz = floor_ref(z);
  int cx = fx + 1, cy = fy + 1, cz = fz + 1;
  R t = x - fx, u = y - fy, v = z - fz;
  R omt = R(1) - t, omu = R(1) - u, omv = R(1) - v;
  fx = clampi(fx, nx); cx = clampi(cx, nx);
  fy = clampi(fy, ny); cy = clampi(cy, ny);
  fz = clampi(fz, nz); cz = clampi(cz, nz);
#define AT3(a, b, c) img[((size_t)(a) * ny + (b)) * nz + (c)]
  R v0 = AT3(fx, fy, fz), v1 = AT3(cx, fy, fz), v2 = AT3(cx, cy, fz), v3 = AT3(fx, cy, fz);
  R v4 = AT3(fx, fy, cz), v5 = AT3(cx, fy, cz), v6 = AT3(cx, cy, cz), v7 = AT3(fx, cy, cz);
#undef AT3
  gx = omv * (omu * (v1 - v0) + u * (v2 - v3)) + v * (omu * (v5 - v4) + u * (v6 - v7));
  gy = omv * (omt * (v3 - v0) + t * (v2 - v1)) + v * (omt * (v7 - v4) + t * (v6 - v5));
  gz = omu * (omt * (v4 - v0) + t * (v5 - v1)) + u * (omt * (v7 - v3) + t * (v6 - v2));
}

template <typename R>
inline void atomic_add(R* p, R v) {
#pragma omp atomic
  *p += v;
}

// include/interp.h:403-426 (atomicSplat 2-D) + :330-364 (splat_neighbor, CLAMP)
template <typename R>
inline void splat2(R* d, R mass, R x, R y, int nx, int ny) {
  int xi0 = floor_ref(x), yi0 = floor_ref(y);
  R dx = R(1) - (x - (R)xi0);
  R dy = R(1) - (y - (R)yi0);
  for (int xi = xi0; xi < xi0 + 2; xi++) {
    for (int yi = yi0; yi < yi0 + 2; yi++) {
      R ww = dx * dy;
      atomic_add(&d[clampi(xi, nx) * ny + clampi(yi, ny)], (R)(ww * mass));
      dy = R(1) - dy;
    }
    dx = R(1) - dx;
  }
}

// include/interp.h:427-454 (atomicSplat 3-D) + :366-401
template <typename R>
inline void splat3(R* d, R mass, R x, R y, R z, int nx, int ny, int nz) {
  int xi0 = floor_ref(x), yi0 = floor_ref(y), zi0 = floor_ref(z);
  R dx = R(1) - (x - xi0);
  R dy = R(1) - (y - yi0);
  R dz = R(1) - (z - zi0);
  for (int xi = xi0; xi < xi0 + 2; xi++) {
    for (int yi = yi0; yi < yi0 + 2; yi++) {
      for (int zi = zi0; zi < zi0 + 2; zi++) {
        R ww = dx * dy * dz;
        atomic_add(&d[((size_t)clampi(xi, nx) * ny + clampi(yi, ny)) * nz + clampi(zi, nz)],
                   (R)(ww * mass));
        dz = R(1) - dz;
      }
      dy = R(1) - dy;
    }
    dx = R(1) - dx;
  }
}

// include/diff.h:6-76 with include/extrap.h:110-157 (get_value_safe, CLAMP)
template <typename R>
inline void grad2(R& gx, R& gy, const R* a, int nx, int ny, int i, int j) {
  gx = 0.5f * (a[clampi(i + 1, nx) * ny + j] - a[clampi(i - 1, nx) * ny + j]);
  gy = 0.5f * (a[i * ny + clampi(j + 1, ny)] - a[i * ny + clampi(j - 1, ny)]);
}
template <typename R>
inline void grad3(R& gx, R& gy, R& gz, const R* a, int nx, int ny, int nz, int i, int j, int k) {
#define AT3(p, q, r) a[((size_t)(p) * ny + (q)) * nz + (r)]
  gx = 0.5f * (AT3(clampi(i + 1, nx), j, k) - AT3(clampi(i - 1, nx), j, k));
  gy = 0.5f * (AT3(i, clampi(j + 1, ny), k) - AT3(i, clampi(j - 1, ny), k));
  gz = 0.5f * (AT3(i, j, clampi(k + 1, nz)) - AT3(i, j, clampi(k - 1, nz)));
#undef AT3
}

// Exact transpose of the clamped central difference along one axis applied to
// the product a*b, evaluated at position p of an axis of length n with element
// stride s: cuda/diff.cu:432-460 (and every "if (i == 0) ... else if (i == nx-1)").
// The reference multiplies by the double literal -.5/.5 and accumulates with +=
// (so the add is done in double and rounded to Real): reproduce that.
template <typename R>
inline void dT_acc(R& acc, const R* a, const R* b, long idx, long s, int p, int n) {
  if (p == 0)
    acc += -.5 * (a[idx] * b[idx] + a[idx + s] * b[idx + s]);
  else if (p == n - 1)
    acc += .5 * (a[idx] * b[idx] + a[idx - s] * b[idx - s]);
  else
    acc += -.5 * (a[idx + s] * b[idx + s] - a[idx - s] * b[idx - s]);
}

// ---- interp ----------------------------------------------------------------

// cuda/interp.cu:15-78 (forward kernels), :80-130 (host: batch/broadcast rule)
template <typename R>
void interp_fwd(R* out, const R* I, const R* u, long N, long NI, long C, int dim, const long* sh,
                double dt) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const long V = (long)nx * ny * nz;
  const bool bcast = NI < N;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      R fi = (R)i, fj = (R)j;
      for (long n = 0; n < N; ++n) {
        const R* In = I + (bcast ? 0 : n * C * V);
        const R* un = u + n * dim * V;
        for (long c = 0; c < C; ++c) {
          R* o = out + (n * C + c) * V;
          if (dim == 2) {
            long ix = (long)i * ny + j;
            R hx = (R)(fi + dt * un[ix]);  // double, rounded at the call: interp.cu:36-40
            R hy = (R)(fj + dt * un[ix + V]);
            o[ix] = bilerp<R>(In + c * V, hx, hy, nx, ny);
          } else {
            for (int k = 0; k < nz; ++k) {
              long ix = ((long)i * ny + j) * nz + k;
              R fk = (R)k;
              R hx = (R)(fi + dt * un[ix]);  // interp.cu:68-73
              R hy = (R)(fj + dt * un[ix + V]);
              R hz = (R)(fk + dt * un[ix + 2 * V]);
              o[ix] = trilerp<R>(In + c * V, hx, hy, hz, nx, ny, nz);
            }
          }
        }
      }
    }
}

// cuda/interp.cu:132-244 (backward kernels), :246-313 (host). d_I, d_u are
// zero-filled here like the reference's at::zeros_like (:263-264).
template <typename R>
void interp_bwd(R* d_I, R* d_u, const R* go, const R* I, const R* u, long N, long NI, long C,
                int dim, const long* sh, double dt, int need_I, int need_u) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const long V = (long)nx * ny * nz;
  const bool bcast = NI < N;
  for (long q = 0; q < NI * C * V; ++q) d_I[q] = 0;
  for (long q = 0; q < N * dim * V; ++q) d_u[q] = 0;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      for (long n = 0; n < N; ++n) {
        const R* un = u + n * dim * V;
        R* dun = d_u + n * dim * V;
        for (long c = 0; c < C; ++c) {
          const R* In = I + ((bcast ? 0 : n * C) + c) * V;
          R* dIn = d_I + ((bcast ? 0 : n * C) + c) * V;
          const R* gon = go + (n * C + c) * V;
          if (dim == 2) {
            long ix = (long)i * ny + j;
            R hx = (R)(i + dt * un[ix]);  // interp.cu:160-161
            R hy = (R)(j + dt * un[ix + V]);
            R diff = gon[ix];
            if (need_I) splat2<R>(dIn, diff, hx, hy, nx, ny);
            if (need_u) {
              R gx, gy;
              bilerp_grad<R>(gx, gy, In, hx, hy, nx, ny);
              diff = (R)(diff * dt);  // "diff *= dt" with double dt: interp.cu:169
              dun[ix] = dun[ix] + gx * diff;
              dun[ix + V] = dun[ix + V] + gy * diff;
            }
          } else {
            for (int k = 0; k < nz; ++k) {
              long ix = ((long)i * ny + j) * nz + k;
              R hx = (R)(i + dt * un[ix]);  // interp.cu:216-218
              R hy = (R)(j + dt * un[ix + V]);
              R hz = (R)(k + dt * un[ix + 2 * V]);
              R diff = gon[ix];
              if (need_I) splat3<R>(dIn, diff, hx, hy, hz, nx, ny, nz);
              if (need_u) {
                R gx, gy, gz;
                trilerp_grad<R>(gx, gy, gz, In, hx, hy, hz, nx, ny, nz);
                diff = (R)(diff * dt);  // interp.cu:230
                dun[ix] = dun[ix] + gx * diff;
                dun[ix + V] = dun[ix + V] + gy * diff;
                dun[ix + 2 * V] = dun[ix + 2 * V] + gz * diff;
              }
            }
          }
        }
      }
    }
}

// ---- jacobian_times_vectorfield -------------------------------------------

// cuda/diff.cu:17-127 (forward kernels), :129-185 (host)
template <typename R>
void jtvf_fwd(R* out, const R* v, const R* w, long N, long C, int dim, const long* sh, int disp,
              int trans) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const long V = (long)nx * ny * nz;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      for (long n = 0; n < N; ++n) {
        const R* vn = v + n * C * V;
        const R* wn = w + n * dim * V;
        R* on = out + n * C * V;
        for (int k = 0; k < nz; ++k) {
          long ix = ((long)i * ny + j) * nz + k;
          if (trans) {  // diff.cu:34-44 (2-D), :81-102 (3-D); C == dim
            for (int c = 0; c < dim; ++c) {
              R g[3] = {0, 0, 0};
              if (dim == 2) grad2<R>(g[0], g[1], vn + c * V, nx, ny, i, j);
              else grad3<R>(g[0], g[1], g[2], vn + c * V, nx, ny, nz, i, j, k);
              if (disp) g[c] += 1.0;
              for (int d = 0; d < dim; ++d) {
                if (c == 0) on[ix + d * V] = g[d] * wn[ix + c * V];
                else on[ix + d * V] += g[d] * wn[ix + c * V];
              }
            }
          } else {  // diff.cu:46-56, :104-122
            for (long c = 0; c < C; ++c) {
              R g[3] = {0, 0, 0};
              if (dim == 2) grad2<R>(g[0], g[1], vn + c * V, nx, ny, i, j);
              else grad3<R>(g[0], g[1], g[2], vn + c * V, nx, ny, nz, i, j, k);
              if (disp && c < dim) g[c] += 1.0;
              if (dim == 2) on[ix + c * V] = g[0] * wn[ix] + g[1] * wn[ix + V];
              else on[ix + c * V] = g[0] * wn[ix] + g[1] * wn[ix + V] + g[2] * wn[ix + 2 * V];
            }
          }
        }
      }
}

// cuda/diff.cu:187-473 (backward kernels), :475-540 (host forces both grads)
template <typename R>
void jtvf_bwd(R* d_v, R* d_w, const R* go, const R* v, const R* w, long N, long C, int dim,
              const long* sh, int disp, int trans) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const long V = (long)nx * ny * nz;
  const long sx = (long)ny * nz, sy = nz, sz = 1;
  const long strides[3] = {sx, sy, sz};
  const int dims_[3] = {nx, ny, nz};
  for (long q = 0; q < N * C * V; ++q) d_v[q] = 0;
  for (long q = 0; q < N * dim * V; ++q) d_w[q] = 0;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      for (long n = 0; n < N; ++n) {
        const R* vn = v + n * C * V;
        const R* wn = w + n * dim * V;
        const R* gon = go + n * C * V;
        R* dvn = d_v + n * C * V;
        R* dwn = d_w + n * dim * V;
        for (int k = 0; k < nz; ++k) {
          long ix = ((long)i * ny + j) * nz + k;
          const int pos[3] = {i, j, k};
          if (trans) {
            // d_w[c] = sum_d (D_d v_c + delta) gout_d : diff.cu:210-220, :311-335
            for (int c = 0; c < dim; ++c) {
              R g[3] = {0, 0, 0};
              if (dim == 2) grad2<R>(g[0], g[1], vn + c * V, nx, ny, i, j);
              else grad3<R>(g[0], g[1], g[2], vn + c * V, nx, ny, nz, i, j, k);
              if (disp) g[c] += 1.0;
              if (dim == 2) dwn[ix + c * V] += g[0] * gon[ix] + g[1] * gon[ix + V];
              else dwn[ix + c * V] += g[0] * gon[ix] + g[1] * gon[ix + V] + g[2] * gon[ix + 2 * V];
            }
            // d_v[c] += sum_d D_d^T (w_c gout_d) : diff.cu:221-249, :336-407
            for (int d = 0; d < dim; ++d)
              for (int c = 0; c < dim; ++c)
                dT_acc<R>(dvn[ix + c * V], wn + c * V, gon + d * V, ix, strides[d],
                          pos[d], dims_[d]);
          } else {
            for (long c = 0; c < C; ++c) {
              R g[3] = {0, 0, 0};
              if (dim == 2) grad2<R>(g[0], g[1], vn + c * V, nx, ny, i, j);
              else grad3<R>(g[0], g[1], g[2], vn + c * V, nx, ny, nz, i, j, k);
              if (disp && c < dim) g[c] += 1.0;
              // diff.cu:254-262, :417-431
              for (int d = 0; d < dim; ++d) dwn[ix + d * V] += g[d] * gon[ix + c * V];
              // d_v[c] += sum_d D_d^T (w_d gout_c) : diff.cu:263-279, :432-460
              for (int d = 0; d < dim; ++d)
                dT_acc<R>(dvn[ix + c * V], wn + d * V, gon + c * V, ix, strides[d], pos[d], dims_[d]);
            }
          }
        }
      }
}

// cuda/diff.cu:546-632 (adjoint forward), :634-672 (host; out zero-filled)
template <typename R>
void jtvf_adj_fwd(R* out, const R* z, const R* w, long N, long C, int dim, const long* sh) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const long V = (long)nx * ny * nz;
  const long strides[3] = {(long)ny * nz, (long)nz, 1};
  const int dims_[3] = {nx, ny, nz};
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      for (long n = 0; n < N; ++n)
        for (long c = 0; c < C; ++c) {
          const R* zn = z + (n * C + c) * V;
          const R* wn = w + n * dim * V;
          R* on = out + (n * C + c) * V;
          for (int k = 0; k < nz; ++k) {
            long ix = ((long)i * ny + j) * nz + k;
            const int pos[3] = {i, j, k};
            R acc = 0;
            for (int d = 0; d < dim; ++d)
              dT_acc<R>(acc, wn + d * V, zn, ix, strides[d], pos[d], dims_[d]);
            on[ix] = acc;
          }
        }
}

// cuda/diff.cu:674-780 (adjoint backward), :783-835 (host; C == dim assumed)
template <typename R>
void jtvf_adj_bwd(R* d_z, R* d_w, const R* go, const R* z, const R* w, long N, long C, int dim,
                  const long* sh) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const long V = (long)nx * ny * nz;
  (void)C;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      for (long n = 0; n < N; ++n) {
        const R* zn = z + n * dim * V;
        const R* wn = w + n * dim * V;
        const R* gon = go + n * dim * V;
        R* dzn = d_z + n * dim * V;
        R* dwn = d_w + n * dim * V;
        for (int k = 0; k < nz; ++k) {
          long ix = ((long)i * ny + j) * nz + k;
          for (int c = 0; c < dim; ++c) {
            R g[3] = {0, 0, 0};
            if (dim == 2) grad2<R>(g[0], g[1], gon + c * V, nx, ny, i, j);
            else grad3<R>(g[0], g[1], g[2], gon + c * V, nx, ny, nz, i, j, k);
            for (int d = 0; d < dim; ++d) {
              if (c == 0) dwn[ix + d * V] = g[d] * zn[ix + c * V];
              else dwn[ix + d * V] += g[d] * zn[ix + c * V];
            }
            R s = (dim == 2) ? (g[0] * wn[ix] + g[1] * wn[ix + V])
                             : (g[0] * wn[ix] + g[1] * wn[ix + V] + g[2] * wn[ix + 2 * V]);
            dzn[ix + c * V] = R(0) + s;  // "+=" onto a zero-filled tensor
          }
        }
      }
}

// ---- fluid operator (Fourier multiplier) ------------------------------------

// cuda/metric.cu:14-18
template <typename R>
inline R safe_sqrt(R x) {
  if (x < 1e-8) return (R)1e-4;
  return std::sqrt(x);
}

// cuda/metric.cu:162-218 (2-D) and :220-306 (3-D). Fm is the interleaved
// half-spectrum (N, dim, X, Y[, Zc], 2) modified in place. LUT element type is
// Real (metric.py:65-75 rounds the float64 tables to the tensor dtype).
template <typename R>
void fluid_op(R* Fm, int inverse, const R* cosX, const R* sinX, const R* cosY, const R* sinY,
              const R* cosZ, const R* sinZ, double alpha, double beta, double gamma, long N,
              int dim, const long* sh) {
  const long nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
#pragma omp parallel for collapse(2) schedule(static)
  for (long i = 0; i < nx; ++i)
    for (long j = 0; j < ny; ++j) {
      const R wx = cosX[i], wy = cosY[j];
      if (dim == 2) {
        const long nxy = 2 * nx * ny;
        long ix = 2 * (j + i * ny), iy = ix + nxy;
        const R lambda = (R)(gamma + alpha * (wx + wy));
        R l00 = (R)(lambda - beta * wx);
        R l11 = (R)(lambda - beta * wy);
        R l10 = (R)(beta * sinX[i] * sinY[j]);
        R L00 = l00 * l00 + l10 * l10;
        R L10 = l00 * l10 + l10 * l11;
        R L11 = l11 * l11 + l10 * l10;
        R ooG00 = 0, G10 = 0, ooG11 = 0;
        if (inverse) {  // metric.cu:20-45
          ooG00 = (R)(1. / safe_sqrt(L00));
          G10 = L10 * ooG00;
          ooG11 = L11 - G10 * G10;
          ooG11 = (R)(1. / safe_sqrt(ooG11));
        }
        for (long n = 0; n < N; ++n, ix += 2 * nxy, iy += 2 * nxy)
          for (int p = 0; p < 2; ++p) {  // real part, then imaginary part
            R bX = Fm[ix + p], bY = Fm[iy + p];
            if (inverse) {  // metric.cu:80-100
              R y0 = bX * ooG00;
              R y1 = (bY - G10 * y0) * ooG11;
              bY = y1 * ooG11;
              bX = (y0 - G10 * bY) * ooG00;
            } else {  // metric.cu:132-144
              R x = L00 * bX + L10 * bY;
              bY = L10 * bX + L11 * bY;
              bX = x;
            }
            Fm[ix + p] = bX;
            Fm[iy + p] = bY;
          }
      } else {
        const long nxyz = 2 * nx * ny * nz;
        for (long k = 0; k < nz; ++k) {
          const R wz = cosZ[k];
          const R lambda = (R)(gamma + alpha * (wx + wy + wz));
          R l00 = (R)(lambda - beta * wx);
          R l11 = (R)(lambda - beta * wy);
          R l22 = (R)(lambda - beta * wz);
          R l10 = (R)(beta * sinX[i] * sinY[j]);
          R l20 = (R)(beta * sinX[i] * sinZ[k]);
          R l21 = (R)(beta * sinY[j] * sinZ[k]);
          R L00 = l00 * l00 + l10 * l10 + l20 * l20;
          R L10 = l00 * l10 + l10 * l11 + l20 * l21;
          R L11 = l10 * l10 + l11 * l11 + l21 * l21;
          R L20 = l00 * l20 + l10 * l21 + l20 * l22;
          R L21 = l10 * l20 + l11 * l21 + l21 * l22;
          R L22 = l20 * l20 + l21 * l21 + l22 * l22;
          R ooG00 = 0, G10 = 0, ooG11 = 0, G20 = 0, G21 = 0, ooG22 = 0;
          if (inverse) {  // metric.cu:47-78
            ooG00 = (R)(1. / safe_sqrt(L00));
            G10 = L10 * ooG00;
            G20 = L20 * ooG00;
            ooG11 = L11 - G10 * G10;
            ooG11 = (R)(1. / safe_sqrt(ooG11));
            G21 = (L21 - G20 * G10) * ooG11;
            ooG22 = L22 - G20 * G20 - G21 * G21;
            ooG22 = (R)(1. / safe_sqrt(ooG22));
          }
          long ix = 2 * (j + i * ny) * nz + 2 * k, iy = ix + nxyz, iz = iy + nxyz;
          for (long n = 0; n < N; ++n, ix += 3 * nxyz, iy += 3 * nxyz, iz += 3 * nxyz)
            for (int p = 0; p < 2; ++p) {
              R bX = Fm[ix + p], bY = Fm[iy + p], bZ = Fm[iz + p];
              if (inverse) {  // metric.cu:102-130
                R y0 = bX * ooG00;
                R y1 = (bY - G10 * y0) * ooG11;
                R y2 = (bZ - G20 * y0 - G21 * y1) * ooG22;
                bZ = y2 * ooG22;
                bY = (y1 - G21 * bZ) * ooG11;
                bX = (y0 - G10 * bY - G20 * bZ) * ooG00;
              } else {  // metric.cu:146-160
                R x = L00 * bX + L10 * bY + L20 * bZ;
                R y = L10 * bX + L11 * bY + L21 * bZ;
                bZ = L20 * bX + L21 * bY + L22 * bZ;
                bX = x;
                bY = y;
              }
              Fm[ix + p] = bX;
              Fm[iy + p] = bY;
              Fm[iz + p] = bZ;
            }
        }
      }
    }
}

// ---- regrid ------------------------------------------------------------------

// cuda/affine.cu:612-681 (forward), :683-734 (host). osh = output shape.
template <typename R>
void regrid_fwd(R* out, const R* I, long N, long C, int dim, const long* sh, const long* osh,
                const double* origin, const double* spacing) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const int Nx = osh[0], Ny = osh[1], Nz = dim == 3 ? osh[2] : 1;
  const long V = (long)nx * ny * nz, W = (long)Nx * Ny * Nz;
  const R Ox = (R)origin[0], Oy = (R)origin[1], Oz = dim == 3 ? (R)origin[2] : R(0);
  const R Sx = (R)spacing[0], Sy = (R)spacing[1], Sz = dim == 3 ? (R)spacing[2] : R(0);
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < Nx; ++i)
    for (int j = 0; j < Ny; ++j) {
      R ox = (R)(.5 * static_cast<R>(Nx - 1));
      R oy = (R)(.5 * static_cast<R>(Ny - 1));
      R oz = (R)(.5 * static_cast<R>(Nz - 1));
      R hx = (i - ox) * Sx + Ox;
      R hy = (j - oy) * Sy + Oy;
      for (long nc = 0; nc < N * C; ++nc) {
        const R* In = I + nc * V;
        R* on = out + nc * W;
        if (dim == 2) {
          on[(long)i * Ny + j] = bilerp<R>(In, hx, hy, nx, ny);
        } else {
          R hz = Oz - oz * Sz;  // accumulated, affine.cu:669-675
          for (int k = 0; k < Nz; ++k) {
            on[((long)i * Ny + j) * Nz + k] = trilerp<R>(In, hx, hy, hz, nx, ny, nz);
            hz += Sz;
          }
        }
      }
    }
}

// cuda/affine.cu:736-800 (backward), :802-855 (host; d_I zero-filled)
template <typename R>
void regrid_bwd(R* d_I, const R* go, long N, long C, int dim, const long* sh, const long* osh,
                const double* origin, const double* spacing) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const int Nx = osh[0], Ny = osh[1], Nz = dim == 3 ? osh[2] : 1;
  const long V = (long)nx * ny * nz, W = (long)Nx * Ny * Nz;
  const R Ox = (R)origin[0], Oy = (R)origin[1], Oz = dim == 3 ? (R)origin[2] : R(0);
  const R Sx = (R)spacing[0], Sy = (R)spacing[1], Sz = dim == 3 ? (R)spacing[2] : R(0);
  for (long q = 0; q < N * C * V; ++q) d_I[q] = 0;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < Nx; ++i)
    for (int j = 0; j < Ny; ++j) {
      R ox = (R)(.5 * static_cast<R>(Nx - 1));
      R oy = (R)(.5 * static_cast<R>(Ny - 1));
      R oz = (R)(.5 * static_cast<R>(Nz - 1));
      R hx = (i - ox) * Sx + Ox;
      R hy = (j - oy) * Sy + Oy;
      for (long nc = 0; nc < N * C; ++nc) {
        const R* gon = go + nc * W;
        R* dIn = d_I + nc * V;
        if (dim == 2) {
          splat2<R>(dIn, gon[(long)i * Ny + j], hx, hy, nx, ny);
        } else {
          for (int k = 0; k < Nz; ++k) {
            R hz = (k - oz) * Sz + Oz;  // not accumulated, affine.cu:792
            splat3<R>(dIn, gon[((long)i * Ny + j) * Nz + k], hx, hy, hz, nx, ny, nz);
          }
        }
      }
    }
}

// ---- affine_interp forward ------------------------------------------------------
// cuda/affine.cu:23-112 (GPU forward kernels; the closed-form coordinate, not the
// CPU kernel's incremental one in cpu/affine.cpp:35-61).
template <typename R>
void affine_fwd(R* out, const R* I, const R* A, const R* T, long N, long NI, long C, int dim,
                const long* sh) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const long V = (long)nx * ny * nz;
  const bool bcast = (NI == 1 && N > 1);
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      R ox = (R)(.5 * static_cast<R>(nx - 1));
      R oy = (R)(.5 * static_cast<R>(ny - 1));
      R oz = (R)(.5 * static_cast<R>(nz - 1));
      R fi = static_cast<R>(i) - ox, fj = static_cast<R>(j) - oy;
      for (long n = 0; n < N; ++n) {
        const R* An = A + n * dim * dim;
        const R* Tn = T + n * dim;
        for (long c = 0; c < C; ++c) {
          const R* In = I + ((bcast ? 0 : n * C) + c) * V;
          R* on = out + (n * C + c) * V;
          if (dim == 2) {
            R hx = An[0] * fi + An[1] * fj + Tn[0] + ox;
            R hy = An[2] * fi + An[3] * fj + Tn[1] + oy;
            on[(long)i * ny + j] = bilerp<R>(In, hx, hy, nx, ny);
          } else {
            for (int k = 0; k < nz; ++k) {
              R fk = static_cast<R>(k) - oz;
              R hx = An[0] * fi + An[1] * fj + An[2] * fk + Tn[0] + ox;
              R hy = An[3] * fi + An[4] * fj + An[5] * fk + Tn[1] + oy;
              R hz = An[6] * fi + An[7] * fj + An[8] * fk + Tn[2] + oz;
              on[((long)i * ny + j) * nz + k] = trilerp<R>(In, hx, hy, hz, nx, ny, nz);
            }
          }
        }
      }
    }
}

// ---- affine_interp backward ------------------------------------------------------
// cuda/affine.cu:171-328 (2-D), :330-536 (3-D), host :538-610. One block of 16 x 32 threads per
// (n, c); thread (ii, jj) walks i = ii, ii+16, ..., j = jj, jj+32, ..., all k, accumulating its
// partial sums of d_A / d_T in index order; the 512 partials are then tree-reduced by halving
// (256, 128, ..., 1) with tid = ii*32 + jj, exactly the fp32 summation order of the reference.
// d_I is the splat of grad_out (atomic, unordered); C > 1 adds the channels' block results.
template <typename R>
void affine_bwd(R* d_I, R* d_A, R* d_T, const R* go, const R* I, const R* A, const R* T, long N,
                long NI, long C, int dim, const long* sh, int need_I, int need_A, int need_T) {
  const int nx = sh[0], ny = sh[1], nz = dim == 3 ? sh[2] : 1;
  const long V = (long)nx * ny * nz;
  const bool bcast = (NI == 1 && N > 1);
  const int BX = 16, BY = 32, NT = BX * BY;
  const int nA = dim * dim;
  if (need_I) for (long q = 0; q < NI * C * V; ++q) d_I[q] = 0;
  if (need_A) for (long q = 0; q < N * nA; ++q) d_A[q] = 0;
  if (need_T) for (long q = 0; q < N * dim; ++q) d_T[q] = 0;
  const R ox = (R)(.5 * static_cast<R>(nx - 1));
  const R oy = (R)(.5 * static_cast<R>(ny - 1));
  const R oz = (R)(.5 * static_cast<R>(nz - 1));
  for (long n = 0; n < N; ++n)
    for (long c = 0; c < C; ++c) {
      const R* gon = go + (n * C + c) * V;
      const R* In = I + ((bcast ? 0 : n * C) + c) * V;
      R* dIn = need_I ? d_I + ((bcast ? 0 : n * C) + c) * V : nullptr;
      const R* An = A + n * nA;
      const R* Tn = T + n * dim;
      std::vector<R> part((size_t)NT * 12, R(0));  // [tid][0..8] = d_A partials, [9..11] = d_T
#pragma omp parallel for collapse(2) schedule(static)
      for (int ii = 0; ii < BX; ++ii)
        for (int jj = 0; jj < BY; ++jj) {
          R a[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
          for (int i = ii; i < nx; i += BX) {
            R fi = static_cast<R>(i) - ox;
            for (int j = jj; j < ny; j += BY) {
              R fj = static_cast<R>(j) - oy;
              if (dim == 2) {
                long ix = (long)i * ny + j;
                R hx = An[0] * fi + An[1] * fj + Tn[0] + ox;
                R hy = An[2] * fi + An[3] * fj + Tn[1] + oy;
                R diff = gon[ix];
                if (need_I) splat2<R>(dIn, diff, hx, hy, nx, ny);
                if (need_A || need_T) {
                  R gx, gy;
                  bilerp_grad<R>(gx, gy, In, hx, hy, nx, ny);
                  gx *= diff;
                  gy *= diff;
                  if (need_A) {
                    a[0] += gx * fi; a[1] += gx * fj;
                    a[2] += gy * fi; a[3] += gy * fj;
                  }
                  if (need_T) { a[9] += gx; a[10] += gy; }
                }
              } else {
                for (int k = 0; k < nz; ++k) {
                  long ix = ((long)i * ny + j) * nz + k;
                  R fk = static_cast<R>(k) - oz;
                  R hx = An[0] * fi + An[1] * fj + An[2] * fk + Tn[0] + ox;
                  R hy = An[3] * fi + An[4] * fj + An[5] * fk + Tn[1] + oy;
                  R hz = An[6] * fi + An[7] * fj + An[8] * fk + Tn[2] + oz;
                  R diff = gon[ix];
                  if (need_I) splat3<R>(dIn, diff, hx, hy, hz, nx, ny, nz);
                  if (need_A || need_T) {
                    R gx, gy, gz;
                    trilerp_grad<R>(gx, gy, gz, In, hx, hy, hz, nx, ny, nz);
                    gx *= diff;
                    gy *= diff;
                    gz *= diff;
                    if (need_A) {
                      a[0] += gx * fi; a[1] += gx * fj; a[2] += gx * fk;
                      a[3] += gy * fi; a[4] += gy * fj; a[5] += gy * fk;
                      a[6] += gz * fi; a[7] += gz * fj; a[8] += gz * fk;
                    }
                    if (need_T) { a[9] += gx; a[10] += gy; a[11] += gz; }
                  }
                }
              }
            }
          }
          const int tid = ii * BY + jj;
          for (int q = 0; q < 12; ++q) part[(size_t)tid * 12 + q] = a[q];
        }
      for (int h = NT / 2; h >= 1; h /= 2)
        for (int tid = 0; tid < h; ++tid)
          for (int q = 0; q < 12; ++q) part[(size_t)tid * 12 + q] += part[(size_t)(tid + h) * 12 + q];
      if (need_A)
        for (int q = 0; q < nA; ++q) d_A[n * nA + q] += part[q];   // C == 1: plain store; C > 1: atomicAdd
      if (need_T)
        for (int q = 0; q < dim; ++q) d_T[n * dim + q] += part[9 + q];
    }
}

}  // namespace

#define DISPATCH(dtype, CALL_F, CALL_D) \
  do {                                  \
    if ((dtype) == 0) { CALL_F; }       \
    else { CALL_D; }                    \
  } while (0)

extern "C" {

int orc_version() { return 1; }

void orc_interp_fwd(int dtype, void* out, const void* I, const void* u, long N, long NI, long C,
                    int dim, const long* sh, double dt) {
  DISPATCH(dtype, interp_fwd<float>((float*)out, (const float*)I, (const float*)u, N, NI, C, dim, sh, dt),
           interp_fwd<double>((double*)out, (const double*)I, (const double*)u, N, NI, C, dim, sh, dt));
}
void orc_interp_bwd(int dtype, void* d_I, void* d_u, const void* go, const void* I, const void* u,
                    long N, long NI, long C, int dim, const long* sh, double dt, int need_I,
                    int need_u) {
  DISPATCH(dtype,
           interp_bwd<float>((float*)d_I, (float*)d_u, (const float*)go, (const float*)I,
                             (const float*)u, N, NI, C, dim, sh, dt, need_I, need_u),
           interp_bwd<double>((double*)d_I, (double*)d_u, (const double*)go, (const double*)I,
                              (const double*)u, N, NI, C, dim, sh, dt, need_I, need_u));
}
void orc_jtvf_fwd(int dtype, void* out, const void* v, const void* w, long N, long C, int dim,
                  const long* sh, int disp, int trans) {
  DISPATCH(dtype, jtvf_fwd<float>((float*)out, (const float*)v, (const float*)w, N, C, dim, sh, disp, trans),
           jtvf_fwd<double>((double*)out, (const double*)v, (const double*)w, N, C, dim, sh, disp, trans));
}
void orc_jtvf_bwd(int dtype, void* d_v, void* d_w, const void* go, const void* v, const void* w,
                  long N, long C, int dim, const long* sh, int disp, int trans) {
  DISPATCH(dtype,
           jtvf_bwd<float>((float*)d_v, (float*)d_w, (const float*)go, (const float*)v,
                           (const float*)w, N, C, dim, sh, disp, trans),
           jtvf_bwd<double>((double*)d_v, (double*)d_w, (const double*)go, (const double*)v,
                            (const double*)w, N, C, dim, sh, disp, trans));
}
void orc_jtvf_adj_fwd(int dtype, void* out, const void* z, const void* w, long N, long C, int dim,
                      const long* sh) {
  DISPATCH(dtype, jtvf_adj_fwd<float>((float*)out, (const float*)z, (const float*)w, N, C, dim, sh),
           jtvf_adj_fwd<double>((double*)out, (const double*)z, (const double*)w, N, C, dim, sh));
}
void orc_jtvf_adj_bwd(int dtype, void* d_z, void* d_w, const void* go, const void* z,
                      const void* w, long N, long C, int dim, const long* sh) {
  DISPATCH(dtype,
           jtvf_adj_bwd<float>((float*)d_z, (float*)d_w, (const float*)go, (const float*)z,
                               (const float*)w, N, C, dim, sh),
           jtvf_adj_bwd<double>((double*)d_z, (double*)d_w, (const double*)go, (const double*)z,
                                (const double*)w, N, C, dim, sh));
}
void orc_fluid_operator(int dtype, void* Fm, int inverse, const void* cosX, const void* sinX,
                        const void* cosY, const void* sinY, const void* cosZ, const void* sinZ,
                        double alpha, double beta, double gamma, long N, int dim, const long* sh) {
  DISPATCH(dtype,
           fluid_op<float>((float*)Fm, inverse, (const float*)cosX, (const float*)sinX,
                           (const float*)cosY, (const float*)sinY, (const float*)cosZ,
                           (const float*)sinZ, alpha, beta, gamma, N, dim, sh),
           fluid_op<double>((double*)Fm, inverse, (const double*)cosX, (const double*)sinX,
                            (const double*)cosY, (const double*)sinY, (const double*)cosZ,
                            (const double*)sinZ, alpha, beta, gamma, N, dim, sh));
}
void orc_regrid_fwd(int dtype, void* out, const void* I, long N, long C, int dim, const long* sh,
                    const long* osh, const double* origin, const double* spacing) {
  DISPATCH(dtype, regrid_fwd<float>((float*)out, (const float*)I, N, C, dim, sh, osh, origin, spacing),
           regrid_fwd<double>((double*)out, (const double*)I, N, C, dim, sh, osh, origin, spacing));
}
void orc_regrid_bwd(int dtype, void* d_I, const void* go, long N, long C, int dim, const long* sh,
                    const long* osh, const double* origin, const double* spacing) {
  DISPATCH(dtype, regrid_bwd<float>((float*)d_I, (const float*)go, N, C, dim, sh, osh, origin, spacing),
           regrid_bwd<double>((double*)d_I, (const double*)go, N, C, dim, sh, osh, origin, spacing));
}
void orc_affine_interp_fwd(int dtype, void* out, const void* I, const void* A, const void* T,
                           long N, long NI, long C, int dim, const long* sh) {
  DISPATCH(dtype,
           affine_fwd<float>((float*)out, (const float*)I, (const float*)A, (const float*)T, N, NI, C, dim, sh),
           affine_fwd<double>((double*)out, (const double*)I, (const double*)A, (const double*)T, N, NI, C, dim, sh));
}
void orc_affine_interp_bwd(int dtype, void* d_I, void* d_A, void* d_T, const void* go, const void* I,
                           const void* A, const void* T, long N, long NI, long C, int dim, const long* sh,
                           int need_I, int need_A, int need_T) {
  DISPATCH(dtype,
           affine_bwd<float>((float*)d_I, (float*)d_A, (float*)d_T, (const float*)go, (const float*)I,
                             (const float*)A, (const float*)T, N, NI, C, dim, sh, need_I, need_A, need_T),
           affine_bwd<double>((double*)d_I, (double*)d_A, (double*)d_T, (const double*)go, (const double*)I,
                              (const double*)A, (const double*)T, N, NI, C, dim, sh, need_I, need_A, need_T));
}

}  // extern "C"
