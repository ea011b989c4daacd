// ref_cuda_capi.cu -- TEST INFRASTRUCTURE ONLY (oracle pin).
//
// Compiles the REFERENCE's own CUDA translation units, unmodified, from where
// they lie (-DLGM_REF_DIR=/root/reference/lagomorph/extension) against the ATen
// stand-in in shim/, and wraps their host entry points (extension.cpp:29-103)
// in a C ABI taking raw device pointers. Output: oracle/_ref/libref_cuda.so,
// used only by `-m gpu` tests (to pin the oracle and the CUDA product against
// the reference itself) and by tests/golden/make_golden.py.
#include <ATen/ATen.h>
#define LGM_STR2(x) #x
#define LGM_STR(x) LGM_STR2(x)
#define LGM_REF_FILE(f) LGM_STR(LGM_REF_DIR/cuda/f)

bool lagomorph_debug_mode = false;

#include LGM_REF_FILE(interp.cu)
#include LGM_REF_FILE(diff.cu)
#undef PI
#include LGM_REF_FILE(metric.cu)
#include LGM_REF_FILE(affine.cu)

static at::Tensor wrap(int dtype, const void* p, std::initializer_list<int64_t> sizes) {
  at::Tensor t;
  t.ptr = const_cast<void*>(p);
  t.sizes = sizes;
  t.st = dtype == 0 ? at::ScalarType::Float : at::ScalarType::Double;
  return t;
}
static at::Tensor wrapf(int dtype, const void* p, long N, long C, int dim, const long* sh) {
  if (dim == 2) return wrap(dtype, p, {N, C, sh[0], sh[1]});
  return wrap(dtype, p, {N, C, sh[0], sh[1], sh[2]});
}
static void copy_out(void* dst, const at::Tensor& t) {
  cudaMemcpy(dst, t.ptr, (size_t)t.numel() * t.itemsize(), cudaMemcpyDeviceToDevice);
}
#define GUARD(...)                                                                   \
  try {                                                                              \
    cudaGetLastError(); /* clear */                                                  \
    __VA_ARGS__;                                                                     \
    cudaError_t e = cudaGetLastError(); /* the reference never checks its launches */ \
    if (e == cudaSuccess) e = cudaDeviceSynchronize();                               \
    if (e != cudaSuccess) fprintf(stderr, "ref_cuda: %s\n", cudaGetErrorString(e));  \
    return e == cudaSuccess ? 0 : (int)e;                                            \
  } catch (const std::exception& ex) {                                               \
    fprintf(stderr, "ref_cuda: %s\n", ex.what());                                    \
    return -1;                                                                       \
  }

extern "C" {

int refcu_interp_fwd(int dtype, void* out, const void* I, const void* u, long N, long NI, long C,
                     int dim, const long* sh, double dt) {
  GUARD(copy_out(out, interp_cuda_forward(wrapf(dtype, I, NI, C, dim, sh), wrapf(dtype, u, N, dim, dim, sh), dt)))
}
int refcu_interp_bwd(int dtype, void* d_I, void* d_u, const void* go, const void* I, const void* u,
                     long N, long NI, long C, int dim, const long* sh, double dt, int need_I, int need_u) {
  GUARD(auto r = interp_cuda_backward(wrapf(dtype, go, N, C, dim, sh), wrapf(dtype, I, NI, C, dim, sh),
                                      wrapf(dtype, u, N, dim, dim, sh), dt, need_I, need_u);
        copy_out(d_I, r[0]); copy_out(d_u, r[1]))
}
int refcu_jtvf_fwd(int dtype, void* out, const void* v, const void* w, long N, long C, int dim,
                   const long* sh, int disp, int trans) {
  GUARD(copy_out(out, jacobian_times_vectorfield_forward(wrapf(dtype, v, N, C, dim, sh),
                                                         wrapf(dtype, w, N, dim, dim, sh), disp, trans)))
}
int refcu_jtvf_bwd(int dtype, void* d_v, void* d_w, const void* go, const void* v, const void* w,
                   long N, long C, int dim, const long* sh, int disp, int trans) {
  GUARD(auto r = jacobian_times_vectorfield_backward(wrapf(dtype, go, N, C, dim, sh), wrapf(dtype, v, N, C, dim, sh),
                                                     wrapf(dtype, w, N, dim, dim, sh), disp, trans, true, true);
        copy_out(d_v, r[0]); copy_out(d_w, r[1]))
}
int refcu_jtvf_adj_fwd(int dtype, void* out, const void* z, const void* w, long N, long C, int dim,
                       const long* sh) {
  GUARD(copy_out(out, jacobian_times_vectorfield_adjoint_forward(wrapf(dtype, z, N, C, dim, sh),
                                                                 wrapf(dtype, w, N, dim, dim, sh))))
}
int refcu_jtvf_adj_bwd(int dtype, void* d_z, void* d_w, const void* go, const void* z, const void* w,
                       long N, long C, int dim, const long* sh) {
  GUARD(auto r = jacobian_times_vectorfield_adjoint_backward(wrapf(dtype, go, N, C, dim, sh), wrapf(dtype, z, N, C, dim, sh),
                                                             wrapf(dtype, w, N, dim, dim, sh), true, true);
        copy_out(d_z, r[0]); copy_out(d_w, r[1]))
}
// Fm: (N, dim, X, Y[, Zc], 2) interleaved, modified in place. LUT pointers on device.
int refcu_fluid_operator(int dtype, void* Fm, int inverse, const void* cosX, const void* sinX,
                         const void* cosY, const void* sinY, const void* cosZ, const void* sinZ,
                         double alpha, double beta, double gamma, long N, int dim, const long* sh) {
  GUARD(at::Tensor F = dim == 2 ? wrap(dtype, Fm, {N, 2, sh[0], sh[1], 2})
                                : wrap(dtype, Fm, {N, 3, sh[0], sh[1], sh[2], 2});
        std::vector<at::Tensor> c, s;
        c.push_back(wrap(dtype, cosX, {sh[0]})); s.push_back(wrap(dtype, sinX, {sh[0]}));
        c.push_back(wrap(dtype, cosY, {sh[1]})); s.push_back(wrap(dtype, sinY, {sh[1]}));
        if (dim == 3) { c.push_back(wrap(dtype, cosZ, {sh[2]})); s.push_back(wrap(dtype, sinZ, {sh[2]})); }
        fluid_operator_cuda(F, inverse, c, s, alpha, beta, gamma))
}
int refcu_regrid_fwd(int dtype, void* out, const void* I, long N, long C, int dim, const long* sh,
                     const long* osh, const double* origin, const double* spacing) {
  GUARD(std::vector<int> shape(osh, osh + dim); std::vector<double> o(origin, origin + dim), s(spacing, spacing + dim);
        copy_out(out, regrid_forward(wrapf(dtype, I, N, C, dim, sh), shape, o, s)))
}
int refcu_regrid_bwd(int dtype, void* d_I, const void* go, long N, long C, int dim, const long* sh,
                     const long* osh, const double* origin, const double* spacing) {
  GUARD(std::vector<int> inshape(sh, sh + dim), shape(osh, osh + dim);
        std::vector<double> o(origin, origin + dim), s(spacing, spacing + dim);
        copy_out(d_I, regrid_backward(wrapf(dtype, go, N, C, dim, osh), inshape, shape, o, s)))
}
int refcu_affine_interp_fwd(int dtype, void* out, const void* I, const void* A, const void* T,
                            long N, long NI, long C, int dim, const long* sh) {
  GUARD(copy_out(out, affine_interp_cuda_forward(wrapf(dtype, I, NI, C, dim, sh), wrap(dtype, A, {N, dim, dim}),
                                                 wrap(dtype, T, {N, dim}))))
}
int refcu_affine_interp_bwd(int dtype, void* d_I, void* d_A, void* d_T, const void* go, const void* I,
                            const void* A, const void* T, long N, long NI, long C, int dim, const long* sh) {
  GUARD(auto r = affine_interp_cuda_backward(wrapf(dtype, go, N, C, dim, sh), wrapf(dtype, I, NI, C, dim, sh),
                                             wrap(dtype, A, {N, dim, dim}), wrap(dtype, T, {N, dim}), true, true, true);
        copy_out(d_I, r[0]); copy_out(d_A, r[1]); copy_out(d_T, r[2]))
}

}  // extern "C"
