/* lagomorph_b200.h -- C ABI of liblagomorph_b200.so (sm_100a).
 *
 * Drop-in boundary for the LDDMM geodesic-shooting hot path of
 * jacobhinkle/lagomorph. Each entry point replaces one function of the
 * reference's pybind11 module `lagomorph_ext` (lagomorph/extension/extension.cpp,
 * PYBIND11_MODULE at :175-189); the replaced binding is cited per function.
 *
 * Conventions (all entry points)
 *  - Tensors are CONTIGUOUS, layout N C X Y [Z]; channel d of a vector field is
 *    the component along spatial axis d; displacements are in voxels.
 *  - `dtype`: LGM_F32 or LGM_F64 (all tensors of one call share it).
 *  - `dim` is 2 or 3; `shape` points to `dim` int64 extents (X, Y[, Z]).
 *  - All pointers are DEVICE pointers on the current CUDA device, borrowed for
 *    the call. The library never allocates tensors: the caller provides outputs
 *    (and the FFT workspace). Outputs need not be zeroed unless stated.
 *  - `stream` is a cudaStream_t (as void*); work is enqueued, never synchronised
 *    (unless debug mode is on), so calls are CUDA-graph capturable.
 *  - Return value: 0 on success, otherwise a negative LGM_E* code or a positive
 *    cudaError_t; lgm_last_error() gives a message. No exceptions cross the ABI.
 */
#ifndef LAGOMORPH_B200_H
#define LAGOMORPH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGM_F32 0
#define LGM_F64 1

#define LGM_OK 0
#define LGM_EINVAL (-1)  /* bad argument (dim, dtype, shape, channel count) */
#define LGM_ENOSPC (-2)  /* workspace too small */
#define LGM_EUNSUP (-3)  /* unsupported configuration */

/* library identity / diagnostics ------------------------------------------- */
int lgm_version(void);                 /* ABI version, currently 1 */
const char* lgm_last_error(void);      /* message for the last non-zero return (thread-local) */
/* replaces lagomorph_ext.set_debug_mode (extension.cpp:105-107,176): when on,
 * every entry point synchronises the stream and returns the CUDA error. */
void lgm_set_debug_mode(int on);
int lgm_get_debug_mode(void);
/* number of kernel launches issued by this library since process start */
int64_t lgm_launch_count(void);

/* Bench-only profiling: between begin and end an event is recorded after every kernel this
 * library launches on `stream`; end() synchronises and writes a JSON object
 * {"kernel": {"launches": n, "ms": total}} (time of a launch = gap to the previous event). */
int lgm_profile_begin(void* stream);
int lgm_profile_end(char* json, int64_t json_bytes);

/* interp --------------------------------------------------------------------
 * replaces interp_forward (extension.cpp:135-143 -> cuda/interp.cu:80-130):
 *   out[n,c,x] = lerp_clamp(I[n or 0, c], x + dt*u[n,:,x]);  I broadcasts iff NI < N.
 * I: (NI,C,..)  u: (N,dim,..)  out: (N,C,..) */
int lgm_interp_fwd(int dtype, void* out, const void* I, const void* u, int64_t N, int64_t NI,
                   int64_t C, int dim, const int64_t* shape, double dt, void* stream);
/* replaces interp_backward (extension.cpp:145-156 -> cuda/interp.cu:246-313).
 * d_I (like I) is the adjoint splat of gout (== `interp_adjoint`), d_u (like u)
 * the gradient w.r.t. the displacement. Either may be NULL (not needed).
 * d_I is zero-filled by the callee before accumulation. */
int lgm_interp_bwd(int dtype, void* d_I, void* d_u, const void* gout, const void* I,
                   const void* u, int64_t N, int64_t NI, int64_t C, int dim,
                   const int64_t* shape, double dt, void* stream);

/* jacobian_times_vectorfield -------------------------------------------------
 * replaces jacobian_times_vectorfield_forward (extension.cpp:69-73 -> cuda/diff.cu:129-185):
 *   out_c = sum_d (D_d v_c + [displacement] delta_cd) w_d        (transpose = 0)
 *   out_d = sum_c (D_d v_c + [displacement] delta_cd) w_c        (transpose = 1)
 * v,out: (N,C,..)  w: (N,dim,..); transpose/displacement need C == dim. */
int lgm_jtvf_fwd(int dtype, void* out, const void* v, const void* w, int64_t N, int64_t C,
                 int dim, const int64_t* shape, int displacement, int transpose, void* stream);
/* replaces jacobian_times_vectorfield_backward (cuda/diff.cu:475-540); NULL = not needed */
int lgm_jtvf_bwd(int dtype, void* d_v, void* d_w, const void* gout, const void* v, const void* w,
                 int64_t N, int64_t C, int dim, const int64_t* shape, int displacement,
                 int transpose, void* stream);
/* replaces jacobian_times_vectorfield_adjoint_forward (cuda/diff.cu:634-672):
 *   out_c = sum_d D_d^T (w_d z_c) */
int lgm_jtvf_adj_fwd(int dtype, void* out, const void* z, const void* w, int64_t N, int64_t C,
                     int dim, const int64_t* shape, void* stream);
/* replaces jacobian_times_vectorfield_adjoint_backward (cuda/diff.cu:783-835); C == dim */
int lgm_jtvf_adj_bwd(int dtype, void* d_z, void* d_w, const void* gout, const void* z,
                     const void* w, int64_t N, int64_t C, int dim, const int64_t* shape,
                     void* stream);

/* fluid metric ---------------------------------------------------------------
 * replaces the rfft -> fluid_operator -> irfft triple of FluidMetricOperator
 * (lagomorph/metric.py:11-19; fluid_operator: extension.cpp:158-173 ->
 * cuda/metric.cu:308-355) with own batched R2C/C2R FFT passes and the Fourier
 * multiplier fused into the middle pass. out may alias in.
 *   inverse=1: sharp, out = F^-1[ L(k)^-1 F[in] ];  inverse=0: flat, L(k) F[in].
 * in/out: (N,dim,..) real. workspace: lgm_fluid_workspace_bytes() bytes. */
int64_t lgm_fluid_workspace_bytes(int dtype, int64_t N, int dim, const int64_t* shape);
int lgm_fluid_apply(int dtype, void* out, const void* in, int64_t N, int dim,
                    const int64_t* shape, int inverse, double alpha, double beta, double gamma,
                    void* workspace, int64_t workspace_bytes, void* stream);
/* replaces fluid_operator itself (in place on the interleaved half spectrum
 * (N,dim,X,Y[,Zc],2) produced by a unitary rfft); LUTs as metric.py:65-75.
 * spec_shape = (X, Y[, Zc]). cos/sin LUT d has spec_shape[d] entries. */
int lgm_fluid_operator(int dtype, void* Fm, int inverse, const void* const* cos_luts,
                       const void* const* sin_luts, double alpha, double beta, double gamma,
                       int64_t N, int dim, const int64_t* spec_shape, void* stream);

/* regrid -----------------------------------------------------------------------
 * replaces regrid_forward / regrid_backward (extension.cpp:56-67 ->
 * cuda/affine.cu:683-734, :802-855): out[i] = lerp_clamp(I, (i-(N'-1)/2)*S + O) */
int lgm_regrid_fwd(int dtype, void* out, const void* I, int64_t N, int64_t C, int dim,
                   const int64_t* shape, const int64_t* out_shape, const double* origin,
                   const double* spacing, void* stream);
int lgm_regrid_bwd(int dtype, void* d_I, const void* gout, int64_t N, int64_t C, int dim,
                   const int64_t* shape, const int64_t* out_shape, const double* origin,
                   const double* spacing, void* stream);

/* affine_interp ------------------------------------------------------------------
 * replaces affine_interp_forward/backward (extension.cpp:109-133 -> cuda/affine.cu:114-169, :538-610)
 * A: (N,dim,dim) row-major, T: (N,dim); I broadcasts iff NI == 1 && N > 1.
 * d_A/d_T are zero-filled by the callee; any of d_I/d_A/d_T may be NULL. */
int lgm_affine_interp_fwd(int dtype, void* out, const void* I, const void* A, const void* T,
                          int64_t N, int64_t NI, int64_t C, int dim, const int64_t* shape,
                          void* stream);
int lgm_affine_interp_bwd(int dtype, void* d_I, void* d_A, void* d_T, const void* gout,
                          const void* I, const void* A, const void* T, int64_t N, int64_t NI,
                          int64_t C, int dim, const int64_t* shape, void* stream);

/* fused hot-path operators (no reference FFI twin: each fuses a chain of the
 * reference's Python-level calls into one kernel; results equal the chain) ------
 * Ad_star(phiinv, m) = jtvf(phiinv, interp(m, phiinv), displacement=1)
 *   (lagomorph/adjrep.py:86-97). */
int lgm_Ad_star_fwd(int dtype, void* out, const void* phiinv, const void* m, int64_t N, int dim,
                    const int64_t* shape, void* stream);
/* ad_star(v, m) = jtvf(v, m, 0, transpose=1) - jtvf_adjoint(m, v)  (adjrep.py:69-83) */
int lgm_ad_star_fwd(int dtype, void* out, const void* v, const void* m, int64_t N, int dim,
                    const int64_t* shape, void* stream);
/* ad(v, w) = jtvf(v, w, 0, 0) - jtvf(w, v, 0, 0)  (adjrep.py:37-47) */
int lgm_ad_fwd(int dtype, void* out, const void* v, const void* w, int64_t N, int dim,
               const int64_t* shape, void* stream);
/* compose(u, v, ds, dt) = ds*u(x) + dt*v(x + ds*u(x))  (lagomorph/deform.py:53-55); C = dim */
int lgm_compose_fwd(int dtype, void* out, const void* u, const void* v, int64_t N, int dim,
                    const int64_t* shape, double ds, double dt, void* stream);
/* One forward EPDiff step (lagomorph/lddmm.py:39-44), mommask optional (NULL):
 *   m = Ad_star(phiinv, m0) [* mommask]; v = sharp(m); phiinv_out = compose(v, phiinv, -dt, 1)
 * scratch: lgm_epdiff_scratch_bytes() bytes (holds m/v and the FFT workspace).
 * phiinv_out must not alias phiinv. */
int64_t lgm_epdiff_scratch_bytes(int dtype, int64_t N, int dim, const int64_t* shape);
int lgm_epdiff_step_fwd(int dtype, void* phiinv_out, const void* phiinv, const void* m0,
                        const void* mommask, int64_t N, int dim, const int64_t* shape, double dt,
                        double alpha, double beta, double gamma, void* scratch,
                        int64_t scratch_bytes, void* stream);
/* The whole forward shoot, expmap without autograd (lagomorph/lddmm.py:73-91, the loop :87-91):
 *   phiinv_0 = phiinv_in (NULL: zeros, lddmm.py:84-85); num_steps times lgm_epdiff_step_fwd.
 * Same kernels and results as that loop; one call per shoot (no per-step host work, explicit
 * traversal directions alternating from step to step, CUDA-graph capturable).
 * scratch: lgm_expmap_scratch_bytes() bytes. phiinv_out must not alias phiinv_in. */
int64_t lgm_expmap_scratch_bytes(int dtype, int64_t N, int dim, const int64_t* shape);
int lgm_expmap_fwd(int dtype, void* phiinv_out, const void* phiinv_in, const void* m0,
                   const void* mommask, int64_t N, int dim, const int64_t* shape, double dt,
                   int num_steps, double alpha, double beta, double gamma, void* scratch,
                   int64_t scratch_bytes, void* stream);
/* Backward of one EPDiff step: what autograd does for lddmm.py:39-44 through
 * InterpFunction.backward (deform.py:31-41), JacobianTimesVectorFieldFunction.backward
 * (diff.py:29-35) and FluidMetricOperator.backward (metric.py:21-34), as three fused kernels
 * around one sharp. fp32, 3-D, shape[2] % 32 == 0 only: lgm_epdiff_bwd_scratch_bytes returns -1
 * (and the step LGM_EUNSUP) otherwise; callers then chain lgm_interp_bwd / lgm_jtvf_bwd.
 *   g_phi     in: dL/dphiinv_out; out (need_phi): dL/dphiinv. Updated in place.
 *   d_m0      accumulator (need_m0): dL/dm0 of this step is ADDED; zero it before the first step.
 *   splat_acc (N,3,...) accumulator used when need_phi: all zeros on entry, all zeros again on return.
 *   phiinv, v the step's input displacement and its velocity v = sharp(Ad_star(phiinv, m0)*mommask). */
int64_t lgm_epdiff_bwd_scratch_bytes(int dtype, int64_t N, int dim, const int64_t* shape);
int lgm_epdiff_step_bwd(int dtype, void* g_phi, void* d_m0, void* splat_acc, const void* phiinv,
                        const void* v, const void* m0, const void* mommask, int64_t N, int dim,
                        const int64_t* shape, double dt, double alpha, double beta, double gamma,
                        void* scratch, int64_t scratch_bytes, int need_phi, int need_m0, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAGOMORPH_B200_H */
