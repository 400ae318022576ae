/*
 * of2d_cuda.h -- C ABI of libof2d_cuda.so: the hand-written sm_100a kernels that replace the
 * CPU loops of tjwdraper/OpticalFlow2d's per-iteration registration solve.
 *
 * Boundary rules: extern "C", plain pointers and sizes, int status returns (0 = ok), no C++
 * or torch types, no exceptions.  Every `d_` pointer is DEVICE memory; every `h_` pointer is
 * HOST memory.  The reference has no FFI below its MEX entry point (everything is one C++
 * process); this header is the seam the host C++ classes in opticalflow2d_b200/host/ call
 * instead of running the reference's loops, and each entry point cites the reference
 * function (file:line under /root/reference) whose result it reproduces.
 *
 * Layout (identical to the reference): a field of dimx*dimy elements is column-major with x
 * fastest, idx = i + j*dimx (src/Field.tpp:13).  Images are `real`, motion fields are
 * array-of-structs {x, y} of `real` (src/coord2d.h:149).  `_f32` entry points follow the
 * reference as written (float); `_f64` follow the float->double build (fp64 mode).  Batched
 * entry points take `batch` fields stored back to back.
 *
 * There is no CPU fallback: every entry point fails with OF2D_ERR_CUDA when no device is
 * usable.
 */
#ifndef OF2D_CUDA_H
#define OF2D_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---- */
#define OF2D_SUCCESS 0
#define OF2D_ERR_CUDA 1        /* CUDA runtime/driver failure (message in of2d_last_error) */
#define OF2D_ERR_INVALID 2     /* bad argument: maps to std::invalid_argument on the host side */
#define OF2D_ERR_DIVZERO 3     /* a pixel hit the reference's "Divide by zero exception" (src/coord2d.h:95-100) */
#define OF2D_ERR_UNSUPPORTED 4

/* flag bits of the per-pair status word written by kernels */
#define OF2D_FLAG_DIVZERO 1u

typedef struct of2d_ctx of2d_ctx;

/* ---- context, memory, streams ---- */
int of2d_device_count(int *count);
int of2d_ctx_create(int device, of2d_ctx **out);
void of2d_ctx_destroy(of2d_ctx *ctx);
/* run on an externally owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); a NULL handle is the
   legacy default stream, exactly as in the CUDA runtime */
int of2d_ctx_set_stream(of2d_ctx *ctx, void *cuda_stream);
/* back to the context's own non-blocking stream (the state after of2d_ctx_create) */
int of2d_ctx_use_own_stream(of2d_ctx *ctx);
void *of2d_ctx_get_stream(of2d_ctx *ctx);
int of2d_ctx_sync(of2d_ctx *ctx);
/* several contexts in one process (one per GPU and host thread, or extra copy streams on one GPU): make the context's
   device current for the calling host thread; order the waiter's stream after everything enqueued on the signaller's so far */
int of2d_ctx_make_current(of2d_ctx *ctx);
int of2d_ctx_device(of2d_ctx *ctx);
int of2d_ctx_wait_for(of2d_ctx *waiter, of2d_ctx *signaller);
/* arithmetic level of the context (read when an engine is created / a driver loop starts):
     0  strict : the host classes run the reference's loop literally, one per-step kernel per call; every value is
                 the reference's expression in the reference's operation order, unfused (bit-exact parity pin);
     1  exact  : the device-resident engine with the same unfused arithmetic (bit-identical fields to level 0 for
                 Diffusion / Thirion / Diffeomorphic; Elastic / Fluid to 2^-40 of a step; Curvature to 1e-9 px);
     2  relaxed: the engine compiled a second time with FMA contraction, approximate division and algebraically
                 equivalent shortcuts (separable Gaussian taps, linear carry correction in the SOR sweep): results
                 within the north-star tolerances (1e-3 px / 1e-4 SSD in fp32, 1e-6 px in fp64), not bit-identical.
   Default 2; the environment variable OF2D_MATH = strict | exact | relaxed sets the initial level. */
int of2d_ctx_set_fast_math(of2d_ctx *ctx, int on);
int of2d_ctx_get_fast_math(of2d_ctx *ctx);
const char *of2d_last_error(void);
/* how many of this library's kernels have been launched through ctx since creation */
uint64_t of2d_ctx_launch_count(of2d_ctx *ctx);
/* per-kernel timing of the iteration engine with CUDA events on the launching stream: enable (resets), run, read
   ("name launches total_ms" lines; synchronises).  Used by bench.py for the roofline of the dominant kernel. */
int of2d_ctx_profile_enable(of2d_ctx *ctx, int on);
int of2d_ctx_profile_read(of2d_ctx *ctx, char *buf, size_t cap);
/* reads and clears the sticky per-pair flag words (OF2D_FLAG_*) raised by kernels; synchronises the stream */
int of2d_poll_status(of2d_ctx *ctx, int batch, unsigned *h_status);

int of2d_malloc(of2d_ctx *ctx, size_t bytes, void **d_ptr);
int of2d_free(of2d_ctx *ctx, void *d_ptr);
int of2d_host_alloc(size_t bytes, void **h_ptr);   /* pinned */
int of2d_host_free(void *h_ptr);
int of2d_memset(of2d_ctx *ctx, void *d_ptr, int byte, size_t bytes);
int of2d_h2d(of2d_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);  /* stream-ordered; blocks if h_src is pageable */
int of2d_d2h(of2d_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);  /* synchronises the stream before returning */
int of2d_d2h_async(of2d_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int of2d_d2d(of2d_ctx *ctx, void *d_dst, const void *d_src, size_t bytes);

/* ---- I/O casts (K12) ---- */
/* Image::set_image, src/Image.cpp:15-29: double -> real */
int of2d_image_from_double_f32(of2d_ctx *ctx, size_t n, const double *d_in, float *d_out);
int of2d_image_from_double_f64(of2d_ctx *ctx, size_t n, const double *d_in, double *d_out);
/* Image::copy_image_to_input, src/Image.cpp:36-50: real -> double */
int of2d_image_to_double_f32(of2d_ctx *ctx, size_t n, const float *d_in, double *d_out);
int of2d_image_to_double_f64(of2d_ctx *ctx, size_t n, const double *d_in, double *d_out);
/* Motion::copy_motion_to_input, src/Motion.cpp:23-39: AoS real -> planar double (x plane, y plane) */
int of2d_motion_to_planar_double_f32(of2d_ctx *ctx, size_t n, const float *d_u, double *d_out);
int of2d_motion_to_planar_double_f64(of2d_ctx *ctx, size_t n, const double *d_u, double *d_out);
/* `batch` fields in one launch: d_u = [batch][n]{x,y}, d_out = [batch][2][n] (the batch protocols convert a whole wave at once: one launch
   finds its way between the kernels of a running solve, a launch per pair waits for a gap each time) */
int of2d_motion_to_planar_double_batch_f32(of2d_ctx *ctx, size_t n, int batch, const float *d_u, double *d_out);
int of2d_motion_to_planar_double_batch_f64(of2d_ctx *ctx, size_t n, int batch, const double *d_u, double *d_out);

/* ---- field primitives ---- */
/* Image::warp2d, src/Image.cpp:119-182 (out of place: d_dst must not alias d_src) */
int of2d_warp2d_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, const float *d_src, const float *d_u, float *d_dst);
int of2d_warp2d_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, const double *d_src, const double *d_u, double *d_dst);
/* Motion::accumulate, src/Motion.cpp:113-178: d_out = v + u o (id + v) (out of place) */
int of2d_compose_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, const float *d_u, const float *d_v, float *d_out);
int of2d_compose_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, const double *d_u, const double *d_v, double *d_out);
/* Field<vector2d>::convolute, src/Field.tpp:210-269 with a w x w double kernel (column-major, as
   Kernel::get_kernel(), src/Kernel.cpp:40-42); h_kernel is HOST memory; out of place */
int of2d_convolute_motion_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, const float *d_in, float *d_out, const double *h_kernel, int kw, int kh);
int of2d_convolute_motion_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, const double *d_in, double *d_out, const double *h_kernel, int kw, int kh);
/* Field<float>::convolute (Image::convolute, src/Image.cpp:184-187); the reference leaves `val`
   uninitialised for float (src/Field.tpp:240) -- here it starts from zero */
int of2d_convolute_image_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, const float *d_in, float *d_out, const double *h_kernel, int kw, int kh);
int of2d_convolute_image_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, const double *d_in, double *d_out, const double *h_kernel, int kw, int kh);
/* IterativeSolver::set_derivatives, src/regularization/IterativeSolver.cpp:22-56 + src/gradients.h:9-32 */
int of2d_derivatives_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, const float *d_Iref, const float *d_Imov, float *d_gradI, float *d_It);
int of2d_derivatives_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, const double *d_Iref, const double *d_Imov, double *d_gradI, double *d_It);
/* Image::jacobian + Image::min, src/Image.cpp:189-218, :96-104; d_jac may be NULL (min only); h_min receives the minimum */
int of2d_jacobian_f32(of2d_ctx *ctx, int dimx, int dimy, const float *d_u, float *d_jac, float *h_min);
int of2d_jacobian_f64(of2d_ctx *ctx, int dimx, int dimy, const double *d_u, double *d_jac, double *h_min);
/* Field<T>::operator+=, -=, *=, src/Field.tpp:272-345 (n = number of real scalars) */
int of2d_axpy_f32(of2d_ctx *ctx, size_t n, float a, const float *d_x, float *d_y);   /* y += a*x with a in {+1,-1} exact */
int of2d_axpy_f64(of2d_ctx *ctx, size_t n, double a, const double *d_x, double *d_y);
int of2d_scale_f32(of2d_ctx *ctx, size_t n, float a, float *d_x);
int of2d_scale_f64(of2d_ctx *ctx, size_t n, double a, double *d_x);
int of2d_scale_xy_f32(of2d_ctx *ctx, size_t npix, float ax, float ay, float *d_u);   /* Motion::up/downSample rescale, src/Motion.cpp:75-81 */
int of2d_scale_xy_f64(of2d_ctx *ctx, size_t npix, double ax, double ay, double *d_u);
/* Motion::norm and Motion::maxabs, src/Motion.cpp:42-58 (maxabs uses y twice, as the reference does) */
int of2d_motion_norm_f32(of2d_ctx *ctx, size_t npix, const float *d_u, float *h_norm);
int of2d_motion_norm_f64(of2d_ctx *ctx, size_t npix, const double *d_u, double *h_norm);
int of2d_motion_maxabs_f32(of2d_ctx *ctx, size_t npix, const float *d_u, float *h_maxabs);
int of2d_motion_maxabs_f64(of2d_ctx *ctx, size_t npix, const double *d_u, double *h_maxabs);
/* Image::sum / max / min, src/Image.cpp:78-104 */
int of2d_image_stats_f32(of2d_ctx *ctx, size_t n, const float *d_img, float *h_sum, float *h_max, float *h_min);
int of2d_image_stats_f64(of2d_ctx *ctx, size_t n, const double *d_img, double *h_sum, double *h_max, double *h_min);
/* Image::normalize, src/Image.cpp:107-116 */
int of2d_image_normalize_f32(of2d_ctx *ctx, size_t n, float lo, float hi, float *d_img);
int of2d_image_normalize_f64(of2d_ctx *ctx, size_t n, double lo, double hi, double *d_img);
/* Motion::exp, src/Motion.cpp:253-277 (scaling and squaring); d_tmp: scratch of the same size; h_nsquares optional */
int of2d_motion_exp_f32(of2d_ctx *ctx, int dimx, int dimy, float *d_u, float *d_tmp, int *h_nsquares);
int of2d_motion_exp_f64(of2d_ctx *ctx, int dimx, int dimy, double *d_u, double *d_tmp, int *h_nsquares);
/* Field<T>::downSample / upSample, src/Field.tpp:76-206 (ncomp = 1 image, 2 motion; the Motion rescale is separate) */
int of2d_downsample_f32(of2d_ctx *ctx, int ncomp, int inx, int iny, const float *d_in, int outx, int outy, float *d_out);
int of2d_downsample_f64(of2d_ctx *ctx, int ncomp, int inx, int iny, const double *d_in, int outx, int outy, double *d_out);
int of2d_upsample_f32(of2d_ctx *ctx, int ncomp, int inx, int iny, const float *d_in, int outx, int outy, float *d_out);
int of2d_upsample_f64(of2d_ctx *ctx, int ncomp, int inx, int iny, const double *d_in, int outx, int outy, double *d_out);
/* Motion::Neumann_/Dirichlet_boundaryconditions, src/Motion.cpp:181-251 (kind 0 = Neumann, 1 = Dirichlet) */
int of2d_boundary_conditions_f32(of2d_ctx *ctx, int dimx, int dimy, int kind, float *d_u);
int of2d_boundary_conditions_f64(of2d_ctx *ctx, int dimx, int dimy, int kind, double *d_u);

/* ---- Logger (src/Logger.cpp:32-59) ----
 * err = mean||u - prev|| / mean||prev|| (0 when prev == 0), then prev <- u.  h_err receives err. */
int of2d_logger_update_f32(of2d_ctx *ctx, size_t npix, const float *d_u, float *d_prev, float *h_err);
int of2d_logger_update_f64(of2d_ctx *ctx, size_t npix, const double *d_u, double *d_prev, double *h_err);

/* ---- per-iteration solver steps (src/regularization) ----
 * Each *_step entry point performs exactly one get_update() of the reference solver.  h_status
 * (optional) receives OF2D_FLAG_* bits raised by the kernels of that step. */

/* OpticalFlow::get_force, OpticalFlow.cpp:15-39: f = gradI (It + u . gradI) */
int of2d_lssd_force_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, const float *d_gradI, const float *d_It, const float *d_u, float *d_force);
int of2d_lssd_force_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, const double *d_gradI, const double *d_It, const double *d_u, double *d_force);

/* OpticalFlowDiffusion::get_update, OpticalFlowDiffusion.cpp:43-84 (Horn-Schunck Jacobi); out of place */
int of2d_diffusion_step_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, const float *d_u, float *d_unew, const float *d_gradI, const float *d_It, float alpha, unsigned *h_status);
int of2d_diffusion_step_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, const double *d_u, double *d_unew, const double *d_gradI, const double *d_It, double alpha, unsigned *h_status);

/* OpticalFlowElastic::get_update, OpticalFlowElastic.cpp:13-55: force from the pre-sweep u, then one
   in-place lexicographic SOR sweep (exact sequential order, executed as a t = 2i + j wavefront) */
int of2d_elastic_step_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, float *d_u, const float *d_gradI, const float *d_It, float mu, float lambda, float omega);
int of2d_elastic_step_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, double *d_u, const double *d_gradI, const double *d_It, double mu, double lambda, double omega);

/* OpticalFlowFluid::get_update, OpticalFlowFluid.cpp:123-140: force(u) -> SOR sweep on the persistent
   velocity -> increment R -> dt = 0.65/maxabs(R) -> u += dt R unless dt >= 65.
   h_maxabs / h_dt (optional) receive the values the reference prints (OpticalFlowFluid.cpp:94). */
int of2d_fluid_step_f32(of2d_ctx *ctx, int dimx, int dimy, float *d_u, float *d_velocity, float *d_increment, const float *d_gradI, const float *d_It, float mu, float lambda, float omega, float *h_maxabs, float *h_dt);
int of2d_fluid_step_f64(of2d_ctx *ctx, int dimx, int dimy, double *d_u, double *d_velocity, double *d_increment, const double *d_gradI, const double *d_It, double mu, double lambda, double omega, double *h_maxabs, double *h_dt);

/* Demons::demons_iteration fused with the warp and derivative passes that precede it in
   DemonsThirions::get_update (DemonsThirions.cpp:18-27, Demons.cpp:34-63):
   c = -gradI*It / (|gradI|^2 + It^2 sigma_i^2 / sigma_x^2), Iwar = warp(Imov, u). */
int of2d_demons_force_f32(of2d_ctx *ctx, int dimx, int dimy, int batch, const float *d_Iref, const float *d_Imov, const float *d_u, float *d_corr, float sigma_i, float sigma_x, unsigned *h_status);
int of2d_demons_force_f64(of2d_ctx *ctx, int dimx, int dimy, int batch, const double *d_Iref, const double *d_Imov, const double *d_u, double *d_corr, double sigma_i, double sigma_x, unsigned *h_status);

/* Demons::demons_iteration alone (Demons.cpp:34-63) on stored derivatives */
int of2d_demons_correspondence_f32(of2d_ctx *ctx, size_t npix, const float *d_gradI, const float *d_It, float *d_corr, float sigma_i, float sigma_x);
int of2d_demons_correspondence_f64(of2d_ctx *ctx, size_t npix, const double *d_gradI, const double *d_It, double *d_corr, double sigma_i, double sigma_x);

/* OpticalFlowCurvature (OpticalFlowCurvature.cpp:6-167): plan = eigenvalue table + twiddles for a
   dimx x dimy grid; step = force, rhs, 2-D DCT-II, eigenvalue multiply, 2-D DCT-III, rescale. */
typedef struct of2d_curvature_plan of2d_curvature_plan;
int of2d_curvature_plan_create(of2d_ctx *ctx, int dimx, int dimy, double alpha, double tau, int real_is_double, of2d_curvature_plan **out);
void of2d_curvature_plan_destroy(of2d_curvature_plan *plan);
int of2d_curvature_step_f32(of2d_curvature_plan *plan, const float *d_u, float *d_unew, const float *d_gradI, const float *d_It);
int of2d_curvature_step_f64(of2d_curvature_plan *plan, const double *d_u, double *d_unew, const double *d_gradI, const double *d_It);
/* unnormalised 2-D DCT-II (kind 2) / DCT-III (kind 3) of a row-major n0 x n1 double array, the
   transform fftw_plan_r2r_2d(REDFT10 / REDFT01) computes in the reference; in place on the device */
int of2d_dct2d_f64(of2d_ctx *ctx, int n0, int n1, int kind, double *d_data);

/* ---- iteration engine ----
 * One refine pass of the reference's driver loops for `batch` independent pairs, entirely enqueued on
 * the device: Iaux = Imov o (id + motion); [derivatives]; up to niter x get_update + Logger::update_error
 * with the break `err < 0.001 && iter > 1`; Fluid: Jacobian-triggered regridding; finally
 * motion <- estimate + motion o (id + estimate).
 *   ImageRegistrationOpticalFlow.cpp:97-151, ImageRegistrationDemons.cpp:86-137, ImageRegistrationFluid.cpp:67-142.
 * Decisions (break, time step, regrid, number of squarings) are taken by device-side reductions; the host
 * only polls a counter of running pairs.  The engine exists in two builds, picked from the context's arithmetic level
 * when the engine is created (of2d_ctx_set_fast_math): level 1 "exact" = the reference's unfused arithmetic (Elastic /
 * Fluid sweep as overlapped tiles, csrc/sor_tile.cuh), level 2 "relaxed" = FMA contraction, approximate division and
 * equivalent shortcuts (csrc/engine_relaxed.cu); the per-step entry points above are bit-exact at every level. */
typedef struct of2d_engine of2d_engine;
typedef struct {
    int method;                 /* enum Regularisation, src/SolverOptions.h:4: 0 Diffusion .. 5 Fluid */
    int dimx, dimy, batch;      /* fields of the batch are stored back to back */
    int real_is_double;
    int max_iter;               /* largest niter that will be passed to of2d_engine_refine (trace capacity) */
    double alpha, tau;          /* Diffusion: alpha; Curvature: alpha, tau */
    double mu, lambda, omega;   /* Elastic, Fluid */
    double sigma_i, sigma_x;    /* Demons */
    int kernel_w;               /* Demons: width of both Gaussian kernels */
    const double *kernel_fluid;      /* HOST, kernel_w * kernel_w doubles as Kernel::get_kernel() (smooths the correspondence) */
    const double *kernel_diffusion;  /* HOST (smooths the motion) */
    int accumulation;           /* Thirion: 0 composition, 1 addition (enum MotionAccumulation) */
} of2d_engine_desc;
/* OF2D_ERR_UNSUPPORTED: the parameters need the exact (per-step) path -- callers fall back to the *_step entry points */
int of2d_engine_create(of2d_ctx *ctx, const of2d_engine_desc *desc, of2d_engine **out);
void of2d_engine_destroy(of2d_engine *engine);
/* zeroes the state the reference carries between calls (the Fluid velocity) */
int of2d_engine_reset_state(of2d_engine *engine);
/* d_Iref, d_Imov: batch images; d_motion: batch level motions, updated in place.  Synchronises before returning.
   OF2D_ERR_DIVZERO when a pixel hit the reference's divide-by-zero throw. */
int of2d_engine_refine_f32(of2d_engine *engine, const float *d_Iref, const float *d_Imov, float *d_motion, int niter);
int of2d_engine_refine_f64(of2d_engine *engine, const double *d_Iref, const double *d_Imov, double *d_motion, int niter);
/* what the last refine did for one pair: iterations executed, regrid events, last Logger error */
int of2d_engine_pair_result(of2d_engine *engine, int pair, int *iterations, int *nregrid, double *last_err);
/* per-iteration series of the last refine: which = 0 Logger error, 1 maxabs (fluid increment / demons correspondence),
   2 fluid time step, 3 fluid min Jacobian, 4 fluid regrid flag, 5 diffeomorphic squarings */
int of2d_engine_trace(of2d_engine *engine, int pair, int which, double *h_out, int count);
uint64_t of2d_engine_iterations_enqueued(of2d_engine *engine);
/* 1 when the engine was created at arithmetic level 2 (relaxed), 0 for the exact build */
int of2d_engine_is_relaxed(of2d_engine *engine);

#ifdef __cplusplus
}
#endif

#endif /* OF2D_CUDA_H */
