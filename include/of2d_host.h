/*
 * of2d_host.h -- C entry points of libof2d_host{32,64}.so, the C++ host layer that mirrors the
 * reference's class API (ImageRegistration{OpticalFlow,Demons,Fluid}, Image, Motion, Kernel, ...)
 * on top of include/of2d_cuda.h.  Built twice: host32 = float fields (the reference as written),
 * host64 = double fields (fp64 mode).
 *
 * The reference's only foreign-function boundary is its MEX entry point
 *     void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
 * (WrapperOpticalFlow2d.cpp:18-20); both libraries export exactly that symbol with the same five
 * call shapes (see opticalflow2d_b200/host/mex/WrapperOpticalFlow2d.cpp and INTEGRATION.md).  The
 * functions below exist so that a process without an Octave interpreter (tests, bench.py, another
 * FFI) can build the mxArray arguments, call mexFunction, and read back the control-flow trace the
 * reference only prints.
 */
#ifndef OF2D_HOST_H
#define OF2D_HOST_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OF2D_HOST_OK 0
#define OF2D_HOST_EINVAL 2    /* std::invalid_argument escaped (reference: wrong nparams, dimension mismatch) */
#define OF2D_HOST_ERUNTIME 3  /* std::runtime_error escaped (divide by zero, mexErrMsgTxt, CUDA failure) */
#define OF2D_HOST_EOTHER 4

int of2d_host_real_bits(void);                 /* 32 or 64 */
const char *of2d_host_last_error(void);
void of2d_host_capture_printf(int on);         /* collect mexPrintf output instead of dropping it */
const char *of2d_host_printed(void);
int of2d_host_set_strict(int strict);          /* 1: arithmetic level 0 (the reference's loop literally, bit-exact); 0: back to the default level */
int of2d_host_set_math(int level);             /* arithmetic level of the process context: 0 strict, 1 exact engine, 2 relaxed engine (of2d_cuda.h) */
int of2d_host_get_math(void);
int of2d_host_set_stream(void *cuda_stream);   /* run on the caller's cudaStream_t (NULL = legacy default stream) */
int of2d_host_use_own_stream(void);
int of2d_host_sync(void);
unsigned long long of2d_host_launch_count(void);
void of2d_host_shutdown(void);
/* per-kernel CUDA-event timing of the iteration engine (of2d_ctx_profile_* of include/of2d_cuda.h) */
int of2d_host_profile_enable(int on);
int of2d_host_profile_read(char *buf, size_t cap);

/* in-process mxArray (real double only), as the interpreter would provide */
void *of2d_mx_create(int ndim, const size_t *dims);
double *of2d_mx_data(void *mx);
size_t of2d_mx_numel(void *mx);
int of2d_mx_ndim(void *mx);
size_t of2d_mx_dim(void *mx, int d);
void of2d_mx_free(void *mx);

/* mexFunction behind a C status: replaces Octave's `OpticalFlow2d(...)` dispatch (WrapperOpticalFlow2d.cpp:18-155) */
int of2d_mex_call(int nlhs, void **plhs, int nrhs, void **prhs);

/* the same classes without the MEX singleton: ImageRegistration ctor / set_*_image / estimate_motion /
   copy_estimated_motion (src/ImageRegistration.h:14-29) */
typedef struct of2d_session of2d_session;
int of2d_session_create(int dimx, int dimy, int nscales, const int *niter, int nrefine, int reg, const double *regparams, int nparams, int verbose, of2d_session **out);
void of2d_session_destroy(of2d_session *s);
int of2d_session_set_images(of2d_session *s, const double *Iref, const double *Imov);
int of2d_session_estimate(of2d_session *s);
/* extension: cold start for the next estimate (zero motion pyramid and the fluid velocity; the reference warm-starts) */
int of2d_session_reset(of2d_session *s);
int of2d_session_get_motion(of2d_session *s, double *planar_out);      /* 2*N doubles: x plane, y plane */
int of2d_session_get_motion_aos(of2d_session *s, void *out_real);       /* N {x,y} pairs in the field precision */
int of2d_session_warp(of2d_session *s, const double *img, double *out);
/* extension: of2d_session_set_images + of2d_session_estimate + of2d_session_get_motion of n distinct sessions in one call,
   job k+1's host -> device copies and job k-1's device -> host copy running under job k's solve (three streams; pinned host
   buffers make the copies asynchronous, pageable ones still work).  Same results as the separate calls. */
int of2d_sessions_register(of2d_session *const *sessions, int n, const double *const *Iref, const double *const *Imov, double *const *planar_out);

/* the reference's public Image / Motion / Kernel methods that no driver calls, on host arrays:
   op 0: Image::sum / max / min -> scalars[0..2] (src/Image.cpp:78-104); op 1: Image::normalize -> out (:107-116);
   op 2: Image::convolute with Kernel::set_gaussian(sigma) when sigma > 0, else Kernel::set_average -> out (:184-187;
   the reference leaves its float accumulator uninitialised there, src/Field.tpp:240 -- here it starts from zero).
   Images: dimx*dimy column-major doubles; motions: dimx*dimy {x, y} double pairs. */
int of2d_host_image_op(int op, int dimx, int dimy, const double *in, double *out, double *scalars, int kernel_w, double sigma);
int of2d_host_motion_boundary(int kind, int dimx, int dimy, const double *aos_in, double *aos_out);   /* 0 Neumann, 1 Dirichlet (src/Motion.cpp:181-251) */
int of2d_host_kernel(int kind, int w, double sigma, double *out);                                      /* 0 Gaussian, 1 average (src/Kernel.cpp:45-82) */

/* extension: `batch` independent pairs of one size registered together on this process's GPU (BASELINE.json
   configs[4]; one level, cold start per pair = what a fresh ImageRegistration* object with nscales = 0 computes).
   Iref / Imov: batch images back to back; planar_out: per pair the x plane then the y plane.
   wave: pairs resident in the engine at a time (0 = default 128); the waves are balanced and a partial last wave is
   padded internally, so any batch size works. */
typedef struct of2d_batch of2d_batch;
int of2d_batch_create(int dimx, int dimy, int batch, int niter, int nrefine, int reg, const double *regparams, int nparams, int wave, of2d_batch **out);
void of2d_batch_destroy(of2d_batch *b);
int of2d_batch_set_images(of2d_batch *b, const double *Iref, const double *Imov);
int of2d_batch_estimate(of2d_batch *b);
int of2d_batch_get_motion(of2d_batch *b, double *planar_out);
int of2d_batch_iterations(of2d_batch *b, int *iterations, int *regrids);
int of2d_batch_wave(of2d_batch *b);
/* streamed protocol: set_images + estimate + get_motion in one call, wave by wave, with the host -> device copy of wave
   k + 1 and the device -> host copy of wave k - 1 under the solve of wave k (overlap needs pinned host buffers) */
int of2d_batch_register(of2d_batch *b, const double *Iref, const double *Imov, double *planar_out);
/* cine chains: `frames` consecutive frame pairs of batch / frames independent sequences, frame-major (pair f * S + s);
   frame f starts from the motion (Fluid: and the velocity) frame f - 1 ended with -- what the reference does when
   estimate_motion() is called again on one object (src/ImageRegistration.cpp:135-139, OpticalFlowFluid.cpp:50,128) */
int of2d_batch_create_chain(int dimx, int dimy, int batch, int frames, int niter, int nrefine, int reg, const double *regparams, int nparams, of2d_batch **out);
/* one process, several GPUs: the pairs are sharded in contiguous ranges over `devices`; every shard has its own
   context, streams and host thread; there is no exchange between shards.  All of2d_batch_* calls work on the result. */
int of2d_batch_create_multi(int dimx, int dimy, int batch, int niter, int nrefine, int reg, const double *regparams, int nparams, int wave,
                            const int *devices, int ndevices, of2d_batch **out);
int of2d_shard_range(int total, int world, int rank, int *lo, int *hi);   /* the partition (pure host arithmetic; 0 = ok) */
int of2d_batch_num_shards(of2d_batch *b);
int of2d_batch_shard_info(of2d_batch *b, int shard, int *device, int *lo, int *hi);

/* trace of the last estimate_motion(); s == NULL addresses the MEX singleton */
int of2d_trace_num_levels(of2d_session *s);
long of2d_trace_total_iterations(of2d_session *s);
int of2d_trace_level_info(of2d_session *s, int level, int *scale, int *refine, int *iterations, int *nregrid);
int of2d_trace_level_series(of2d_session *s, int level, int which, double *out, int cap);

#ifdef __cplusplus
}
#endif

#endif
