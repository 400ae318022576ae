/*
 * of2d_oracle.h -- plain-C CPU restatement of the OpticalFlow2d per-iteration
 * registration solve.  TEST INFRASTRUCTURE ONLY: this is the checker the CUDA
 * path is compared against; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED against the reference itself.  The reference has no
 * tests or golden vectors, so the pin is
 * oracle/_ref/libof2d_ref{32,64}.so -- the reference's own sources compiled
 * unchanged in the authoring container -- and tests/test_oracle_cpu.py
 * requires this restatement to reproduce its outputs bit for bit, plus the
 * fixtures in tests/golden/ generated from it (tests/golden/make_golden.py).
 *
 * Built twice (oracle/Makefile): OF2D_REAL_IS_DOUBLE=0 follows the reference as
 * written (float fields); =1 follows the `sed s/float/double/` fp64 variant
 * (float-suffixed literals such as 0.66f, 0.65f, 0.001f stay float-valued).
 *
 * Layout: column-major with x fastest, idx = i + j*dimx (src/Field.tpp:13);
 * motion is array-of-structs {x, y} (src/coord2d.h:149).
 */
#ifndef OF2D_ORACLE_H
#define OF2D_ORACLE_H

#if OF2D_REAL_IS_DOUBLE
typedef double real;
#else
typedef float real;
#endif

typedef struct { real x, y; } vec2;

/* status codes shared with oracle/ref_shim.cpp */
#define OF2D_OK 0
#define OF2D_EINVAL 2   /* std::invalid_argument in the reference */
#define OF2D_ERUNTIME 3 /* std::runtime_error: divide by zero (coord2d.h:95-100), mexErrMsgTxt */

#ifdef __cplusplus
extern "C" {
#endif

int of2d_oracle_sizeof_real(void);
const char *of2d_oracle_last_error(void);

int of2d_oracle_mex_register(int dimx, int dimy, int nscales, const double *niter, int nrefine, int reg,
                             const double *regparams, int nparams, int verbose, const double *Iref,
                             const double *Imov, double *motion_out, double *warped_out);
int of2d_oracle_mex_badcall(int nlhs, int nrhs);

int of2d_oracle_set_image(int dimx, int dimy, const double *in, real *out);
int of2d_oracle_copy_motion_to_input(int dimx, int dimy, const real *u, double *out);
int of2d_oracle_warp2d(int dimx, int dimy, real *img, const real *u);
int of2d_oracle_accumulate(int dimx, int dimy, real *u, const real *v);
int of2d_oracle_gaussian_kernel(int w, real sigma, double *out);
int of2d_oracle_convolute_motion(int dimx, int dimy, real *u, int w, real sigma);
int of2d_oracle_image_stats(int dimx, int dimy, const real *img, real *sum, real *mx, real *mn);
int of2d_oracle_image_normalize(int dimx, int dimy, real *img);
int of2d_oracle_boundary_conditions(int dimx, int dimy, int kind, real *u);
int of2d_oracle_average_kernel(int w, double *out);
int of2d_oracle_convolute_image(int dimx, int dimy, real *img, int w, real sigma);
int of2d_oracle_exp(int dimx, int dimy, real *u);
int of2d_oracle_norm_maxabs(int dimx, int dimy, const real *u, real *norm, real *maxabs);
int of2d_oracle_jacobian(int dimx, int dimy, const real *u, real *jac, real *minjac);
int of2d_oracle_derivatives(int dimx, int dimy, const real *Iref, const real *Imov, real *grad, real *It);
int of2d_oracle_image_resample(int inx, int iny, const real *in, int outx, int outy, real *out, int up);
int of2d_oracle_motion_resample(int inx, int iny, const real *in, int outx, int outy, real *out, int up);
int of2d_oracle_logger(int dimx, int dimy, const real *useq, int nseq, real *err);
int of2d_oracle_solver_steps(int reg, const real *params, int nparams, int dimx, int dimy, const real *Iref,
                             const real *Imov, real *u, int nsteps);

/* captured control-flow trace (same meaning as oracle/mex_standin.cpp) */
void of2d_trace_reset(void);
void of2d_trace_enable(int on);
int of2d_trace_count(int which);
void of2d_trace_get(int which, double *a, double *b, int cap);

void of2d_standin_dct1d(double *x, int n, int type);

#ifdef __cplusplus
}
#endif

#endif
