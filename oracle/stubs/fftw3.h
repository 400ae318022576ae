/*
 * Stand-in for fftw3.h, TEST INFRASTRUCTURE ONLY.
 *
 * fftw3 is an un-vendored, un-pinned dependency of the reference (linked with
 * -lfftw3, compile_mex_function.m:30-32) and is not installed here.  The
 * reference uses exactly three entry points, all from
 * src/regularization/OpticalFlow/OpticalFlowCurvature.cpp:52-55,64-67,152-160:
 * fftw_plan_r2r_2d with kinds REDFT10/REDFT01, fftw_execute_r2r and
 * fftw_destroy_plan.  oracle/dct_standin.c restates the published FFTW
 * definitions of those two transforms (FFTW manual, "1d Real-even DFTs"):
 *   REDFT10:  Y_k = 2 * sum_j X_j cos(pi (j+1/2) k / n)
 *   REDFT01:  Y_k = X_0 + 2 * sum_{j>=1} X_j cos(pi j (k+1/2) / n)
 * and tests/ checks it against scipy.fft.dct(type=2|3, norm=None).
 */
#ifndef OF2D_ORACLE_STUB_FFTW3_H
#define OF2D_ORACLE_STUB_FFTW3_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { FFTW_REDFT10 = 5, FFTW_REDFT01 = 4 } fftw_r2r_kind;
#define FFTW_MEASURE (0U)

struct of2d_dct_plan;
typedef struct of2d_dct_plan *fftw_plan;

fftw_plan fftw_plan_r2r_2d(int n0, int n1, double *in, double *out,
                           fftw_r2r_kind kind0, fftw_r2r_kind kind1, unsigned flags);
void fftw_execute_r2r(const fftw_plan p, double *in, double *out);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif

#endif
