/*
 * mex.h -- the subset of the Octave / MATLAB MEX API that the OpticalFlow2d entry point uses
 * (mxGetPr, mxCreateNumericArray, mexPrintf, mexErrMsgTxt; reference WrapperOpticalFlow2d.cpp:18-155).
 *
 * When the library is built with mkoctfile / mex, the interpreter's own mex.h comes first on the
 * include path and this file is not used.  In this repository (no Octave in the image) it is paired
 * with mex/mex_harness.cpp, which implements the same four functions in-process so that tests and
 * benchmarks can replay the 5-call protocol of test_opticalflow2d.m:42-59 through ctypes.
 */
#ifndef OF2D_HOST_MEX_H
#define OF2D_HOST_MEX_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef size_t mwSize;
typedef size_t mwIndex;

typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

typedef struct mxArray_tag mxArray;

double *mxGetPr(const mxArray *a);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity cplx);
void mxDestroyArray(mxArray *a);

int mexPrintf(const char *fmt, ...);
void mexErrMsgTxt(const char *msg);

/* the entry point every MEX file exports */
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

#ifdef __cplusplus
}
#endif

#endif
