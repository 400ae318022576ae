/*
 * Stand-in for Octave/MATLAB's mex.h, TEST INFRASTRUCTURE ONLY.
 *
 * Neither Octave nor MATLAB is installed in the build container, so the
 * reference sources (which only use mxGetPr, mxCreateNumericArray, mexPrintf
 * and mexErrMsgTxt) are compiled against this header when building the
 * checker under oracle/_ref/.  The implementations live in
 * oracle/mex_standin.cpp.  Nothing under opticalflow2d_b200/ includes this
 * file: the product has its own MEX shim (opticalflow2d_b200/host/mex).
 */
#ifndef OF2D_ORACLE_STUB_MEX_H
#define OF2D_ORACLE_STUB_MEX_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef size_t mwSize;
typedef size_t mwIndex;

typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

typedef struct mxArray_tag {
    double *data;
    mwSize  ndim;
    mwSize  dims[4];
    mwSize  numel;
} mxArray;

double  *mxGetPr(const mxArray *a);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity cplx);
void     mxDestroyArray(mxArray *a);

int  mexPrintf(const char *fmt, ...);
void mexErrMsgTxt(const char *msg);   /* throws std::runtime_error in the stand-in */

#ifdef __cplusplus
}
#endif

#endif
