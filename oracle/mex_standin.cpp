/*
 * mex_standin.cpp -- implementation of oracle/stubs/mex.h.
 * TEST INFRASTRUCTURE ONLY: linked into the checker libraries under
 * oracle/_ref/ so the unmodified reference sources can run without Octave.
 *
 * mexPrintf is a capturing no-op: the three printf call sites that carry
 * control-flow information are recognised by their format string and their
 * varargs recorded at full precision:
 *   src/Logger.cpp:64                         "Iteration: %d\tError:%.4f\n"
 *   src/ImageRegistrationFluid.cpp:110        "Regridding on iteration: %d\tMin Jacobian: %.3f\n"
 *   src/regularization/OpticalFlow/OpticalFlowFluid.cpp:94  "Dumax: %.3f\tMaxabs increment: %.3f\t Timestep: %.3f\n"
 * mexErrMsgTxt throws (under Octave it never returns).
 */
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "stubs/mex.h"

namespace {
struct Trace {
    std::vector<int> err_iter;
    std::vector<double> err_val;
    std::vector<int> regrid_iter;
    std::vector<double> regrid_minjac;
    std::vector<double> fluid_maxabs;
    std::vector<double> fluid_dt;
    bool enabled = true;
};
Trace g_trace;
}  // namespace

extern "C" {

double *mxGetPr(const mxArray *a) { return a->data; }

mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID, mxComplexity) {
    mxArray *a = new mxArray;
    a->ndim = ndim;
    a->numel = 1;
    for (mwSize d = 0; d < 4; d++) a->dims[d] = 1;
    for (mwSize d = 0; d < ndim && d < 4; d++) { a->dims[d] = dims[d]; a->numel *= dims[d]; }
    a->data = static_cast<double *>(calloc(a->numel ? a->numel : 1, sizeof(double)));
    return a;
}

void mxDestroyArray(mxArray *a) {
    if (!a) return;
    free(a->data);
    delete a;
}

int mexPrintf(const char *fmt, ...) {
    if (!g_trace.enabled) return 0;
    va_list ap;
    va_start(ap, fmt);
    if (strncmp(fmt, "Iteration: %d", 13) == 0) {
        int it = va_arg(ap, int);
        double e = va_arg(ap, double);
        g_trace.err_iter.push_back(it);
        g_trace.err_val.push_back(e);
    } else if (strncmp(fmt, "Regridding on iteration", 23) == 0) {
        int it = va_arg(ap, int);
        double mj = va_arg(ap, double);
        g_trace.regrid_iter.push_back(it);
        g_trace.regrid_minjac.push_back(mj);
    } else if (strncmp(fmt, "Dumax:", 6) == 0) {
        (void)va_arg(ap, double);
        double ma = va_arg(ap, double);
        double dt = va_arg(ap, double);
        g_trace.fluid_maxabs.push_back(ma);
        g_trace.fluid_dt.push_back(dt);
    }
    va_end(ap);
    return 0;
}

void mexErrMsgTxt(const char *msg) { throw std::runtime_error(std::string("mexErrMsgTxt: ") + msg); }

/* ---- trace access for the tests ---- */
void of2d_trace_reset(void) { g_trace = Trace(); }
void of2d_trace_enable(int on) { g_trace.enabled = on != 0; }
int of2d_trace_count(int which) {
    switch (which) {
        case 0: return (int)g_trace.err_val.size();
        case 1: return (int)g_trace.regrid_iter.size();
        case 2: return (int)g_trace.fluid_dt.size();
    }
    return 0;
}
/* which: 0 = logger error (iter, err), 1 = regrid (iter, minjac), 2 = fluid (maxabs, dt) */
void of2d_trace_get(int which, double *a, double *b, int cap) {
    int n = of2d_trace_count(which);
    if (n > cap) n = cap;
    for (int i = 0; i < n; i++) {
        if (which == 0) { a[i] = g_trace.err_iter[i]; b[i] = g_trace.err_val[i]; }
        if (which == 1) { a[i] = g_trace.regrid_iter[i]; b[i] = g_trace.regrid_minjac[i]; }
        if (which == 2) { a[i] = g_trace.fluid_maxabs[i]; b[i] = g_trace.fluid_dt[i]; }
    }
}

}  // extern "C"
