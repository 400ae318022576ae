// engine_relaxed.cu -- the iteration engine at arithmetic level 2 ("relaxed"): engine.cu compiled a second time with
// OF2D_RELAXED=1 and -fmad=true -prec-div=false -prec-sqrt=false (build.py gives *_relaxed.cu these flags).  Same kernels,
// same control protocol; the arithmetic may contract a*b+c, divide approximately and take the shortcuts marked
// `#if OF2D_RELAXED` in engine_kernels.cuh / sor_tile.cuh.  Results are held to the north-star tolerances (1e-3 px and 1e-4
// relative SSD in fp32, 1e-6 px in fp64) by tests/test_relaxed_gpu.py and tests/test_configs_gpu.py, not to bit-identity.
#define OF2D_RELAXED 1
#include "engine.cu"
