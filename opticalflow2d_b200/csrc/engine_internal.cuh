// engine_internal.cuh -- declarations shared between engine.cu and dct.cu (not part of the C ABI)
#pragma once

#include "common.cuh"
#include "engine_ctl.cuh"

// one curvature iteration inside the engine loop: u = est[sel], result -> est[sel ^ 1], Logger epilogue
// in the last kernel (dct.cu)
int of2d_curvature_engine_step(of2d_curvature_plan *plan, PairCtl *ctl, int *n_active, double *partials, size_t pstride, TraceDev tr, void *est0, void *est1,
                               const void *gradI, const void *It);
int of2d_curvature_plan_set_batch(of2d_curvature_plan *plan, int batch);
