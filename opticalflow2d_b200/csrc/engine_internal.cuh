// engine_internal.cuh -- declarations shared between engine.cu and dct.cu (not part of the C ABI)
#pragma once

#include "common.cuh"
#include "engine_ctl.cuh"

// one curvature iteration inside the engine loop: u = est[sel], result -> est[sel ^ 1], Logger epilogue
// in the last kernel (dct.cu)
// flags: the register-blocked path (line lengths 512 .. 4096) can run the forward row pass of iteration k+1 inside the
// inverse row pass of iteration k; the caller (the engine's iteration loop) says which iterations are chained
enum { OF2D_CURV_SKIP_FWD = 1, OF2D_CURV_FUSE_NEXT = 2 };
int of2d_curvature_engine_step(of2d_curvature_plan *plan, PairCtl *ctl, int *n_active, double *partials, size_t pstride, TraceDev tr, void *est0, void *est1,
                               const void *gradI, const void *It, int flags);
int of2d_curvature_plan_fuses_rows(const of2d_curvature_plan *plan);   // 1 if the plan takes that path (else flags are ignored: pass 0)
int of2d_curvature_plan_set_batch(of2d_curvature_plan *plan, int batch);
