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
int of2d_curvature_plan_set_relaxed(of2d_curvature_plan *plan, int on);   // fp32 fields on the register path: single-precision transform and spectrum

// ---- the two builds of the iteration engine (engine.cu compiled as is, and through engine_relaxed.cu) ---------------
// Both define the same object behind an `of2d_engine_head`; engine_dispatch.cu exports the public of2d_engine_* entry
// points of include/of2d_cuda.h and forwards to the build the object belongs to.
struct of2d_engine_head { int relaxed; };
#define OF2D_ENGINE_BUILD_DECLS(SUF)                                                                                                          \
    extern "C" {                                                                                                                              \
    int of2d_engine_create_##SUF(of2d_ctx *ctx, const of2d_engine_desc *desc, of2d_engine_head **out);                                          \
    void of2d_engine_destroy_##SUF(of2d_engine_head *engine);                                                                                 \
    int of2d_engine_reset_state_##SUF(of2d_engine_head *engine);                                                                              \
    int of2d_engine_refine_f32_##SUF(of2d_engine_head *engine, const float *d_Iref, const float *d_Imov, float *d_motion, int niter);         \
    int of2d_engine_refine_f64_##SUF(of2d_engine_head *engine, const double *d_Iref, const double *d_Imov, double *d_motion, int niter);      \
    int of2d_engine_pair_result_##SUF(of2d_engine_head *engine, int pair, int *iterations, int *nregrid, double *last_err);                   \
    int of2d_engine_trace_##SUF(of2d_engine_head *engine, int pair, int which, double *h_out, int count);                                     \
    uint64_t of2d_engine_iterations_enqueued_##SUF(of2d_engine_head *engine);                                                                 \
    }
OF2D_ENGINE_BUILD_DECLS(exact)
OF2D_ENGINE_BUILD_DECLS(relaxed)
