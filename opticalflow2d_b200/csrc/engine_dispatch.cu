// engine_dispatch.cu -- the public of2d_engine_* entry points (include/of2d_cuda.h): an engine is created by the build
// that matches the context's arithmetic level at creation time (1: exact, 2: relaxed; of2d_ctx_set_fast_math) and every
// later call goes to the build that owns the object.
#include "engine_internal.cuh"

#define FWD(call) (reinterpret_cast<of2d_engine_head *>(engine)->relaxed ? of2d_engine_##call##_relaxed : of2d_engine_##call##_exact)

extern "C" {

int of2d_engine_create(of2d_ctx *ctx, const of2d_engine_desc *desc, of2d_engine **out) {
    of2d_engine_head **o = reinterpret_cast<of2d_engine_head **>(out);
    return ctx->fast_math >= 2 ? of2d_engine_create_relaxed(ctx, desc, o) : of2d_engine_create_exact(ctx, desc, o);
}
void of2d_engine_destroy(of2d_engine *engine) {
    if (engine) FWD(destroy)(reinterpret_cast<of2d_engine_head *>(engine));
}
int of2d_engine_reset_state(of2d_engine *engine) { return FWD(reset_state)(reinterpret_cast<of2d_engine_head *>(engine)); }
int of2d_engine_refine_f32(of2d_engine *engine, const float *d_Iref, const float *d_Imov, float *d_motion, int niter) {
    return FWD(refine_f32)(reinterpret_cast<of2d_engine_head *>(engine), d_Iref, d_Imov, d_motion, niter);
}
int of2d_engine_refine_f64(of2d_engine *engine, const double *d_Iref, const double *d_Imov, double *d_motion, int niter) {
    return FWD(refine_f64)(reinterpret_cast<of2d_engine_head *>(engine), d_Iref, d_Imov, d_motion, niter);
}
int of2d_engine_pair_result(of2d_engine *engine, int pair, int *iterations, int *nregrid, double *last_err) {
    return FWD(pair_result)(reinterpret_cast<of2d_engine_head *>(engine), pair, iterations, nregrid, last_err);
}
int of2d_engine_trace(of2d_engine *engine, int pair, int which, double *h_out, int count) {
    return FWD(trace)(reinterpret_cast<of2d_engine_head *>(engine), pair, which, h_out, count);
}
uint64_t of2d_engine_iterations_enqueued(of2d_engine *engine) { return FWD(iterations_enqueued)(reinterpret_cast<of2d_engine_head *>(engine)); }
int of2d_engine_is_relaxed(of2d_engine *engine) { return reinterpret_cast<of2d_engine_head *>(engine)->relaxed; }

}  // extern "C"
