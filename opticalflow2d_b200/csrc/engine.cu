// engine.cu -- the device-resident iteration engine: one refine pass of the reference's driver loops
// (ImageRegistrationOpticalFlow.cpp:97-151, ImageRegistrationDemons.cpp:86-137,
// ImageRegistrationFluid.cpp:67-142) for `batch` independent image pairs, without a host round trip
// per iteration.
//
//   * every iteration is a fixed sequence of kernels; each kernel starts by reading the pair's control
//     block (engine_ctl.cuh) and returns at once when the pair has converged, so the host enqueues
//     iterations ahead and only polls a device counter every few iterations;
//   * the Logger norms (Logger.cpp:32-51) ride in the epilogue of the kernel that produces the new
//     estimate (the previous estimate is the other ping-pong buffer, so no `prev` copy exists);
//   * break / time-step / regrid / number-of-squarings decisions are taken by the last CTA of the
//     kernel that completes the statistic, in a fixed reduction order (deterministic).
//
// Elastic and Fluid keep their fields in a TRANSPOSED working layout (element (i,j) at i*P + j) for
// the whole loop so that the column-sequential SOR sweep (sor_tile.cuh) streams contiguous memory.
//
// This file is compiled TWICE (build.py): as is (-fmad=false, IEEE division: the "exact" engine, arithmetic level 1) and
// through engine_relaxed.cu (OF2D_RELAXED=1, -fmad=true -prec-div=false: the "relaxed" engine, level 2, where the kernels
// also take algebraically equivalent shortcuts).  engine_dispatch.cu exports the C ABI and forwards to one of the two.
#ifndef OF2D_RELAXED
#define OF2D_RELAXED 0
#endif
#if OF2D_RELAXED
#define ENG(name) of2d_engine_##name##_relaxed
#else
#define ENG(name) of2d_engine_##name##_exact
#endif

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <new>
#include <tuple>
#include <vector>

#include "device_math.cuh"
#include "engine_ctl.cuh"
#include "engine_internal.cuh"

#include "engine_kernels.cuh"
#include "engine_fused.cuh"
#include "engine_fused_tma.cuh"
#include "engine_fluid_tma.cuh"
#include "engine_hs_tma.cuh"
#include "sor_tile.cuh"

// =================================================================================================
// host side
// =================================================================================================
namespace {
struct Engine {
    of2d_engine_head head;       // first member: engine_dispatch.cu reads which build owns the object
    of2d_ctx *ctx;
    of2d_engine_desc d;
    bool dbl, transposed;
    int P;
    size_t n, nT, elem;          // elem = sizeof(real)
    // device buffers (real-typed; vec2 fields have 2*elem per element)
    void *aux, *gradI, *It, *est[2], *c[2], *lvl[2], *vel[2], *incr;
    PairCtl *d_ctl;
    int *d_nactive;
    double *d_partials;
    size_t pstride;
    TraceDev tr;
    void *d_taps[2];             // [fluid kernel, diffusion kernel]: doubles followed by reals
    double full_weight[2];
    std::vector<double> h_taps[2];
    int nsq_cap;
    of2d_curvature_plan *plan;
    SorPlan sor;
    // host
    int *h_snap;                 // pinned [2]
    cudaEvent_t ev[2];
    PairCtl *h_ctl;              // pinned [batch]: the control blocks after the last refine pass
    uint64_t iterations_enqueued;
    const void *cur_Imov;
    int last_niter;
#if OF2D_RELAXED
    // tensor maps of the pipelined Demons kernels (engine_fused_tma.cuh; fp32 fields, kernel widths 3 and 5)
    bool tma_ready;
    CUtensorMap tm_est_win[2], tm_est_tile[2], tm_c_tile[2], tm_imov_win, tm_iref_tile;
    const void *tm_iref_ptr;
    bool tma_fluid;              // engine_fluid_tma.cuh: maps of the two estimate buffers and the increment (transposed layout)
    CUtensorMap tm_fl[5], tm_rw_lvl[2], tm_rw_imov, tm_rw_iref;   // k_rt_fl_rewarp: level-motion buffers; Imov window and Iref tile (per refine call)
    const void *tm_rw_imov_ptr, *tm_rw_iref_ptr;
    bool tma_rewarp;
    bool tma_hs;                 // engine_hs_tma.cuh: estimate buffers (36 x 36 halo tile), gradI, It (34-row halo tiles)
    CUtensorMap tm_hs[4];
#endif
};
#define of2d_engine Engine

#define TRY(x) do { int _s = (x); if (_s) return _s; } while (0)

template <class R>
EngK<R> make_k(of2d_engine *E) {
    EngK<R> K;
    K.nx = E->d.dimx; K.ny = E->d.dimy; K.batch = E->d.batch; K.P = E->P;
    K.n = E->n; K.nT = E->nT;
    K.ctl = E->d_ctl; K.n_active = E->d_nactive; K.partials = E->d_partials; K.pstride = E->pstride; K.tr = E->tr;
    for (int k = 0; k < 2; k++) { K.est[k] = (vec2_t<R> *)E->est[k]; K.c[k] = (vec2_t<R> *)E->c[k]; K.lvl[k] = (vec2_t<R> *)E->lvl[k]; }
    K.ext = nullptr;
    return K;
}

// grid = (CTAs per pair, batch): exactly ONE wave of the kernel -- as many CTAs as are resident at once
// (occupancy of this kernel x number of SMs), never more than there are tiles.  With a single wave every CTA
// walks ntiles / grid tiles (+-1), so no SM waits for a partially filled last wave, and the per-CTA fixed costs
// (control-block load, reduction partial) are paid once per resident slot.
inline int ctas_per_sm(const void *kernel, size_t smem) {
    static int forced = -1;
    if (forced < 0) { const char *e = getenv("OF2D_CTAS_PER_SM"); forced = e && atoi(e) > 0 ? atoi(e) : 0; }
    if (forced) return forced;
    // occupancy depends on the kernel's attributes on the CURRENT device: keyed by (device, kernel, smem), guarded
    static std::map<std::tuple<int, const void *, size_t>, int> cache;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple(dev, kernel, smem);
    const auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, TX * TY, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 4; }
    cache[key] = occ;
    return occ;
}
template <class KernelT>
inline dim3 grid_tiles(const of2d_engine *E, KernelT kernel, size_t smem = 0) {
    const int ntiles = ceil_div(E->d.dimx, TILE) * ceil_div(E->d.dimy, TILE);
    // batches: the CTAs of all pairs together must not exceed the resident slots (a second, partly filled wave would double the time)
    static int use_ceil = -1;
    if (use_ceil < 0) { const char *e = getenv("OF2D_GRID_CEIL"); use_ceil = e && atoi(e) != 0 ? 1 : 0; }
    const long slots = (long)E->ctx->sm_count * ctas_per_sm((const void *)kernel, smem);
    int per_pair = use_ceil ? ceil_div(slots, E->d.batch) : (int)(slots / E->d.batch);
    if (per_pair > ntiles) per_pair = ntiles;
    if (per_pair < 1) per_pair = 1;
    return dim3(per_pair, E->d.batch);
}

template <class R>
ConvW<R> conv_weights(of2d_engine *E, int which) {
    ConvW<R> W;
    memset(&W, 0, sizeof(W));
    const int kw = E->d.kernel_w;
    W.kw = kw;
    W.taps_d = (const double *)E->d_taps[which];
    for (int t = 0; t < kw * kw; t++) W.w[t] = (R)E->h_taps[which][(size_t)t];
    W.full_weight = E->full_weight[which];
    W.neg_zero = -0.0f;
#if OF2D_RELAXED
    {   // rank-1 test in double: w[ii + jj kw] == w[ii + c kw] w[c + jj kw] / w[c + c kw] to 1e-6 of the largest tap
        const std::vector<double> &h = E->h_taps[which];
        const int c = (kw - 1) / 2;
        const double wc = h[(size_t)(c + c * kw)];
        double worst = 0.0, big = 0.0;
        W.separable = (kw & 1) && wc != 0.0 && W.full_weight != 0.0;
        if (W.separable) {
            for (int jj = 0; jj < kw; jj++)
                for (int ii = 0; ii < kw; ii++) {
                    const double w = h[(size_t)(ii + jj * kw)], prod = h[(size_t)(ii + c * kw)] * h[(size_t)(c + jj * kw)] / wc;
                    worst = fmax(worst, fabs(w - prod)); big = fmax(big, fabs(w));
                }
            W.separable = worst <= 1e-6 * big;
        }
        if (W.separable)
            for (int t = 0; t < kw; t++) {
                W.sx[t] = (R)(h[(size_t)(t + c * kw)] / wc);
                W.sy[t] = (R)(h[(size_t)(c + t * kw)] / W.full_weight);
            }
    }
#endif
    return W;
}

template <class R, int EPI, int KW>
int launch_conv_kw(of2d_engine *E, const EngK<R> &K, int src, int dst, int which) {
    const ConvW<R> W = conv_weights<R>(E, which);
    const int cx = (W.kw - 1) / 2;
    const int shift = sizeof(vec2_t<R>) == 8 ? (cx & 1) : 0;
    const size_t smem = 2 * sizeof(vec2_t<R>) * (size_t)(TILE + 2 * cx) * (size_t)((TILE + 2 * cx + shift + 1) & ~1);
    TRY(of2d_ensure_dynamic_smem((const void *)k_e_conv<R, EPI, KW>, smem));
    { ProfScope _ps(E->ctx, EPI == 1 ? "conv_logger" : EPI == 2 ? "conv_maxabs" : "conv"); pdl_launch(k_e_conv<R, EPI, KW>, grid_tiles(E, k_e_conv<R, EPI, KW>, smem), dim3(TX, TY), smem, E->ctx->stream, K, src, dst, W, E->nsq_cap); }
    OF2D_LAUNCH_CHECK(E->ctx);
    return OF2D_SUCCESS;
}

template <class R, int EPI>
int launch_conv(of2d_engine *E, const EngK<R> &K, int src, int dst, int which) {
    switch (E->d.kernel_w) {
        case 3: return launch_conv_kw<R, EPI, 3>(E, K, src, dst, which);
        case 5: return launch_conv_kw<R, EPI, 5>(E, K, src, dst, which);
        case 7: return launch_conv_kw<R, EPI, 7>(E, K, src, dst, which);
        case 9: return launch_conv_kw<R, EPI, 9>(E, K, src, dst, which);
        default: return launch_conv_kw<R, EPI, 0>(E, K, src, dst, which);
    }
}


#if OF2D_RELAXED
// Demons in two fused kernels (engine_fused.cuh): kernel widths 3 and 5 (the default) whose taps are separable (Gaussians are)
// OF2D_FUSED: 0 off, 1 both kernels (default), 2 only the force + smoothing kernel, 3 only the compose + smoothing kernel;
// OF2D_FUSED_NOFAST=1: every tile takes the general (border) path -- experiments only
inline int fused_demons_mode() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("OF2D_FUSED"); on = e ? atoi(e) : 1; if (on < 0 || on > 3) on = 1; }
    return on;
}
inline int fused_nofast() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("OF2D_FUSED_NOFAST"); v = e && atoi(e) != 0 ? 1 : 0; }
    return v;
}
// OF2D_FUSED_TMA: 3 (default) both fused kernels as tensor-map pipelines (engine_fused_tma.cuh), 1 / 2 only the first / second, 0 neither
inline int fused_tma_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("OF2D_FUSED_TMA"); v = e ? (atoi(e) & 3) : 3; }
    return v;
}
template <class R, int KW>
int enqueue_fused_demons_kw(of2d_engine *E, const EngK<R> &K, const R *d_Iref, const ConvW<R> &Wf, const ConvW<R> &Wd) {
    cudaStream_t s = E->ctx->stream;
    const dim3 b(TX, TY);
    const of2d_engine_desc &d = E->d;
    const R si = (R)d.sigma_i, sx = (R)d.sigma_x;
    const R sratio = (si * si) / (sx * sx);
    // fp32: the kernels hold their arrays in static shared memory
    const size_t sm1 = sizeof(R) == 4 ? 0 : fused_smem_force<R, KW>(), sm2 = sizeof(R) == 4 ? 0 : fused_smem_compose<R, KW>();
    const int mode = fused_demons_mode(), nf = fused_nofast();
    // the pipelined instances (engine_fused_tma.cuh): fp32 fields whose layout meets TMA's alignment rules
    bool tma_force = false, tma_compose = false;
    if constexpr (sizeof(R) == 4) {
        if (mode == 1 && !nf && fused_tma_enabled() && E->tma_ready) {
            if (E->tm_iref_ptr != (const void *)d_Iref)
                E->tm_iref_ptr = make_field_map(&E->tm_iref_tile, d_Iref, 4, d.dimx, d.dimy, d.batch, RT_TW, RtGeom<KW>::CW) ? (const void *)d_Iref : nullptr;
            tma_force = E->tm_iref_ptr != nullptr && (fused_tma_enabled() & 1);
            tma_compose = (fused_tma_enabled() & 2) != 0;
        }
    }
    const R sxsq = sx * sx;
    int ex = 0;
    const R inv_sxsq = (sxsq > 0 && frexp((double)sxsq, &ex) == 0.5 && ex > -100 && ex < 100) ? (R)ldexp(1.0, 1 - ex) : (R)0;
    if (mode == 3) {   // unfused first half: force -> C0, smoothing C0 -> C1
        { ProfScope _ps(E->ctx, "demons_force"); pdl_launch(k_e_demons_force<R>, grid_tiles(E, k_e_demons_force<R>), b, 0, s, K, d_Iref, (const R *)E->aux, si * si, sxsq, inv_sxsq); }
        OF2D_LAUNCH_CHECK(E->ctx);
        if (d.method == 3) TRY((launch_conv<R, 0>(E, K, B_C0, B_C1, 0))); else TRY((launch_conv<R, 2>(E, K, B_C0, B_C1, 0)));
    } else if (tma_force) {
        if constexpr (sizeof(R) == 4) {
            using RG = RtGeom<KW>;
            TmaMaps4 MF;
            MF.m[0] = E->tm_est_tile[0]; MF.m[1] = E->tm_est_tile[1]; MF.m[2] = E->tm_imov_win; MF.m[3] = E->tm_iref_tile;
            if (d.method == 3) {
                TRY(of2d_ensure_dynamic_smem((const void *)k_rt_force_conv<0, KW>, RG::F_SMEM));
                { ProfScope _ps(E->ctx, "force_conv"); pdl_launch(k_rt_force_conv<0, KW>, grid_tiles(E, k_rt_force_conv<0, KW>, RG::F_SMEM), b, RG::F_SMEM, s, K, d_Iref, (const R *)E->aux, sratio, Wf, (int)B_C1, E->nsq_cap, MF); }
            } else {
                TRY(of2d_ensure_dynamic_smem((const void *)k_rt_force_conv<2, KW>, RG::F_SMEM));
                { ProfScope _ps(E->ctx, "force_conv_maxabs"); pdl_launch(k_rt_force_conv<2, KW>, grid_tiles(E, k_rt_force_conv<2, KW>, RG::F_SMEM), b, RG::F_SMEM, s, K, d_Iref, (const R *)E->aux, sratio, Wf, (int)B_C1, E->nsq_cap, MF); }
            }
            OF2D_LAUNCH_CHECK(E->ctx);
        }
    } else if (d.method == 3) {
        TRY(of2d_ensure_dynamic_smem((const void *)k_rx_force_conv<R, 0, KW>, sm1));
        { ProfScope _ps(E->ctx, "force_conv"); pdl_launch(k_rx_force_conv<R, 0, KW>, grid_tiles(E, k_rx_force_conv<R, 0, KW>, sm1), b, sm1, s, K, d_Iref, (const R *)E->aux, sratio, Wf, (int)B_C1, E->nsq_cap, nf); }
        OF2D_LAUNCH_CHECK(E->ctx);
    } else {
        TRY(of2d_ensure_dynamic_smem((const void *)k_rx_force_conv<R, 2, KW>, sm1));
        { ProfScope _ps(E->ctx, "force_conv_maxabs"); pdl_launch(k_rx_force_conv<R, 2, KW>, grid_tiles(E, k_rx_force_conv<R, 2, KW>, sm1), b, sm1, s, K, d_Iref, (const R *)E->aux, sratio, Wf, (int)B_C1, E->nsq_cap, nf); }
        OF2D_LAUNCH_CHECK(E->ctx);
    }
    if (d.method == 4)
        for (int q = 0; q < E->nsq_cap; q++) {
            { ProfScope _ps(E->ctx, "square"); pdl_launch(k_e_square<R>, grid_tiles(E, k_e_square<R>), b, 0, s, K, q); }
            OF2D_LAUNCH_CHECK(E->ctx);
        }
    if (mode == 2) {   // unfused second half
        if (d.method == 3) {
            { ProfScope _pc(E->ctx, "compose"); pdl_launch(k_e_compose<R, false>, grid_tiles(E, k_e_compose<R, false>), b, 0, s, K, G_ACTIVE, B_EST_CUR, B_C1, B_C0, d.accumulation == 1); }
            OF2D_LAUNCH_CHECK(E->ctx);
            TRY((launch_conv<R, 1>(E, K, B_C0, B_EST_NEXT, 1)));
        } else {
            { ProfScope _pc(E->ctx, "compose"); pdl_launch(k_e_compose<R, false>, grid_tiles(E, k_e_compose<R, false>), b, 0, s, K, G_ACTIVE, B_EST_CUR, B_CRES, B_CTMP, 0); }
            OF2D_LAUNCH_CHECK(E->ctx);
            TRY((launch_conv<R, 1>(E, K, B_CTMP, B_EST_NEXT, 1)));
        }
        return OF2D_SUCCESS;
    }
    if (tma_compose) {
        if constexpr (sizeof(R) == 4) {
            using RG = RtGeom<KW>;
            TmaMaps4 MC;
            MC.m[0] = E->tm_est_win[0]; MC.m[1] = E->tm_est_win[1]; MC.m[2] = E->tm_c_tile[0]; MC.m[3] = E->tm_c_tile[1];
            TRY(of2d_ensure_dynamic_smem((const void *)k_rt_compose_conv<KW>, RG::C_SMEM));
            { ProfScope _pc(E->ctx, "compose_conv_logger"); pdl_launch(k_rt_compose_conv<KW>, grid_tiles(E, k_rt_compose_conv<KW>, RG::C_SMEM), b, RG::C_SMEM, s, K, d.method == 3 ? (int)B_C1 : (int)B_CRES, d.method == 3 && d.accumulation == 1 ? 1 : 0, Wd, MC); }
            OF2D_LAUNCH_CHECK(E->ctx);
        }
        return OF2D_SUCCESS;
    }
    TRY(of2d_ensure_dynamic_smem((const void *)k_rx_compose_conv<R, KW>, sm2));
    { ProfScope _pc(E->ctx, "compose_conv_logger"); pdl_launch(k_rx_compose_conv<R, KW>, grid_tiles(E, k_rx_compose_conv<R, KW>, sm2), b, sm2, s, K, d.method == 3 ? (int)B_C1 : (int)B_CRES, d.method == 3 && d.accumulation == 1 ? 1 : 0, Wd, nf); }
    OF2D_LAUNCH_CHECK(E->ctx);
    return OF2D_SUCCESS;
}
// returns 1 when the iteration was enqueued, 0 when the unfused sequence has to run, < 0 on error
template <class R>
int enqueue_fused_demons(of2d_engine *E, const EngK<R> &K, const R *d_Iref) {
    if (!fused_demons_mode()) return 0;
    const int kw = E->d.kernel_w;
    if ((kw != 3 && kw != 5) || E->n >= (size_t)0x7fff0000u) return 0;   // wider kernels / larger fields: the unfused sequence
    const ConvW<R> Wf = conv_weights<R>(E, 0), Wd = conv_weights<R>(E, 1);
    if (!Wf.separable || !Wd.separable) return 0;
    int st = kw == 3 ? enqueue_fused_demons_kw<R, 3>(E, K, d_Iref, Wf, Wd) : enqueue_fused_demons_kw<R, 5>(E, K, d_Iref, Wf, Wd);
    return st ? -st : 1;
}
#endif

// one iteration of method `m`, enqueued on the context's stream
template <class R>
int enqueue_iteration(of2d_engine *E, const EngK<R> &K, const R *d_Iref, int curv_flags = 0) {
    cudaStream_t s = E->ctx->stream;
    const dim3 b(TX, TY);
    const of2d_engine_desc &d = E->d;
    switch (d.method) {
        case 0: {
            const R alpha = (R)d.alpha;
            { ProfScope _ps(E->ctx, "hs_iter"); pdl_launch(k_hs_iter<R>, grid_tiles(E, k_hs_iter<R>), b, 0, s, K, (const vec2_t<R> *)E->gradI, (const R *)E->It, alpha * alpha, 0); }
            OF2D_LAUNCH_CHECK(E->ctx);
            break;
        }
        case 1:
            TRY(of2d_curvature_engine_step(E->plan, K.ctl, K.n_active, K.partials, K.pstride, K.tr, E->est[0], E->est[1], E->gradI, E->It, curv_flags));
            break;
        case 2:
            TRY(sor_tile_launch<R>(E->ctx, E->sor, K.ctl, K.n_active, K.partials, K.pstride, K.tr, 0, (vec2_t<R> *)E->est[0], (vec2_t<R> *)E->est[1], nullptr, nullptr,
                                   (const vec2_t<R> *)E->gradI, (const R *)E->It));
            break;
        case 3:
        case 4: {
#if OF2D_RELAXED
            { const int f = enqueue_fused_demons<R>(E, K, d_Iref); if (f < 0) return -f; if (f > 0) break; }
#endif
            const R si = (R)d.sigma_i, sx = (R)d.sigma_x;
            const R sxsq = sx * sx;
            int ex = 0;
            const R inv_sxsq = (sxsq > 0 && frexp((double)sxsq, &ex) == 0.5 && ex > -100 && ex < 100) ? (R)ldexp(1.0, 1 - ex) : (R)0;   // exact power of two only
            { ProfScope _ps(E->ctx, "demons_force"); pdl_launch(k_e_demons_force<R>, grid_tiles(E, k_e_demons_force<R>), b, 0, s, K, d_Iref, (const R *)E->aux, si * si, sxsq, inv_sxsq); }
            OF2D_LAUNCH_CHECK(E->ctx);
            if (d.method == 3) {
                TRY((launch_conv<R, 0>(E, K, B_C0, B_C1, 0)));
                { ProfScope _pc(E->ctx, "compose"); pdl_launch(k_e_compose<R, false>, grid_tiles(E, k_e_compose<R, false>), b, 0, s, K, G_ACTIVE, B_EST_CUR, B_C1, B_C0, d.accumulation == 1); }
                OF2D_LAUNCH_CHECK(E->ctx);
                TRY((launch_conv<R, 1>(E, K, B_C0, B_EST_NEXT, 1)));
            } else {
                TRY((launch_conv<R, 2>(E, K, B_C0, B_C1, 0)));
                for (int q = 0; q < E->nsq_cap; q++) {
                    { ProfScope _ps(E->ctx, "square"); pdl_launch(k_e_square<R>, grid_tiles(E, k_e_square<R>), b, 0, s, K, q); }
                    OF2D_LAUNCH_CHECK(E->ctx);
                }
                { ProfScope _pc(E->ctx, "compose"); pdl_launch(k_e_compose<R, false>, grid_tiles(E, k_e_compose<R, false>), b, 0, s, K, G_ACTIVE, B_EST_CUR, B_CRES, B_CTMP, 0); }
                OF2D_LAUNCH_CHECK(E->ctx);
                TRY((launch_conv<R, 1>(E, K, B_CTMP, B_EST_NEXT, 1)));
            }
            break;
        }
        case 5: {
            TRY(sor_tile_launch<R>(E->ctx, E->sor, K.ctl, K.n_active, K.partials, K.pstride, K.tr, 1, (vec2_t<R> *)E->vel[0], (vec2_t<R> *)E->vel[1],
                                   (const vec2_t<R> *)E->est[0], (const vec2_t<R> *)E->est[1], (const vec2_t<R> *)E->gradI, (const R *)E->It, (vec2_t<R> *)E->incr));
            // (the increment of the new velocity and the time step are produced by the sweep kernel itself)
            bool integ_done = false;
#if OF2D_RELAXED
            if constexpr (sizeof(R) == 4) {
                if (E->tma_fluid && (fused_tma_enabled() & 1)) {
                    TmaMaps5 MI;
                    for (int q = 0; q < 5; q++) MI.m[q] = E->tm_fl[q];
                    TRY(of2d_ensure_dynamic_smem((const void *)k_rt_fl_integrate, FI_SMEM));
                    { ProfScope _ps(E->ctx, "fluid_integrate"); pdl_launch(k_rt_fl_integrate, grid_tiles(E, k_rt_fl_integrate, FI_SMEM), b, FI_SMEM, s, K, (const float2 *)E->incr, MI); }
                    OF2D_LAUNCH_CHECK(E->ctx);
                    integ_done = true;
                }
            }
#endif
            if (!integ_done) {
                { ProfScope _ps(E->ctx, "fluid_integrate"); pdl_launch(k_fl_integrate<R>, grid_tiles(E, k_fl_integrate<R>), b, 0, s, K, (const vec2_t<R> *)E->incr); }
                OF2D_LAUNCH_CHECK(E->ctx);
            }
            // regrid (ImageRegistrationFluid.cpp:108-124): level <- est + level o (id + est); est <- 0; re-warp; derivatives
            { ProfScope _pc(E->ctx, "regrid_compose"); pdl_launch(k_e_compose<R, true>, grid_tiles(E, k_e_compose<R, true>), b, 0, s, K, G_REGRID, B_LVL_CUR, B_EST_CUR, B_LVL_NEXT, 0); }
            OF2D_LAUNCH_CHECK(E->ctx);
            bool rewarp_done = false;
#if OF2D_RELAXED
            if constexpr (sizeof(R) == 4) {
                if (E->tma_rewarp && (fused_tma_enabled() & 2)) {
                    if (E->tm_rw_imov_ptr != E->cur_Imov)
                        E->tm_rw_imov_ptr = make_field_map(&E->tm_rw_imov, E->cur_Imov, 4, d.dimx, d.dimy, d.batch, FW, FW) ? E->cur_Imov : nullptr;
                    if (E->tm_rw_iref_ptr != (const void *)d_Iref)
                        E->tm_rw_iref_ptr = make_field_map(&E->tm_rw_iref, d_Iref, 4, d.dimx, d.dimy, d.batch, TILE, TILE) ? (const void *)d_Iref : nullptr;
                    if (E->tm_rw_imov_ptr && E->tm_rw_iref_ptr) {
                        TmaMaps4 MR;
                        MR.m[0] = E->tm_rw_lvl[0]; MR.m[1] = E->tm_rw_lvl[1]; MR.m[2] = E->tm_rw_imov; MR.m[3] = E->tm_rw_iref;
                        TRY(of2d_ensure_dynamic_smem((const void *)k_rt_fl_rewarp, RW_SMEM));
                        { ProfScope _pw(E->ctx, "regrid_rewarp"); pdl_launch(k_rt_fl_rewarp, grid_tiles(E, k_rt_fl_rewarp, RW_SMEM), b, RW_SMEM, s, K, (int)G_REGRID, d_Iref, (const float *)E->cur_Imov, (float2 *)E->gradI, (float *)E->It, (int)B_EST_NEXT, MR); }
                        OF2D_LAUNCH_CHECK(E->ctx);
                        rewarp_done = true;
                    }
                }
            }
#endif
            if (!rewarp_done) {
                { ProfScope _pw(E->ctx, "regrid_rewarp"); pdl_launch(k_fl_rewarp<R>, grid_tiles(E, k_fl_rewarp<R>), b, 0, s, K, G_REGRID, d_Iref, (const R *)E->cur_Imov, B_LVL_NEXT, (vec2_t<R> *)E->gradI, (R *)E->It, B_EST_NEXT); }
                OF2D_LAUNCH_CHECK(E->ctx);
            }
            pdl_launch(k_regrid_commit, ceil_div(K.batch, 128), 128, 0, s, K.ctl, K.batch);
            OF2D_LAUNCH_CHECK(E->ctx);
            break;
        }
        default:
            of2d_set_error("engine: unknown method %d", d.method);
            return OF2D_ERR_INVALID;
    }
    return OF2D_SUCCESS;
}


// Diffusion: two Jacobi steps in one launch (temporal blocking, k_hs_pair)
template <class R>
int enqueue_hs_pair(of2d_engine *E, const EngK<R> &K) {
    cudaStream_t s = E->ctx->stream;
    const dim3 b(TX, TY);
    const R alpha = (R)E->d.alpha;
#if OF2D_RELAXED
    if constexpr (sizeof(R) == 4) {
        if (E->tma_hs && (fused_tma_enabled() & 1)) {
            TmaMaps4 MH;
            for (int q = 0; q < 4; q++) MH.m[q] = E->tm_hs[q];
            TRY(of2d_ensure_dynamic_smem((const void *)k_rt_hs_pair, HP_SMEM));
            { ProfScope _ps(E->ctx, "hs_pair"); pdl_launch(k_rt_hs_pair, grid_tiles(E, k_rt_hs_pair, HP_SMEM), b, HP_SMEM, s, K, alpha * alpha, MH); }
            OF2D_LAUNCH_CHECK(E->ctx);
            return OF2D_SUCCESS;
        }
    }
#endif
    { ProfScope _ps(E->ctx, "hs_pair"); pdl_launch(k_hs_pair<R>, grid_tiles(E, k_hs_pair<R>), b, 0, s, K, (const vec2_t<R> *)E->gradI, (const R *)E->It, alpha * alpha); }
    OF2D_LAUNCH_CHECK(E->ctx);
    return OF2D_SUCCESS;
}
inline bool hs_pair_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("OF2D_HS_PAIR"); on = e ? atoi(e) != 0 : 1; }
    return on != 0;
}

template <class R>
int refine_impl(of2d_engine *E, const R *d_Iref, const R *d_Imov, R *d_motion, int niter) {
    of2d_ctx *ctx = E->ctx;
    cudaStream_t s = ctx->stream;
    EngK<R> K = make_k<R>(E);
    K.ext = (vec2_t<R> *)d_motion;
    const of2d_engine_desc &d = E->d;
    const dim3 b(TX, TY);
    const size_t vbytes = sizeof(vec2_t<R>) * E->n * K.batch, vbytesT = sizeof(vec2_t<R>) * E->nT * K.batch;
    if (niter > E->tr.cap) { of2d_set_error("engine: niter %d above the trace capacity %d", niter, E->tr.cap); return OF2D_ERR_INVALID; }
    E->cur_Imov = d_Imov;
    E->last_niter = niter;

    // ---- set-up: Iaux = Imov o (id + motion); derivatives; estimate = 0 (e.g. ImageRegistrationDemons.cpp:97-106)
    OF2D_CUDA_TRY(cudaMemcpyAsync(E->lvl[0], d_motion, vbytes, cudaMemcpyDeviceToDevice, s));
    pdl_launch(k_ctl_begin, ceil_div(K.batch, 128), 128, 0, s, K.ctl, K.batch, niter, K.n_active);
    OF2D_LAUNCH_CHECK(ctx);
    pdl_launch(k_e_warp<R>, grid_tiles(E, k_e_warp<R>), b, 0, s, K, G_NONE, d_Imov, B_LVL_CUR, (R *)E->aux);
    OF2D_LAUNCH_CHECK(ctx);
    if (d.method != 3 && d.method != 4) {
        pdl_launch(k_e_derivatives<R>, grid_tiles(E, k_e_derivatives<R>), b, 0, s, K, G_NONE, d_Iref, (const R *)E->aux, (vec2_t<R> *)E->gradI, (R *)E->It, E->transposed ? 1 : 0);
        OF2D_LAUNCH_CHECK(ctx);
    }
    const size_t eb = E->transposed ? vbytesT : vbytes;
    OF2D_CUDA_TRY(cudaMemsetAsync(E->est[0], 0, eb, s));
    OF2D_CUDA_TRY(cudaMemsetAsync(E->est[1], 0, eb, s));

    // ---- iterations, enqueued ahead in chunks; the device counter of running pairs is polled without stalling
    const int chunk = 8;
    int enq = 0, slot = 0;
    bool pending[2] = {false, false};
    bool done = niter <= 0;
    while (!done && enq < niter) {
        const int m = niter - enq < chunk ? niter - enq : chunk;
        for (int q = 0; q < m;) {
            if (d.method == 0 && m - q >= 2 && hs_pair_enabled()) { TRY(enqueue_hs_pair<R>(E, K)); q += 2; }
            else {
                // Curvature (register-blocked path): iteration k's inverse row pass runs iteration k+1's forward row pass
                int cf = 0;
                if (d.method == 1 && of2d_curvature_plan_fuses_rows(E->plan)) cf = (enq + q > 0 ? OF2D_CURV_SKIP_FWD : 0) | (enq + q + 1 < niter ? OF2D_CURV_FUSE_NEXT : 0);
                TRY(enqueue_iteration<R>(E, K, d_Iref, cf));
                q += 1;
            }
        }
        enq += m;
        E->iterations_enqueued += (uint64_t)m;
        pdl_launch(k_snapshot_active, dim3(1), dim3(32), 0, s, (const int *)E->d_nactive, &E->h_snap[slot]);   // (not a memcpy: see the kernel)
        OF2D_CUDA_TRY(cudaGetLastError());
        OF2D_CUDA_TRY(cudaEventRecord(E->ev[slot], s));
        pending[slot] = true;
        const int other = slot ^ 1;
        if (pending[other]) {   // at most two chunks ahead of the device
            OF2D_CUDA_TRY(cudaEventSynchronize(E->ev[other]));
            pending[other] = false;
            if (E->h_snap[other] == 0) done = true;
        }
        slot = other;
    }

    // a two-step launch whose break test fired on its first step hands that step to the next launch; the closing single step
    // serves a redo raised by the last two-step launch (and the iteration a pair is short of when the two reductions disagree
    // on the test, the error being within rounding of 0.001); an empty launch otherwise
    if (d.method == 0 && niter >= 2 && hs_pair_enabled()) {
        const R alpha = (R)d.alpha;
        { ProfScope _ps(E->ctx, "hs_iter"); pdl_launch(k_hs_iter<R>, grid_tiles(E, k_hs_iter<R>), b, 0, s, K, (const vec2_t<R> *)E->gradI, (const R *)E->It, alpha * alpha, 2); }
        OF2D_LAUNCH_CHECK(ctx);
    }

    // ---- tear-down: motion <- estimate + motion o (id + estimate); the estimate is dropped (:136-137)
    if (E->transposed) {   // the estimate lives in the transposed working layout: its tiles are turned in shared memory
        { ProfScope _pc(E->ctx, "final_compose"); pdl_launch(k_e_compose<R, true>, grid_tiles(E, k_e_compose<R, true>), b, 0, s, K, G_NONE, B_LVL_CUR, B_EST_CUR, B_EXT, 0); }
    } else {
        { ProfScope _pc(E->ctx, "final_compose"); pdl_launch(k_e_compose<R, false>, grid_tiles(E, k_e_compose<R, false>), b, 0, s, K, G_NONE, B_LVL_CUR, B_EST_CUR, B_EXT, 0); }
    }
    OF2D_LAUNCH_CHECK(ctx);
    pdl_launch(k_copy_ctl_to_host, dim3(ceil_div((long)K.batch * (long)(sizeof(PairCtl) / sizeof(int4)), 256)), dim3(256), 0, s, (const PairCtl *)E->d_ctl, E->h_ctl, K.batch);
    OF2D_CUDA_TRY(cudaGetLastError());
    OF2D_CUDA_TRY(cudaStreamSynchronize(s));
    unsigned flags = 0;
    int overflow = 0;
    for (int p = 0; p < K.batch; p++) { flags |= E->h_ctl[p].flags; overflow |= E->h_ctl[p].overflow; }
    if (flags & OF2D_FLAG_DIVZERO) {
        for (int p = 0; p < K.batch; p++) {
            if (E->h_ctl[p].flags) {
                E->h_ctl[p].flags = 0;   // host copy keeps the iteration counts; device word is reset
                OF2D_CUDA_TRY(cudaMemsetAsync(&E->d_ctl[p].flags, 0, sizeof(unsigned), s));
            }
        }
        of2d_set_error("Divide by zero exception");
        return OF2D_ERR_DIVZERO;
    }
    if (overflow) { of2d_set_error("engine: scaling-and-squaring needed more than %d squarings", E->nsq_cap); return OF2D_ERR_UNSUPPORTED; }
    return OF2D_SUCCESS;
}

int alloc(void **p, size_t bytes) {
    *p = nullptr;
    OF2D_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 16));
    OF2D_CUDA_TRY(cudaMemset(*p, 0, bytes ? bytes : 16));
    return OF2D_SUCCESS;
}

int upload_taps(of2d_engine *E, int which, const double *h_kernel) {
    const int kw = E->d.kernel_w, nt = kw * kw;
    std::vector<unsigned char> buf((sizeof(double) + E->elem) * (size_t)nt);
    double *dd = (double *)buf.data();
    for (int t = 0; t < nt; t++) dd[t] = h_kernel[t];
    if (E->dbl) { double *r = (double *)(buf.data() + sizeof(double) * nt); for (int t = 0; t < nt; t++) r[t] = h_kernel[t]; }
    else { float *r = (float *)(buf.data() + sizeof(double) * nt); for (int t = 0; t < nt; t++) r[t] = (float)h_kernel[t]; }
    const int cx = (kw - 1) / 2;
    double full = 0.0;   // visiting order of Field.tpp:242-243
    for (int ii = -cx; ii <= cx; ii++)
        for (int jj = -cx; jj <= cx; jj++) full += h_kernel[(ii + cx) + (jj + cx) * kw];
    E->full_weight[which] = full;
    E->h_taps[which].assign(h_kernel, h_kernel + nt);
    TRY(alloc(&E->d_taps[which], buf.size()));
    OF2D_CUDA_TRY(cudaMemcpy(E->d_taps[which], buf.data(), buf.size(), cudaMemcpyHostToDevice));
    return OF2D_SUCCESS;
}

}  // namespace

extern "C" {

int ENG(create)(of2d_ctx *ctx, const of2d_engine_desc *desc, of2d_engine_head **out_head) {
    Engine **out = reinterpret_cast<Engine **>(out_head);
    *out = nullptr;
    OF2D_REQUIRE(desc && desc->dimx > 1 && desc->dimy > 1 && desc->batch > 0 && desc->method >= 0 && desc->method <= 5 && desc->max_iter >= 0, "bad engine description");
    OF2D_CUDA_TRY(cudaSetDevice(ctx->device));
    of2d_engine *E = new (std::nothrow) of2d_engine();
    OF2D_REQUIRE(E, "out of host memory");
    E->head.relaxed = OF2D_RELAXED;
    E->ctx = ctx; E->d = *desc; E->dbl = desc->real_is_double != 0;
    E->elem = E->dbl ? 8 : 4;
    const int nx = desc->dimx, ny = desc->dimy, B = desc->batch, m = desc->method;
    E->n = (size_t)nx * ny;
    E->transposed = (m == 2 || m == 5);
    int st = OF2D_SUCCESS;
    auto fail = [&](int code) { ENG(destroy)(&E->head); return code; };
    if (E->transposed) {
        E->sor = sor_plan(nx, ny, B, desc->mu, desc->lambda, desc->omega, E->dbl, m == 5, ctx->sm_count);
        if (!E->sor.supported) {
            of2d_set_error("engine: SOR parameters (mu %g, lambda %g, omega %g) do not contract fast enough for the tiled sweep", desc->mu, desc->lambda, desc->omega);
            return fail(OF2D_ERR_UNSUPPORTED);
        }
        E->P = E->sor.P; E->nT = E->sor.nT;
    } else {
        E->P = 0; E->nT = E->n;
    }
    const size_t v = 2 * E->elem, sc = E->elem;
    const size_t nN = E->n * B, nTt = E->nT * B;
    const size_t nE = E->transposed ? nTt : nN;
    if ((st = alloc(&E->aux, sc * nN))) return fail(st);
    if (m != 3 && m != 4) {
        if ((st = alloc(&E->gradI, v * nE))) return fail(st);
        if ((st = alloc(&E->It, sc * nE))) return fail(st);
    }
    for (int k = 0; k < 2; k++) {
        if ((st = alloc(&E->est[k], v * nE))) return fail(st);
        if ((st = alloc(&E->lvl[k], v * nN))) return fail(st);
    }
    if (m == 3 || m == 4) {
        for (int k = 0; k < 2; k++) if ((st = alloc(&E->c[k], v * nN))) return fail(st);
    }
    if (m == 5) {
        for (int k = 0; k < 2; k++) if ((st = alloc(&E->vel[k], v * nTt))) return fail(st);
        if ((st = alloc(&E->incr, v * nTt))) return fail(st);
    }
    if ((st = alloc((void **)&E->d_ctl, sizeof(PairCtl) * B))) return fail(st);
    if ((st = alloc((void **)&E->d_nactive, sizeof(int)))) return fail(st);
    // partials: up to 4 values per CTA of the widest grid (32 x 32 tiles, rows of the curvature pass, SOR tiles)
    size_t nblk = (size_t)ceil_div(nx, TILE) * ceil_div(ny, TILE);
    if ((size_t)ny > nblk) nblk = ny;
    if (E->transposed && (size_t)E->sor.nbands * E->sor.nstrips > nblk) nblk = (size_t)E->sor.nbands * E->sor.nstrips;
    E->pstride = nblk * 4 + 8;
    if ((st = alloc((void **)&E->d_partials, sizeof(double) * E->pstride * B))) return fail(st);
    const int cap = desc->max_iter > 0 ? desc->max_iter : 1;
    E->tr.cap = cap;
    const size_t tn = (size_t)cap * B;
    if ((st = alloc((void **)&E->tr.err, sizeof(double) * tn))) return fail(st);
    if ((st = alloc((void **)&E->tr.maxabs, sizeof(double) * tn))) return fail(st);
    if ((st = alloc((void **)&E->tr.dt, sizeof(double) * tn))) return fail(st);
    if ((st = alloc((void **)&E->tr.minjac, sizeof(double) * tn))) return fail(st);
    if ((st = alloc((void **)&E->tr.regrid, sizeof(int) * tn))) return fail(st);
    if ((st = alloc((void **)&E->tr.nsq, sizeof(int) * tn))) return fail(st);
    if (m == 3 || m == 4) {
        if (!(desc->kernel_w > 0 && desc->kernel_w <= kConvMaxW && desc->kernel_fluid && desc->kernel_diffusion)) {
            of2d_set_error("engine: demons needs two Gaussian kernels of width <= %d", kConvMaxW);
            return fail(OF2D_ERR_UNSUPPORTED);
        }
        if ((st = upload_taps(E, 0, desc->kernel_fluid))) return fail(st);
        if ((st = upload_taps(E, 1, desc->kernel_diffusion))) return fail(st);
        // |c| <= sigma_x / (2 sigma_i) (Demons.cpp:57) and smoothing is a convex combination, so
        // maxabs = sqrt(2) max|c.y| bounds the number of squarings Motion::exp can ask for
        E->nsq_cap = 0;
        if (m == 4) {
            const double bound = sqrt(2.0) * fabs(desc->sigma_x) / (2.0 * fabs(desc->sigma_i)) * 1.0001;
            int cap2 = bound > 0 ? (int)ceil(1.0 + log2(bound)) : 0;
            if (cap2 < 0) cap2 = 0;
            E->nsq_cap = cap2 + 1;
            if (E->nsq_cap > 16) { of2d_set_error("engine: sigma_x / sigma_i too large for scaling and squaring"); return fail(OF2D_ERR_UNSUPPORTED); }
        }
    }
#if OF2D_RELAXED
    E->tma_ready = false;
    E->tm_iref_ptr = nullptr;
    if ((m == 3 || m == 4) && !E->dbl && (desc->kernel_w == 3 || desc->kernel_w == 5)) {
        const int cx = (desc->kernel_w - 1) / 2, cw = TILE + 2 * cx, ww = cw + 2;
        bool ok = true;
        for (int k = 0; k < 2 && ok; k++)
            ok = make_field_map(&E->tm_est_win[k], E->est[k], 8, nx, ny, B, FW, FW) && make_field_map(&E->tm_est_tile[k], E->est[k], 8, nx, ny, B, RT_TW, ww) &&
                 make_field_map(&E->tm_c_tile[k], E->c[k], 8, nx, ny, B, RT_VP, cw);
        ok = ok && make_field_map(&E->tm_imov_win, E->aux, 4, nx, ny, B, FW, FW);
        E->tma_ready = ok;
    }
    E->tma_fluid = false;
    E->tma_rewarp = false;
    E->tma_hs = false;
    if (m == 0 && !E->dbl)
        E->tma_hs = make_field_map(&E->tm_hs[0], E->est[0], 8, nx, ny, B, HP_H0, HP_H0) && make_field_map(&E->tm_hs[1], E->est[1], 8, nx, ny, B, HP_H0, HP_H0) &&
                    make_field_map(&E->tm_hs[2], E->gradI, 8, nx, ny, B, HP_GP, HP_H1) && make_field_map(&E->tm_hs[3], E->It, 4, nx, ny, B, HP_TP, HP_H1);
    if (m == 5 && !E->dbl) {   // transposed layout: P elements per line, nx lines per pair, nT elements per pair
        const void *bufs3[3] = {E->est[0], E->est[1], E->incr};
        bool ok = (E->nT % 2) == 0 && (E->P % 2) == 0;
        for (int k = 0; k < 3 && ok; k++) ok = make_field_map_pitched(&E->tm_fl[k], bufs3[k], 8, E->P, nx, (size_t)E->nT, B, FI_W, FI_H);
        for (int k = 0; k < 2 && ok; k++) ok = make_field_map_pitched(&E->tm_fl[3 + k], E->est[k], 8, E->P, nx, (size_t)E->nT, B, TILE, TILE);
        E->tma_fluid = ok;
        bool okr = true;
        for (int k = 0; k < 2 && okr; k++) okr = make_field_map(&E->tm_rw_lvl[k], E->lvl[k], 8, nx, ny, B, RW_LP, RW_H);
        E->tma_rewarp = okr && (nx % 4) == 0;   // (the Imov / Iref maps are made per refine call: caller-owned arrays)
    }
    E->tm_rw_imov_ptr = E->tm_rw_iref_ptr = nullptr;
#endif
    if (m == 1) {
        if ((st = of2d_curvature_plan_create(ctx, nx, ny, desc->alpha, desc->tau, E->dbl ? 1 : 0, &E->plan))) return fail(st);
        if ((st = of2d_curvature_plan_set_batch(E->plan, B))) return fail(st);
#if OF2D_RELAXED
        if ((st = of2d_curvature_plan_set_relaxed(E->plan, 1))) return fail(st);
#endif
    }
    if (cudaHostAlloc((void **)&E->h_snap, sizeof(int) * 2, cudaHostAllocDefault) != cudaSuccess) { of2d_set_error("engine: pinned allocation failed"); return fail(OF2D_ERR_CUDA); }
    for (int k = 0; k < 2; k++)
        if (cudaEventCreateWithFlags(&E->ev[k], cudaEventDisableTiming) != cudaSuccess) { of2d_set_error("engine: event creation failed"); return fail(OF2D_ERR_CUDA); }
    if (cudaHostAlloc((void **)&E->h_ctl, sizeof(PairCtl) * (size_t)B, cudaHostAllocDefault) != cudaSuccess) { of2d_set_error("engine: pinned allocation failed"); return fail(OF2D_ERR_CUDA); }
    memset(E->h_ctl, 0, sizeof(PairCtl) * (size_t)B);
    *out = E;
    return OF2D_SUCCESS;
}

void ENG(destroy)(of2d_engine_head *H) {
    Engine *E = reinterpret_cast<Engine *>(H);
    if (!E) return;
    cudaSetDevice(E->ctx->device);
    cudaStreamSynchronize(E->ctx->stream);
    void *bufs[] = {E->aux, E->gradI, E->It, E->est[0], E->est[1], E->c[0], E->c[1], E->lvl[0], E->lvl[1], E->vel[0], E->vel[1], E->incr,
                    E->d_ctl, E->d_nactive, E->d_partials, E->tr.err, E->tr.maxabs, E->tr.dt, E->tr.minjac, E->tr.regrid, E->tr.nsq, E->d_taps[0], E->d_taps[1]};
    for (void *p : bufs) if (p) cudaFree(p);
    if (E->plan) of2d_curvature_plan_destroy(E->plan);
    if (E->h_snap) cudaFreeHost(E->h_snap);
    if (E->h_ctl) cudaFreeHost(E->h_ctl);
    for (int k = 0; k < 2; k++) if (E->ev[k]) cudaEventDestroy(E->ev[k]);
    delete E;
}

int ENG(reset_state)(of2d_engine_head *H) {
    Engine *E = reinterpret_cast<Engine *>(H);
    if (E->d.method != 5) return OF2D_SUCCESS;
    const size_t bytes = 2 * E->elem * E->nT * E->d.batch;
    OF2D_CUDA_TRY(cudaMemsetAsync(E->vel[0], 0, bytes, E->ctx->stream));
    OF2D_CUDA_TRY(cudaMemsetAsync(E->vel[1], 0, bytes, E->ctx->stream));
    return OF2D_SUCCESS;
}

int ENG(refine_f32)(of2d_engine_head *H, const float *d_Iref, const float *d_Imov, float *d_motion, int niter) {
    Engine *E = reinterpret_cast<Engine *>(H);
    OF2D_REQUIRE(!E->dbl, "engine was created for double fields");
    return refine_impl<float>(E, d_Iref, d_Imov, d_motion, niter);
}
int ENG(refine_f64)(of2d_engine_head *H, const double *d_Iref, const double *d_Imov, double *d_motion, int niter) {
    Engine *E = reinterpret_cast<Engine *>(H);
    OF2D_REQUIRE(E->dbl, "engine was created for float fields");
    return refine_impl<double>(E, d_Iref, d_Imov, d_motion, niter);
}

int ENG(pair_result)(of2d_engine_head *H, int pair, int *iterations, int *nregrid, double *last_err) {
    Engine *E = reinterpret_cast<Engine *>(H);
    OF2D_REQUIRE(pair >= 0 && pair < E->d.batch, "pair out of range");
    const PairCtl &c = E->h_ctl[(size_t)pair];
    if (iterations) *iterations = c.iter;
    if (nregrid) *nregrid = c.nregrid;
    if (last_err) *last_err = c.err;
    return OF2D_SUCCESS;
}

int ENG(trace)(of2d_engine_head *H, int pair, int which, double *h_out, int count) {
    Engine *E = reinterpret_cast<Engine *>(H);
    OF2D_REQUIRE(pair >= 0 && pair < E->d.batch && which >= 0 && which <= 5 && count >= 0 && count <= E->tr.cap, "bad trace request");
    if (count == 0) return OF2D_SUCCESS;
    const size_t off = (size_t)pair * E->tr.cap;
    if (which <= 3) {
        const double *src = which == 0 ? E->tr.err : which == 1 ? E->tr.maxabs : which == 2 ? E->tr.dt : E->tr.minjac;
        OF2D_CUDA_TRY(cudaMemcpyAsync(h_out, src + off, sizeof(double) * count, cudaMemcpyDeviceToHost, E->ctx->stream));
        OF2D_CUDA_TRY(cudaStreamSynchronize(E->ctx->stream));
    } else {
        std::vector<int> tmp((size_t)count);
        const int *src = which == 4 ? E->tr.regrid : E->tr.nsq;
        OF2D_CUDA_TRY(cudaMemcpyAsync(tmp.data(), src + off, sizeof(int) * count, cudaMemcpyDeviceToHost, E->ctx->stream));
        OF2D_CUDA_TRY(cudaStreamSynchronize(E->ctx->stream));
        for (int k = 0; k < count; k++) h_out[k] = (double)tmp[(size_t)k];
    }
    return OF2D_SUCCESS;
}

uint64_t ENG(iterations_enqueued)(of2d_engine_head *H) { return reinterpret_cast<Engine *>(H)->iterations_enqueued; }

}  // extern "C"
