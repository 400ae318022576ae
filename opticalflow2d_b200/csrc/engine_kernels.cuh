// engine_kernels.cuh -- kernels of the iteration engine (included by engine.cu only).
//
// Shape shared by all of them: grid = (CTAs per pair, batch).  A CTA of 32 x 8 threads walks 32 x 32
// pixel tiles of ITS pair in a grid-stride loop (x fastest, so neighbouring CTAs stream neighbouring
// memory), reads the pair's control block once (one 64-byte load) and publishes ONE partial per
// reduction, so the fixed cost of the device-side control protocol is paid per CTA, not per tile.
// Arithmetic is written in the reference's operation order and the library is compiled with
// -fmad=false, so every field value is bit-identical to the per-step (strict) path; only the reduction
// order of the norms differs (double partial sums instead of one sequential float accumulator).
#pragma once

#include "device_math.cuh"
#include "engine_ctl.cuh"
#include "tma.cuh"

namespace {

#ifndef OF2D_HS_MINB
#define OF2D_HS_MINB 3      // k_hs_pair, fp32: resident CTAs per SM the register allocation aims at
#endif
#ifndef OF2D_INTEG_MINB
#define OF2D_INTEG_MINB 1   // k_fl_integrate (1: no register cap)
#endif
constexpr int TX = 32, TY = 8, PY = 4;
constexpr int TILE = 32;

enum Buf { B_C0 = 0, B_C1 = 1, B_EST_CUR = 2, B_EST_NEXT = 3, B_CRES = 4, B_CTMP = 5, B_LVL_CUR = 6, B_LVL_NEXT = 7, B_EXT = 9 };
enum Gate { G_NONE = 0, G_ACTIVE = 1, G_REGRID = 2 };

template <class R>
struct EngK {
    int nx, ny, batch, P;          // P: pitch of the transposed layout (0 when unused)
    size_t n;                      // nx * ny
    size_t nT;                     // elements per pair in the transposed layout
    PairCtl *ctl;
    int *n_active;
    double *partials;
    size_t pstride;                // doubles per pair in `partials`
    TraceDev tr;
    vec2_t<R> *est[2];
    vec2_t<R> *c[2];
    vec2_t<R> *lvl[2];
    vec2_t<R> *ext;                // caller-provided field (set per launch)
};

// the integer head of PairCtl in registers (one L2 round trip)
struct CtlHot {
    int active, iter, niter;
    unsigned flags;
    int sel, regrid, skip, nsquares;
    int nregrid, msel, overflow, vsel;
    int prev_other, redo;
};
__device__ __forceinline__ CtlHot load_ctl(const PairCtl *c) {
    const int4 *p = reinterpret_cast<const int4 *>(c);
    const int4 a = __ldcg(p), b = __ldcg(p + 1), d = __ldcg(p + 2), e = __ldcg(p + 3);
    CtlHot h;
    h.active = a.x; h.iter = a.y; h.niter = a.z; h.flags = (unsigned)a.w;
    h.sel = b.x; h.regrid = b.y; h.skip = b.z; h.nsquares = b.w;
    h.nregrid = d.x; h.msel = d.y; h.overflow = d.z; h.vsel = d.w;
    h.prev_other = e.x; h.redo = e.y;
    return h;
}
static_assert(offsetof(PairCtl, sel) == 16 && offsetof(PairCtl, nregrid) == 32 && offsetof(PairCtl, prev_other) == 48 && offsetof(PairCtl, redo) == 52, "CtlHot mirrors the head of PairCtl");

template <class R>
__device__ __forceinline__ vec2_t<R> *pick(const EngK<R> &K, int which, const CtlHot &h, int pair, bool transposed = false) {
    const size_t off = (size_t)pair * (transposed ? K.nT : K.n);
    switch (which) {
        case B_C0: return K.c[0] + off;
        case B_C1: return K.c[1] + off;
        case B_EST_CUR: return K.est[h.sel] + off;
        case B_EST_NEXT: return K.est[h.sel ^ 1] + off;
        case B_CRES: return K.c[(h.nsquares & 1) ? 0 : 1] + off;
        case B_CTMP: return K.c[(h.nsquares & 1) ? 1 : 0] + off;
        case B_LVL_CUR: return K.lvl[h.msel] + off;
        case B_LVL_NEXT: return K.lvl[h.msel ^ 1] + off;
        default: return K.ext + off;
    }
}

__device__ __forceinline__ bool gate_open(const CtlHot &h, int gate) {
    if (gate == G_ACTIVE) return h.active != 0;
    if (gate == G_REGRID) return h.regrid != 0;
    return true;
}

// Motion::norm addend (Motion.cpp:45).  The reference takes the square root in double of float data;
// the fp32 build of the engine takes it in float (1 ulp of a term that is then averaged over all pixels).
#if OF2D_RELAXED   // subnormal squares (differences below 1e-19 px) are flushed to zero: no scaling fix-up around MUFU.SQRT
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
__device__ __forceinline__ float norm_term_f(float2 v) { return sqrt_approx(v.x * v.x + v.y * v.y); }
__device__ __forceinline__ double norm_term(float2 v) { return (double)norm_term_f(v); }
__device__ __forceinline__ double norm_term(double2 v) { return sqrt(v.x * v.x + v.y * v.y); }

// packed fp32 pairs (sm_100a FFMA2 / FADD2): two lanes per issue slot, each lane IEEE round-to-nearest
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// per-thread accumulation of the Logger terms: in the field precision within a tile, flushed into double per tile
template <class R> struct NormAcc {
    R sd = 0, sp = 0;
    double dsd = 0.0, dsp = 0.0;
    __device__ __forceinline__ void add(vec2_t<R> nv, vec2_t<R> old) {
        const vec2_t<R> df = mk2<R>(nv.x - old.x, nv.y - old.y);
        if (sizeof(R) == 4) { sd += (R)norm_term_f(make_float2((float)df.x, (float)df.y)); sp += (R)norm_term_f(make_float2((float)old.x, (float)old.y)); }
        else { sd += (R)norm_term(df); sp += (R)norm_term(old); }
    }
    __device__ __forceinline__ void flush() { dsd += (double)sd; dsp += (double)sp; sd = 0; sp = 0; }
};

struct TileWalk {
    int tiles_x, ntiles;
    unsigned magic;   // ceil(2^32 / tiles_x): tile / tiles_x == umulhi(tile, magic) while tile * tiles_x < 2^32
    __device__ __forceinline__ TileWalk(int fast_extent, int slow_extent) {
        tiles_x = (fast_extent + TILE - 1) / TILE;
        ntiles = tiles_x * ((slow_extent + TILE - 1) / TILE);
        magic = tiles_x > 1 ? 0xFFFFFFFFu / (unsigned)tiles_x + 1u : 0u;
    }
    __device__ __forceinline__ int ty(int tile) const { return tiles_x > 1 ? (int)__umulhi((unsigned)tile, magic) : tile; }
    __device__ __forceinline__ int tx(int tile) const { return tile - ty(tile) * tiles_x; }
};

// Logger epilogue shared by the kernels that produce the next estimate
template <class R>
__device__ __forceinline__ void logger_epilogue(const EngK<R> &K, PairCtl *c, int pair, double sd, double sp, bool clear_redo = false) {
    block_sum2(sd, sp);
    const double vals[2] = {sd, sp};
    double *part = K.partials + (size_t)pair * K.pstride;
    if (publish_partials<2>(vals, part, &c->ticket[0], gridDim.x, blockIdx.x)) {
        double out[2];
        reduce_partials<2>(part, gridDim.x, out, 0u, 0u);
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            if (clear_redo) c->redo = 0;
            c->sel ^= 1;
            finalize_logger<R>(c, K.tr, pair, out[0], out[1], (unsigned)K.n, K.n_active);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// control
// ---------------------------------------------------------------------------------------------
__global__ void k_ctl_begin(PairCtl *ctl, int batch, int niter, int *n_active) {
    pdl_enter();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) *n_active = niter > 0 ? batch : 0;
    if (p >= batch) return;
    PairCtl *c = ctl + p;
    c->active = niter > 0;
    c->iter = 0;
    c->niter = niter;
    c->sel = 0;
    c->regrid = 0;
    c->skip = 0;
    c->nsquares = 0;
    c->nregrid = 0;
    c->msel = 0;
    c->overflow = 0;
    c->prev_other = 0;
    c->redo = 0;
    for (int k = 0; k < 4; k++) c->ticket[k] = 0u;
    c->err = 0.0;
}

// Read-backs of the control state into PINNED host memory by a kernel (the host pointer is the device pointer under unified addressing)
// instead of cudaMemcpyAsync: a device -> host memcpy queues on the copy engine behind whatever another stream is downloading (a
// wave's motions, 512 MB), and the iteration loop -- which polls the counter of running pairs every 8 iterations -- stood still
// for the length of that download (11 % of the streamed batch rate, ~1 ms per method of the pipelined sessions call).
__global__ void k_snapshot_active(const int *__restrict__ n_active, int *host_slot) {
    pdl_enter();
    if (threadIdx.x == 0 && blockIdx.x == 0) { *host_slot = *n_active; __threadfence_system(); }
}
__global__ void k_copy_ctl_to_host(const PairCtl *__restrict__ ctl, PairCtl *host_ctl, int batch) {
    pdl_enter();
    const int n4 = batch * (int)(sizeof(PairCtl) / sizeof(int4));
    const int4 *src = reinterpret_cast<const int4 *>(ctl);
    int4 *dst = reinterpret_cast<int4 *>(host_ctl);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += gridDim.x * blockDim.x) dst[k] = __ldcg(src + k);
    __threadfence_system();
}

__global__ void k_regrid_commit(PairCtl *ctl, int batch) {
    pdl_enter();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    PairCtl *c = ctl + p;
    if (!c->regrid) return;
    c->msel ^= 1;
    c->sel ^= 1;          // the zeroed buffer becomes the running estimate; the other one is Logger's prev
    c->prev_other = 1;
    c->regrid = 0;
    c->nregrid += 1;
}

// ---------------------------------------------------------------------------------------------
// Horn-Schunck Jacobi step + Logger (OpticalFlowDiffusion.cpp:19-84, Logger.cpp:32-51)
// ---------------------------------------------------------------------------------------------
// fixup: 0 = a plain iteration (skipped while a redo of k_hs_pair is pending); 1 = runs only to redo the first step of a two-step
// launch; 2 = the closing launch of a refine pass: a plain iteration that also serves a pending redo
template <class R>
__global__ void __launch_bounds__(TX *TY) k_hs_iter(EngK<R> K, const vec2_t<R> *__restrict__ gradI_all, const R *__restrict__ It_all, R alphasq, int fixup) {
    pdl_enter();
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active || (fixup == 0 && h.redo) || (fixup == 1 && !h.redo)) return;   // fixup 2: whether or not a redo is pending
    const int nx = K.nx, ny = K.ny;
    const vec2_t<R> *__restrict__ u = pick(K, B_EST_CUR, h, pair);
    vec2_t<R> *__restrict__ un = pick(K, B_EST_NEXT, h, pair);
    const vec2_t<R> *__restrict__ gradI = gradI_all + (size_t)pair * K.n;
    const R *__restrict__ It = It_all + (size_t)pair * K.n;
    const TileWalk T(nx, ny);
    NormAcc<R> acc;
    bool divzero = false;
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        if (i0 >= 1 && i0 + TILE < nx && j0 >= 1 && j0 + TILE < ny) {
            // interior tile: no edge cases.  A thread owns 4 consecutive rows of one column, so the column of u it
            // needs (rows j-1 .. j+4) is loaded once and every address is a row pointer plus an immediate offset.
            const size_t idx0 = (size_t)(i0 + threadIdx.x) + (size_t)(j0 + 4 * threadIdx.y) * nx;
            const vec2_t<R> *__restrict__ up = u + idx0;
            vec2_t<R> ce[6], le[4], ri[4], dI[4];
            R it[4];
#pragma unroll
            for (int r = 0; r < 6; r++) ce[r] = up[(ptrdiff_t)(r - 1) * nx];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                le[q] = up[(ptrdiff_t)q * nx - 1];
                ri[q] = up[(ptrdiff_t)q * nx + 1];
                dI[q] = gradI[idx0 + (size_t)q * nx];
                it[q] = It[idx0 + (size_t)q * nx];
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const vec2_t<R> qm = mk2<R>((((le[q].x + ri[q].x) + ce[q].x) + ce[q + 2].x) / (R)4.0f, (((le[q].y + ri[q].y) + ce[q].y) + ce[q + 2].y) / (R)4.0f);   // gradients.h:78
                const vec2_t<R> f = lssd_force<R>(dI[q], it[q], qm);
                const R den = alphasq + dI[q].x * dI[q].x + dI[q].y * dI[q].y;
                vec2_t<R> o;
                if (den == 0) { divzero = true; o = qm; }
                else o = mk2<R>(qm.x - f.x / den, qm.y - f.y / den);
                un[idx0 + (size_t)q * nx] = o;
                acc.add(o, ce[q + 1]);
            }
            acc.flush();
            continue;
        }
        const int i = i0 + threadIdx.x;
        const int jb = j0 + threadIdx.y;
        // all loads of the thread's 4 pixels are issued before any arithmetic: indices are clamped into the
        // field so that they are unconditional (the clamped values are never used)
        const int ic = min(i, nx - 1);
        vec2_t<R> old[PY], a[PY], b[PY], cc[PY], d[PY], dI[PY];
        R it[PY];
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jc = min(jb + p * TY, ny - 1);
            const int idx = ic + jc * nx;
            old[p] = u[idx];
            a[p] = u[idx - (ic > 0)];
            b[p] = u[idx + (ic < nx - 1)];
            cc[p] = u[idx - (jc > 0 ? nx : 0)];
            d[p] = u[idx + (jc < ny - 1 ? nx : 0)];
            dI[p] = gradI[idx];
            it[p] = It[idx];
        }
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int j = jb + p * TY;
            vec2_t<R> q;
            if (i == 0 || i >= nx - 1 || j == 0 || j >= ny - 1) q = mk2<R>((R)0.0f, (R)0.0f);
            else q = mk2<R>((((a[p].x + b[p].x) + cc[p].x) + d[p].x) / (R)4.0f, (((a[p].y + b[p].y) + cc[p].y) + d[p].y) / (R)4.0f);   // gradients.h:78
            const vec2_t<R> f = lssd_force<R>(dI[p], it[p], q);
            const R den = alphasq + dI[p].x * dI[p].x + dI[p].y * dI[p].y;
            vec2_t<R> o;
            if (den == 0) { divzero = divzero || (i < nx && j < ny); o = q; }
            else o = mk2<R>(q.x - f.x / den, q.y - f.y / den);
            if (i < nx && j < ny) {
                un[i + j * nx] = o;
                acc.add(o, old[p]);
            }
        }
        acc.flush();
    }
    if (divzero) atomicOr(&c->flags, OF2D_FLAG_DIVZERO);
    logger_epilogue<R>(K, c, pair, acc.dsd, acc.dsp, fixup != 0);
}

// ---------------------------------------------------------------------------------------------
// Two Horn-Schunck Jacobi steps per launch: temporal blocking in shared memory.  A CTA stages u^k on the 36 x 36 halo
// of its 32 x 32 tile, evaluates u^(k+1) on the 34 x 34 halo (its own pixels plus the 132 ring points) into shared
// memory, then u^(k+2) on the tile: u, gradI, It are read once and u written once per TWO iterations.  Every pixel
// value is the same expression as in k_hs_iter (bit-identical); the Logger sums of both steps are reduced together and
// the last CTA replays the driver's loop for the two iterations (Logger.cpp:32-51, ImageRegistrationOpticalFlow.cpp:131-134).
// If the break test fires after the FIRST step the launch leaves the state untouched and raises `redo`: the next launch
// (this kernel again, or the closing k_hs_iter of the refine pass) finds the flag and redoes that one step from the
// intact u^k as a single-step iteration.
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ vec2_t<R> hs_point(vec2_t<R> a, vec2_t<R> b, vec2_t<R> cc, vec2_t<R> d, bool border, vec2_t<R> dI, R it, R alphasq, bool &divzero) {
    vec2_t<R> q;
    if (border) q = mk2<R>((R)0.0f, (R)0.0f);
    else q = mk2<R>((((a.x + b.x) + cc.x) + d.x) / (R)4.0f, (((a.y + b.y) + cc.y) + d.y) / (R)4.0f);   // gradients.h:78
    const vec2_t<R> f = lssd_force<R>(dI, it, q);
    const R den = alphasq + dI.x * dI.x + dI.y * dI.y;
    if (den == 0) { divzero = true; return q; }
    return mk2<R>(q.x - f.x / den, q.y - f.y / den);
}

// end of a two-step Horn-Schunck launch: the Logger test of BOTH steps by the last CTA (the first step's break hands the step to the
// next launch through `redo`)
template <class R>
__device__ __forceinline__ void hs_pair_epilogue(const EngK<R> &K, PairCtl *c, int pair, const NormAcc<R> &acc1, const NormAcc<R> &acc2, bool single, bool divzero) {
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (divzero) atomicOr(&c->flags, OF2D_FLAG_DIVZERO);
    if (single) { logger_epilogue<R>(K, c, pair, acc1.dsd, acc1.dsp, true); return; }
    double sd1 = acc1.dsd, sp1 = acc1.dsp, sd2 = acc2.dsd, sp2 = acc2.dsp;
    block_sum2(sd1, sp1);
    block_sum2(sd2, sp2);
    const double vals[4] = {sd1, sp1, sd2, sp2};
    double *part = K.partials + (size_t)pair * K.pstride;
    if (publish_partials<4>(vals, part, &c->ticket[0], gridDim.x, blockIdx.x)) {
        double out[4];
        reduce_partials<4>(part, gridDim.x, out, 0u, 0u);
        if (tid == 0) {
            const R n = (R)(unsigned)K.n;
            const R dn = (R)out[0] / n, pn = (R)out[1] / n;                       // Motion.cpp:47
            const R err1 = pn == 0 ? (R)0.0f : dn / pn;                           // Logger.cpp:39
            const int itc = c->iter;
            if ((err1 < (R)0.001f && itc > 1) || itc + 1 >= c->niter) {
                c->redo = 1;   // the loop ends after the first step: state untouched, the single-step launch redoes it
            } else {
                c->err = (double)err1;
                if (itc < K.tr.cap) K.tr.err[(size_t)pair * K.tr.cap + itc] = (double)err1;
                c->iter = itc + 1;
                c->sel ^= 1;
                finalize_logger<R>(c, K.tr, pair, out[2], out[3], (unsigned)K.n, K.n_active);
            }
        }
    }
}

template <class R>
__global__ void __launch_bounds__(TX *TY, sizeof(R) == 4 ? OF2D_HS_MINB : 2) k_hs_pair(EngK<R> K, const vec2_t<R> *__restrict__ gradI_all, const R *__restrict__ It_all, R alphasq) {
    pdl_enter();
    constexpr int H0 = TILE + 4, H1 = TILE + 2, NRING = 4 * H1 - 4;
    __shared__ vec2_t<R> s0[H0 * H0];   // u^k on the 36 x 36 halo tile
    __shared__ vec2_t<R> s1[H1 * H1];   // u^(k+1) on the 34 x 34 halo tile
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const bool single = h.redo != 0;   // redo of the first step of the previous two-step launch: one step only
    const int nx = K.nx, ny = K.ny;
    const vec2_t<R> *__restrict__ u = pick(K, B_EST_CUR, h, pair);
    vec2_t<R> *__restrict__ un = pick(K, B_EST_NEXT, h, pair);
    const vec2_t<R> *__restrict__ gradI = gradI_all + (size_t)pair * K.n;
    const R *__restrict__ It = It_all + (size_t)pair * K.n;
    const int tid = threadIdx.x + threadIdx.y * TX;
    // the ring point of this thread (threads 0 .. 131): coordinates in the 34 x 34 halo tile
    int rr = 0, rc = 0;
    if (tid < H1) { rr = 0; rc = tid; }
    else if (tid < 2 * H1) { rr = H1 - 1; rc = tid - H1; }
    else if (tid < 2 * H1 + TILE) { rr = tid - 2 * H1 + 1; rc = 0; }
    else { rr = tid - (2 * H1 + TILE) + 1; rc = H1 - 1; }
    const bool has_ring = tid < NRING;
    const TileWalk T(nx, ny);
    NormAcc<R> acc1, acc2;
    bool divzero = false;
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        const int i = i0 + threadIdx.x;
        // loads first: the thread's share of the u^k halo tile, gradI / It of its 4 pixels and of its ring point
        constexpr int NR0 = (H0 * H0 + TX * TY - 1) / (TX * TY);
        vec2_t<R> u0[NR0], dI[PY], dIr;
        R it[PY], itr;
#pragma unroll
        for (int k = 0; k < NR0; k++) {
            const int e = min(tid + k * TX * TY, H0 * H0 - 1);
            const int r = e / H0, cc = e - r * H0;
            const int gi = min(max(i0 - 2 + cc, 0), nx - 1), gj = min(max(j0 - 2 + r, 0), ny - 1);   // clamped: out-of-image halo values are never used
            u0[k] = u[gi + gj * nx];
        }
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int gi = min(i, nx - 1), gj = min(j0 + threadIdx.y + p * TY, ny - 1);
            dI[p] = gradI[gi + gj * nx]; it[p] = It[gi + gj * nx];
        }
        const int ri = i0 - 1 + rc, rj = j0 - 1 + rr;
        const bool ring_in = has_ring && ri >= 0 && ri < nx && rj >= 0 && rj < ny;
        {
            const int gi = min(max(ri, 0), nx - 1), gj = min(max(rj, 0), ny - 1);
            dIr = gradI[gi + gj * nx]; itr = It[gi + gj * nx];
        }
        __syncthreads();   // the previous tile's reads of s0 / s1 are over
#pragma unroll
        for (int k = 0; k < NR0; k++) { const int e = tid + k * TX * TY; if (e < H0 * H0) s0[e] = u0[k]; }
        __syncthreads();
        // first step on the 34 x 34 halo: own pixels, then the ring point
        vec2_t<R> u1[PY];
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY, j = j0 + jl;
            const int e0 = (jl + 2) * H0 + threadIdx.x + 2;
            const bool border = i == 0 || i >= nx - 1 || j == 0 || j >= ny - 1;
            u1[p] = hs_point<R>(s0[e0 - 1], s0[e0 + 1], s0[e0 - H0], s0[e0 + H0], border, dI[p], it[p], alphasq, divzero);
            s1[(jl + 1) * H1 + threadIdx.x + 1] = u1[p];
            if (i < nx && j < ny) {
                acc1.add(u1[p], s0[e0]);
                if (single) un[i + j * nx] = u1[p];
            }
        }
        if (single) { acc1.flush(); continue; }
        if (has_ring) {
            const int e0 = (rr + 1) * H0 + rc + 1;
            const bool border = ri <= 0 || ri >= nx - 1 || rj <= 0 || rj >= ny - 1;
            bool dz = false;
            const vec2_t<R> v = hs_point<R>(s0[e0 - 1], s0[e0 + 1], s0[e0 - H0], s0[e0 + H0], border, dIr, itr, alphasq, dz);
            s1[rr * H1 + rc] = ring_in ? v : mk2<R>((R)0, (R)0);
        }
        acc1.flush();
        __syncthreads();
        // second step on the tile
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY, j = j0 + jl;
            const int e1 = (jl + 1) * H1 + threadIdx.x + 1;
            const bool border = i == 0 || i >= nx - 1 || j == 0 || j >= ny - 1;
            bool dz = false;
            const vec2_t<R> o = hs_point<R>(s1[e1 - 1], s1[e1 + 1], s1[e1 - H1], s1[e1 + H1], border, dI[p], it[p], alphasq, dz);
            if (i < nx && j < ny) {
                un[i + j * nx] = o;
                acc2.add(o, u1[p]);
            }
        }
        acc2.flush();
    }
    hs_pair_epilogue<R>(K, c, pair, acc1, acc2, single, divzero);
}

// ---------------------------------------------------------------------------------------------
// gated primitives of the loops
// ---------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(TX *TY) k_e_warp(EngK<R> K, int gate, const R *__restrict__ src_all, int u_buf, R *__restrict__ dst_all) {
    pdl_enter();
    const int pair = blockIdx.y;
    const CtlHot h = load_ctl(K.ctl + pair);
    if (!gate_open(h, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const R *__restrict__ src = src_all + (size_t)pair * K.n;
    R *__restrict__ dst = dst_all + (size_t)pair * K.n;
    const vec2_t<R> *__restrict__ u = pick(K, u_buf, h, pair);
    const TileWalk T(nx, ny);
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i = T.tx(tile) * TILE + threadIdx.x;
        const int jb = T.ty(tile) * TILE + threadIdx.y;
        if (i >= nx) continue;
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int j = jb + p * TY;
            if (j < ny) {
                const int idx = i + j * nx;
                dst[idx] = warp_pixel<R>(src, nx, ny, i, j, u[idx], src[idx]);
            }
        }
    }
}

// out = v + u o (id + v)   (Motion::accumulate, Motion.cpp:113-178); add_only: out = u + v (Field::operator+=)
// VT: v is stored in the transposed working layout (element (i,j) at i*P + j): its tile is read with j fastest and turned
// in shared memory, so the Fluid regrid composes straight from the running estimate (no untransposed copy)
template <class R, bool VT>
__global__ void __launch_bounds__(TX *TY, sizeof(R) == 4 ? 4 : 1) k_e_compose(EngK<R> K, int gate, int u_buf, int v_buf, int out_buf, int add_only) {
    pdl_enter();
    __shared__ vec2_t<R> sT[VT ? TILE : 1][VT ? TILE + 1 : 1];
    const int pair = blockIdx.y;
    const CtlHot h = load_ctl(K.ctl + pair);
    if (!gate_open(h, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const vec2_t<R> *__restrict__ u = pick(K, u_buf, h, pair);
    const vec2_t<R> *__restrict__ v = pick(K, v_buf, h, pair, VT);
    vec2_t<R> *__restrict__ out = pick(K, out_buf, h, pair);
    const TileWalk T(nx, ny);
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        const int i = i0 + threadIdx.x;
        const int jb = j0 + threadIdx.y;
        if (VT) {
            __syncthreads();
            const int jt = j0 + threadIdx.x;
#pragma unroll
            for (int p = 0; p < PY; p++) {
                const int il = threadIdx.y + p * TY, it = i0 + il;
                if (it < nx && jt < ny) sT[il][threadIdx.x] = v[(size_t)it * K.P + jt];
            }
            __syncthreads();
        }
        auto v_at = [&](int p) -> vec2_t<R> { return VT ? sT[threadIdx.x][threadIdx.y + p * TY] : v[i + (jb + p * TY) * nx]; };
        if (sizeof(R) == 4 && i0 + TILE <= nx && j0 + TILE <= ny) {
            // full tile (fp32; in fp64 the batch costs more registers than it hides latency): the loads of the thread's 4 pixels are batched (4 x {v, u}, then the 16 gathers), so a warp
            // has up to 16 loads in flight instead of 2; same expressions as compose_pixel
            vec2_t<R> vv[PY], uu[PY];
#pragma unroll
            for (int p = 0; p < PY; p++) { const int idx = i + (jb + p * TY) * nx; vv[p] = v_at(p); uu[p] = u[idx]; }
            if (add_only) {
#pragma unroll
                for (int p = 0; p < PY; p++) out[i + (jb + p * TY) * nx] = mk2<R>(uu[p].x + vv[p].x, uu[p].y + vv[p].y);
                continue;
            }
            Bilin<R> bl[PY];
            vec2_t<R> s00[PY], s10[PY], s01[PY], s11[PY];
#pragma unroll
            for (int p = 0; p < PY; p++) {
                const int j = jb + p * TY;
                bl[p] = bilin_setup<R>(i, j, vv[p].x, vv[p].y, nx, ny);
                const BilinTaps<R> t = bilin_taps<R>(bl[p], i + j * nx, nx);
                s00[p] = u[t.o]; s10[p] = u[t.o + t.ox]; s01[p] = u[t.o + t.oy]; s11[p] = u[t.o + t.ox + t.oy];
            }
#pragma unroll
            for (int p = 0; p < PY; p++) out[i + (jb + p * TY) * nx] = compose_taps<R>(bl[p], s00[p], s10[p], s01[p], s11[p], vv[p], uu[p]);
            continue;
        }
        if (i >= nx) continue;
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int j = jb + p * TY;
            if (j < ny) {
                const int idx = i + j * nx;
                const vec2_t<R> vv = v_at(p), uu = u[idx];
                out[idx] = add_only ? mk2<R>(uu.x + vv.x, uu.y + vv.y) : compose_pixel<R>(u, nx, ny, i, j, vv, uu);
            }
        }
    }
}

// one pixel of a squaring step of Motion::exp (Motion.cpp:262-274): w + w o (id + w) at (i, j), w = sc * src (taps from global memory)
template <class R>
__device__ __forceinline__ vec2_t<R> square_pixel(const vec2_t<R> *__restrict__ src, int nx, int ny, int i, int j, R sc) {
    const int idx = i + j * nx;
    vec2_t<R> v = src[idx];
    v.x *= sc; v.y *= sc;
    const Bilin<R> b = bilin_setup<R>(i, j, v.x, v.y, nx, ny);
    vec2_t<R> o = v;   // out of bounds: keeps the (scaled) value
#if OF2D_RELAXED
    if (b.inside && b.hx && b.hy) {   // all four taps inside: weights sum to 1; the power-of-two scale commutes with the interpolation
        const vec2_t<R> s00 = src[b.idxO], s10 = src[b.idxO + 1], s01 = src[b.idxO + nx], s11 = src[b.idxO + nx + 1];
        const R lx = s00.x + b.fx * (s10.x - s00.x), hx = s01.x + b.fx * (s11.x - s01.x);
        const R ly = s00.y + b.fx * (s10.y - s00.y), hy = s01.y + b.fx * (s11.y - s01.y);
        o = mk2<R>(v.x + sc * (lx + b.fy * (hx - lx)), v.y + sc * (ly + b.fy * (hy - ly)));
    } else
#endif
    if (b.inside) {
        const R one = (R)1;
        vec2_t<R> t = src[b.idxO];
        R vx = (t.x * sc) * (one - b.fx) * (one - b.fy), vy = (t.y * sc) * (one - b.fx) * (one - b.fy);
        R weight = (one - b.fx) * (one - b.fy);
        if (b.hx) { t = src[b.idxO + 1]; vx += (t.x * sc) * b.fx * (one - b.fy); vy += (t.y * sc) * b.fx * (one - b.fy); weight += b.fx * (one - b.fy); }
        if (b.hy) { t = src[b.idxO + nx]; vx += (t.x * sc) * (one - b.fx) * b.fy; vy += (t.y * sc) * (one - b.fx) * b.fy; weight += (one - b.fx) * b.fy; }
        if (b.hx && b.hy) { t = src[b.idxO + 1 + nx]; vx += (t.x * sc) * b.fx * b.fy; vy += (t.y * sc) * b.fx * b.fy; weight += b.fx * b.fy; }
        if (weight != 0) o = mk2<R>(v.x + vx / weight, v.y + vy / weight);
    }
    return o;
}

// one squaring of Motion::exp (Motion.cpp:262-274): dst = w + w o (id + w), w = scale * src (scale only at s == 0;
// a power of two, so scaling the taps on the fly is exact)
template <class R>
__global__ void __launch_bounds__(TX *TY) k_e_square(EngK<R> K, int s) {
    pdl_enter();
    const int pair = blockIdx.y;
    const PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active || s >= h.nsquares) return;
    const int nx = K.nx, ny = K.ny;
    const size_t off = (size_t)pair * K.n;
    const vec2_t<R> *__restrict__ src = K.c[(s & 1) ? 0 : 1] + off;
    vec2_t<R> *__restrict__ dst = K.c[(s & 1) ? 1 : 0] + off;
    const R sc = s == 0 ? (R)__ldcg(&c->scale) : (R)1;
    const TileWalk T(nx, ny);
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i = T.tx(tile) * TILE + threadIdx.x;
        const int jb = T.ty(tile) * TILE + threadIdx.y;
        if (i >= nx) continue;
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int j = jb + p * TY;
            if (j >= ny) continue;
            const int idx = i + j * nx;
            const vec2_t<R> o = square_pixel<R>(src, nx, ny, i, j, sc);
            dst[idx] = o;
        }
    }
}

// Demons force: warp + derivatives + demons_iteration (DemonsThirions.cpp:18-27, Demons.cpp:34-63).
// The warped image of a 32 x 32 tile (+1 halo) is evaluated once into shared memory.
template <class R>
// inv_sigma_xsq: 1 / sigma_xsq when sigma_xsq is a power of two (x / 2^k == x * 2^-k exactly, so the multiplication gives the
// reference's bits without the IEEE division sequence), else 0
__global__ void __launch_bounds__(TX *TY, sizeof(R) == 4 ? 8 : 1) k_e_demons_force(EngK<R> K, const R *__restrict__ Iref_all, const R *__restrict__ Imov_all, R sigma_isq, R sigma_xsq, R inv_sigma_xsq) {
    pdl_enter();
    __shared__ R sw[TILE + 2][TILE + 2 + 1];
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny;
    const R *__restrict__ Iref = Iref_all + (size_t)pair * K.n;
    const R *__restrict__ Imov = Imov_all + (size_t)pair * K.n;
    const vec2_t<R> *__restrict__ u = pick(K, B_EST_CUR, h, pair);
    vec2_t<R> *__restrict__ corr = K.c[0] + (size_t)pair * K.n;
    const int tid = threadIdx.x + threadIdx.y * TX;
    const TileWalk T(nx, ny);
    bool divzero = false;
#if OF2D_RELAXED
    const R sratio = sigma_isq / sigma_xsq;
#endif
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        const int i = i0 + threadIdx.x;
        const bool interior = i0 >= 1 && j0 >= 1 && i0 + TILE < nx && j0 + TILE < ny;   // central differences everywhere
        R iref[PY];
        if (interior) {   // in flight while the halo tile is warped
#pragma unroll
            for (int p = 0; p < PY; p++) iref[p] = Iref[i + (j0 + threadIdx.y + p * TY) * nx];
        }
        __syncthreads();
        for (int e = tid; e < (TILE + 2) * (TILE + 2); e += TX * TY) {
            const int r = e / (TILE + 2), cc = e - r * (TILE + 2);
            const int ii = i0 + cc - 1, j = j0 + r - 1;
            R w = (R)0;
            if (ii >= 0 && ii < nx && j >= 0 && j < ny) {
                const int idx = ii + j * nx;
#if OF2D_RELAXED
                w = warp_pixel_lazy<R>(Imov, nx, ny, ii, j, u[idx], idx);
#else
                w = warp_pixel<R>(Imov, nx, ny, ii, j, u[idx], Imov[idx]);
#endif
            }
            sw[r][cc] = w;
        }
        __syncthreads();
        if (interior) {
#pragma unroll
            for (int p = 0; p < PY; p++) {
                const int jl = threadIdx.y + p * TY;
                const int idx = i + (j0 + jl) * nx;
                const int r = jl + 1, cc = threadIdx.x + 1;
                const R ce = sw[r][cc];
                const R gx = (sw[r][cc + 1] - sw[r][cc - 1]) / (R)2.0f;
                const R gy = (sw[r + 1][cc] - sw[r - 1][cc]) / (R)2.0f;
                const R It = ce - iref[p];
#if OF2D_RELAXED
                const R den = gx * gx + gy * gy + (It * It) * sratio;
                if (den == 0) { divzero = true; corr[idx] = mk2<R>((R)0, (R)0); continue; }
                const R s = -It / den;   // one (approximate, fp32) division per pixel
                corr[idx] = mk2<R>(gx * s, gy * s);
#else
                const R q = It * It * sigma_isq;
                const R den = gx * gx + gy * gy + (inv_sigma_xsq != (R)0 ? q * inv_sigma_xsq : q / sigma_xsq);
                if (den == 0) { divzero = true; corr[idx] = mk2<R>((R)0, (R)0); continue; }
                corr[idx] = mk2<R>(gx * It / den * (R)-1, gy * It / den * (R)-1);
#endif
            }
            continue;
        }
        if (i >= nx) continue;
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY, j = j0 + jl;
            if (j >= ny) continue;
            const int idx = i + j * nx;
            const int r = jl + 1, cc = threadIdx.x + 1;
            const R ce = sw[r][cc];
            R gx, gy;   // gradients.h:9-32 on the warped image
            if (i == 0) gx = sw[r][cc + 1] - ce;
            else if (i == nx - 1) gx = ce - sw[r][cc - 1];
            else gx = (sw[r][cc + 1] - sw[r][cc - 1]) / (R)2.0f;
            if (j == 0) gy = sw[r + 1][cc] - ce;
            else if (j == ny - 1) gy = ce - sw[r - 1][cc];
            else gy = (sw[r + 1][cc] - sw[r - 1][cc]) / (R)2.0f;
            const R It = ce - Iref[idx];
            const R den = gx * gx + gy * gy + It * It * sigma_isq / sigma_xsq;
            if (den == 0) { divzero = true; corr[idx] = mk2<R>((R)0, (R)0); continue; }
            corr[idx] = mk2<R>(gx * It / den * (R)-1, gy * It / den * (R)-1);
        }
    }
    if (divzero) atomicOr(&c->flags, OF2D_FLAG_DIVZERO);
}

// ---------------------------------------------------------------------------------------------
// convolution (Field.tpp:210-269) on a shared-memory tile.  The bounds test of the reference is on
// the LINEAR index, so a tile is simply rows of the flattened array: element (row r, column c) of the
// halo is in[r*nx + c] whenever that flat index is inside [0, n) -- columns outside [0, nx) land in
// the neighbouring row exactly as in the reference.  Each thread produces 4 vertically adjacent
// outputs from a sliding column window; per output the taps are accumulated in the reference's order
// (ii outer, jj inner, unfused multiply-add), so the result is bit-identical.
//   EPI 0: none   EPI 1: Logger epilogue (result is the next estimate)   EPI 2: maxabs epilogue
//   KW > 0: odd kernel width known at compile time; KW == 0: any width (slower)
// ---------------------------------------------------------------------------------------------
constexpr int kConvMaxW = 15;
template <class R>
struct ConvW {
    R w[kConvMaxW * kConvMaxW];   // (real) weights, [kw*kh] column-major as Kernel::get_kernel()
    const double *taps_d;         // the same weights in double (renormalisation of truncated windows)
    double full_weight;           // sum over the visiting order
    int kw;
    float neg_zero;               // -0.0f, opaque to the compiler (see the packed fast path of k_e_conv)
#if OF2D_RELAXED
    // relaxed build: a Gaussian (Kernel::set_gaussian, Kernel.cpp:45-73) is a product w[ii][jj] = sx[ii] sy[jj] up to rounding:
    // full windows take kw + kw taps instead of kw * kw; 1 / full_weight is folded into sy.  separable == 0: dense taps.
    R sx[kConvMaxW], sy[kConvMaxW];
    int separable;
#endif
};

template <class R, int EPI, int KW>
__global__ void __launch_bounds__(TX *TY, sizeof(R) == 4 ? 5 : 1) k_e_conv(EngK<R> K, int src_buf, int dst_buf, const __grid_constant__ ConvW<R> W, int nsq_cap) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny;
    const long n = (long)K.n;
    const vec2_t<R> *__restrict__ in = pick(K, src_buf, h, pair);
    vec2_t<R> *__restrict__ out = pick(K, dst_buf, h, pair);
    const vec2_t<R> *__restrict__ est_cur = pick(K, B_EST_CUR, h, pair);
    const int kw = KW > 0 ? KW : W.kw;
    const int cx = (kw - 1) / 2;
    const int SW = TILE + 2 * cx, SH = TILE + 2 * cx;
    // Two tile stages.  Interior tiles arrive by TMA: one bulk copy per tile row (a tile row is a contiguous range of
    // the FLAT array, which is exactly the reference's addressing), issued by warp 0 one tile ahead of the compute.
    // 16-byte alignment of the copies: for 8-byte elements the window is extended by `shift` elements to the left.
    const int shift = sizeof(vec2_t<R>) == 8 ? (cx & 1) : 0;
    const int SWp = (SW + shift + 1) & ~1;
    const bool rows_aligned = sizeof(vec2_t<R>) == 16 || (nx & 1) == 0;
    vec2_t<R> *const stages = reinterpret_cast<vec2_t<R> *>(smem_raw);   // [2][SH][SWp]; indexed arithmetically so accesses stay LDS/STS
    const int stage_elems = SH * SWp;
    __shared__ uint64_t cbar[2];
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) { mbar_init(&cbar[0], 1); mbar_init(&cbar[1], 1); mbar_init_fence(); }
    __syncthreads();
    const TileWalk T(nx, ny);
    auto tile_start = [&](int tile, int r) -> long {   // flat index of element (r, -shift) of the tile's window
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        return (long)(j0 + r - cx) * nx + (i0 - cx - shift);
    };
    // tile rows 1 .. tiles_y - 3 have their whole window (and its extension to SWp) inside [0, n) for every kernel width
    // up to kConvMaxW: only the first and the last two tile rows pay for the exact 64-bit test
    const int tiles_y = T.ntiles / T.tiles_x;
    auto mid_row = [&](int tile) -> bool { const int ty = T.ty(tile); return ty >= 1 && ty + 2 < tiles_y; };
    auto tma_ok = [&](int tile) -> bool { return rows_aligned && (mid_row(tile) || (tile_start(tile, 0) >= 0 && tile_start(tile, SH - 1) + SWp <= n)); };
    auto issue = [&](int tile, int st) {   // warp 0
        const int lane = threadIdx.x;
        if (lane == 0) { proxy_fence_async(); mbar_expect_tx(&cbar[st], (unsigned)(SH * SWp * sizeof(vec2_t<R>))); }
        __syncwarp();
        for (int r = lane; r < SH; r += 32) bulk_g2s(stages + st * stage_elems + r * SWp, in + tile_start(tile, r), (unsigned)(SWp * sizeof(vec2_t<R>)), &cbar[st]);
    };
    NormAcc<R> acc;
    R mx = (R)0;
    unsigned uses[2] = {0u, 0u};
    if ((int)blockIdx.x < T.ntiles && threadIdx.y == 0 && tma_ok(blockIdx.x)) issue(blockIdx.x, 0);
    int kiter = 0;
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x, kiter++) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        const int st = kiter & 1, next = tile + gridDim.x;
        if (next < T.ntiles && threadIdx.y == 0 && tma_ok(next)) issue(next, st ^ 1);   // stage st^1 was last read one iteration ago
        vec2_t<R> *tile_s = stages + st * stage_elems + shift;    // element (r, cc) of the window at tile_s[r * SWp + cc]
        if (tma_ok(tile)) {
            mbar_wait(&cbar[st], uses[st] & 1u);
            uses[st]++;
        } else {
            for (int e = tid; e < SH * SW; e += TX * TY) {
                const int r = e / SW, cc = e - r * SW;
                const long lin = (long)(j0 + r - cx) * nx + (i0 + cc - cx);
                vec2_t<R> v = mk2<R>((R)0, (R)0);
                if (lin >= 0 && lin < n) v = in[lin];
                tile_s[r * SWp + cc] = v;
            }
            __syncthreads();
        }
        const int i = i0 + threadIdx.x;
        const int jl0 = 4 * threadIdx.y;
        // every tap of every pixel of the tile inside [0, n)?  (first / last flat index of the tile's windows)
        bool tile_interior = mid_row(tile);
        if (!tile_interior) {
            const long lo = (long)(j0 - cx) * nx + (i0 - cx), hi = (long)(min(j0 + TILE, ny) - 1 + cx) * nx + (min(i0 + TILE, nx) - 1 + cx);
            tile_interior = lo >= 0 && hi < n;
        }
        auto epilogue = [&](vec2_t<R> o, vec2_t<R> prev) {
            if (EPI == 1) {
                acc.add(o, prev);
            } else if (EPI == 2) {
                const R s = maxabs_term<R>(o);
                mx = mx < s ? s : mx;
            }
        };
        if (KW > 0 && tile_interior && i0 + TILE <= nx && j0 + TILE <= ny) {
            const long idx0 = i + (long)(j0 + jl0) * nx;
            vec2_t<R> prev[4];
            if (EPI == 1) {   // Logger's prev: in flight while the taps are accumulated
#pragma unroll
                for (int q = 0; q < 4; q++) prev[q] = est_cur[idx0 + (long)q * nx];
            }
            R ax[4], ay[4];
#if OF2D_RELAXED
            if (W.separable) {
                // row pass over the KW + 3 window rows of the thread's 4 outputs, then the column pass: KW + KW taps per output
                // instead of KW * KW (the normalisation is folded into sy)
                constexpr int NRW = 4 + (KW > 0 ? KW : 1) - 1;
                if constexpr (sizeof(R) == 4) {
                    unsigned long long hrow[NRW];
#pragma unroll
                    for (int r = 0; r < NRW; r++) {
                        const float2 e0 = tile_s[(jl0 + r) * SWp + threadIdx.x];
                        hrow[r] = fma_f32x2(pack_f32x2(e0.x, e0.y), pack_f32x2(W.sx[0], W.sx[0]), pack_f32x2(0.0f, 0.0f));
#pragma unroll
                        for (int ii = 1; ii < KW; ii++) {
                            const float2 e = tile_s[(jl0 + r) * SWp + threadIdx.x + ii];
                            hrow[r] = fma_f32x2(pack_f32x2(e.x, e.y), pack_f32x2(W.sx[ii], W.sx[ii]), hrow[r]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        unsigned long long a2 = fma_f32x2(hrow[q], pack_f32x2(W.sy[0], W.sy[0]), pack_f32x2(0.0f, 0.0f));
#pragma unroll
                        for (int jj = 1; jj < KW; jj++) a2 = fma_f32x2(hrow[q + jj], pack_f32x2(W.sy[jj], W.sy[jj]), a2);
                        const float2 e = unpack_f32x2(a2);
                        ax[q] = e.x; ay[q] = e.y;
                    }
                } else {
                    vec2_t<R> hrow[NRW];
#pragma unroll
                    for (int r = 0; r < NRW; r++) {
                        R hx = (R)0, hy = (R)0;
#pragma unroll
                        for (int ii = 0; ii < KW; ii++) {
                            const vec2_t<R> e = tile_s[(jl0 + r) * SWp + threadIdx.x + ii];
                            hx += e.x * W.sx[ii]; hy += e.y * W.sx[ii];
                        }
                        hrow[r] = mk2<R>(hx, hy);
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        R sxq = (R)0, syq = (R)0;
#pragma unroll
                        for (int jj = 0; jj < KW; jj++) { sxq += hrow[q + jj].x * W.sy[jj]; syq += hrow[q + jj].y * W.sy[jj]; }
                        ax[q] = sxq; ay[q] = syq;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const vec2_t<R> o = mk2<R>(ax[q], ay[q]);
                    out[idx0 + (long)q * nx] = o;
                    epilogue(o, prev[q]);
                }
                if (EPI == 1) acc.flush();
                __syncthreads();
                continue;
            }
#endif
            if constexpr (sizeof(R) == 4) {
                // both components of a tap in one packed instruction: FFMA2 with the -0 addend rounds exactly like the
                // reference's multiply (x*t + -0 == x*t, signed zeros included) and FADD2 is its add, so the result is
                // still bit-identical while the issue slots per tap halve.  The addend comes from a kernel parameter:
                // with a literal, ptxas folds the pair into one contracted FFMA2 even under -fmad=false.
                const unsigned long long nz2 = pack_f32x2(W.neg_zero, W.neg_zero);
                unsigned long long a2[4];
#pragma unroll
                for (int q = 0; q < 4; q++) a2[q] = pack_f32x2(0.0f, 0.0f);
#pragma unroll
                for (int ii = 0; ii < KW; ii++) {
                    unsigned long long col[4 + (KW > 0 ? KW : 1) - 1];
#pragma unroll
                    for (int r = 0; r < 4 + KW - 1; r++) { const float2 e = tile_s[(jl0 + r) * SWp + threadIdx.x + ii]; col[r] = pack_f32x2(e.x, e.y); }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
#pragma unroll
                        for (int jj = 0; jj < KW; jj++) {
                            const float t = W.w[ii + jj * KW];
                            a2[q] = add_f32x2(a2[q], fma_f32x2(col[q + jj], pack_f32x2(t, t), nz2));
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) { const float2 e = unpack_f32x2(a2[q]); ax[q] = e.x; ay[q] = e.y; }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) { ax[q] = (R)0; ay[q] = (R)0; }
#pragma unroll
                for (int ii = 0; ii < KW; ii++) {
                    vec2_t<R> col[4 + (KW > 0 ? KW : 1) - 1];
#pragma unroll
                    for (int r = 0; r < 4 + KW - 1; r++) col[r] = tile_s[(jl0 + r) * SWp + threadIdx.x + ii];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
#pragma unroll
                        for (int jj = 0; jj < KW; jj++) {
                            const R t = W.w[ii + jj * KW];
                            ax[q] = ax[q] + col[q + jj].x * t;
                            ay[q] = ay[q] + col[q + jj].y * t;
                        }
                    }
                }
            }
            const R wgt = (R)W.full_weight;
            if (W.full_weight == 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) { const vec2_t<R> ce = tile_s[(jl0 + q + cx) * SWp + threadIdx.x + cx]; ax[q] = ce.x; ay[q] = ce.y; }
            } else if (wgt != (R)1) {   // x / 1 == x: a normalised kernel (Kernel.cpp:66-68) needs no division
#pragma unroll
                for (int q = 0; q < 4; q++) { ax[q] = ax[q] / wgt; ay[q] = ay[q] / wgt; }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const vec2_t<R> o = mk2<R>(ax[q], ay[q]);
                out[idx0 + (long)q * nx] = o;
                epilogue(o, prev[q]);
            }
        } else {
#pragma unroll 1
            for (int q = 0; q < 4; q++) {
                const int j = j0 + jl0 + q;
                const long idx = i + (long)j * nx;
                R ax = (R)0, ay = (R)0;
                double weight = 0.0;
                for (int ii = -cx; ii <= cx; ii++) {
                    for (int jj = -cx; jj <= cx; jj++) {
                        const long lin = idx + ii + (long)jj * nx;
                        if (lin < 0 || lin >= n) continue;
                        const int ik = (ii + cx) + (jj + cx) * kw;
                        weight += W.taps_d[ik];
                        const vec2_t<R> f = tile_s[(jl0 + q + jj + cx) * SWp + (threadIdx.x + ii + cx)];
                        const R t = W.w[ik];
                        ax = ax + f.x * t;
                        ay = ay + f.y * t;
                    }
                }
                vec2_t<R> o;
                if (weight != 0) { const R wg = (R)weight; o = mk2<R>(ax / wg, ay / wg); }
                else o = tile_s[(jl0 + q + cx) * SWp + threadIdx.x + cx];
                if (i < nx && j < ny) {
                    out[idx] = o;
                    epilogue(o, EPI == 1 ? est_cur[idx] : o);
                }
            }
        }
        if (EPI == 1) acc.flush();
        __syncthreads();   // everybody is done with stage st before warp 0 refills it (one iteration from now)
    }
    if (EPI == 0) return;
    if (EPI == 1) {
        logger_epilogue<R>(K, c, pair, acc.dsd, acc.dsp);
    } else {
        mx = block_extreme<R, true>(mx);
        const double vals[1] = {(double)mx};
        double *part = K.partials + (size_t)pair * K.pstride;
        if (publish_partials<1>(vals, part, &c->ticket[1], gridDim.x, blockIdx.x)) {
            double o1[1];
            reduce_partials<1>(part, gridDim.x, o1, 1u, 0u);
            if (tid == 0) {   // Motion::exp, Motion.cpp:253-260
                const R ma = sizeof(R) == 4 ? (R)sqrtf((float)o1[0]) : (R)sqrt(o1[0]);
                int nsq = 0;
                if (ma != 0) {
                    nsq = sizeof(R) == 4 ? (int)ceilf(1 + log2f((float)ma)) : (int)ceil(1 + log2((double)ma));
                    if (nsq < 0) nsq = 0;
                }
                if (nsq > nsq_cap) { c->overflow = 1; nsq = nsq_cap; }
                c->nsquares = nsq;
                c->maxabs = (double)ma;
                c->scale = (double)(R)pow(2.0, (double)-nsq);
                const int it = c->iter;
                if (it < K.tr.cap) { K.tr.nsq[(size_t)pair * K.tr.cap + it] = nsq; K.tr.maxabs[(size_t)pair * K.tr.cap + it] = (double)ma; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// derivatives (IterativeSolver.cpp:22-56), normal and transposed output
// ---------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(TX *TY) k_e_derivatives(EngK<R> K, int gate, const R *__restrict__ Iref_all, const R *__restrict__ Imov_all,
                                                          vec2_t<R> *__restrict__ gradI_all, R *__restrict__ It_all, int transposed) {
    pdl_enter();
    __shared__ vec2_t<R> sg[TILE][TILE + 1];
    __shared__ R st[TILE][TILE + 1];
    const int pair = blockIdx.y;
    const CtlHot h = load_ctl(K.ctl + pair);
    if (!gate_open(h, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const R *__restrict__ Iref = Iref_all + (size_t)pair * K.n;
    const R *__restrict__ Imov = Imov_all + (size_t)pair * K.n;
    const TileWalk T(nx, ny);
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        const int i = i0 + threadIdx.x;
        if (transposed) __syncthreads();
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY, j = j0 + jl;
            if (i < nx && j < ny) {
                const int idx = i + j * nx;
                const vec2_t<R> g = mk2<R>(partial_x<R>(Imov, idx, i, nx), partial_y<R>(Imov, idx, j, nx, ny));
                const R t = Imov[idx] - Iref[idx];
                if (!transposed) {
                    gradI_all[(size_t)pair * K.n + idx] = g;
                    It_all[(size_t)pair * K.n + idx] = t;
                } else {
                    sg[jl][threadIdx.x] = g;
                    st[jl][threadIdx.x] = t;
                }
            }
        }
        if (!transposed) continue;
        __syncthreads();
        const int jt = j0 + threadIdx.x;   // fast thread index runs along j now
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int il = threadIdx.y + p * TY, it = i0 + il;
            if (it < nx && jt < ny) {
                const size_t o = (size_t)pair * K.nT + (size_t)it * K.P + jt;
                gradI_all[o] = sg[threadIdx.x][il];
                It_all[o] = st[threadIdx.x][il];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fluid in the transposed layout (element (i,j) at i*P + j; threadIdx.x runs along j)
// ---------------------------------------------------------------------------------------------
// (the increment R = v - du/dx v.x - du/dy v.y and the time step, OpticalFlowFluid.cpp:60-95, are produced by the sweep kernel: sor_tile.cuh)

// integrate u += dt R (OpticalFlowFluid.cpp:97-121) + Logger + Jacobian minimum of the new field
// (Image.cpp:189-218, :96-104) + break / regrid decisions (ImageRegistrationFluid.cpp:99-124)
// one tile of k_fl_integrate on the general path: any tile (image border, partial tiles, skipped integration), one-sided differences at the edges
template <class R>
__device__ __forceinline__ void fl_integrate_general_tile(const vec2_t<R> *__restrict__ u, vec2_t<R> *un, const vec2_t<R> *__restrict__ incr, int nx, int ny, int P, int i0, int j0,
                                                          bool skip, bool prev_other, R dt, NormAcc<R> &acc, R &mj) {
    auto unew_at = [&](size_t o) -> vec2_t<R> {
        vec2_t<R> v = u[o];
        if (!skip) { const vec2_t<R> r = incr[o]; v.x += r.x * dt; v.y += r.y * dt; }
        return v;
    };
    const int j = j0 + threadIdx.x;
    const int ib = i0 + threadIdx.y;
    if (j >= ny) return;
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int i = ib + p * TY;
        if (i >= nx) continue;
        const size_t o = (size_t)i * P + j;
        const vec2_t<R> nv = unew_at(o);
        const vec2_t<R> prev = prev_other ? un[o] : u[o];   // after a regrid Logger's prev is the pre-reset estimate
        acc.add(nv, prev);
        // Jacobian of the new field (one-sided at the edges, gradients.h:9-32)
        vec2_t<R> dx, dy;
        if (i == 0) { const vec2_t<R> a = unew_at(o + P); dx = mk2<R>(a.x - nv.x, a.y - nv.y); }
        else if (i == nx - 1) { const vec2_t<R> b = unew_at(o - P); dx = mk2<R>(nv.x - b.x, nv.y - b.y); }
        else { const vec2_t<R> a = unew_at(o + P), b = unew_at(o - P); dx = mk2<R>((a.x - b.x) / (R)2.0f, (a.y - b.y) / (R)2.0f); }
        if (j == 0) { const vec2_t<R> a = unew_at(o + 1); dy = mk2<R>(a.x - nv.x, a.y - nv.y); }
        else if (j == ny - 1) { const vec2_t<R> b = unew_at(o - 1); dy = mk2<R>(nv.x - b.x, nv.y - b.y); }
        else { const vec2_t<R> a = unew_at(o + 1), b = unew_at(o - 1); dy = mk2<R>((a.x - b.x) / (R)2.0f, (a.y - b.y) / (R)2.0f); }
        const R J = ((R)1.0f + dx.x) * ((R)1.0f + dy.y) - dx.y * dy.x;
        mj = J < mj ? J : mj;
        un[o] = nv;
    }
    acc.flush();
}

// Logger + break test + regrid decision of a Fluid iteration (ImageRegistrationFluid.cpp:99-124), taken by the last CTA
template <class R>
__device__ __forceinline__ void fl_integrate_epilogue(const EngK<R> &K, PairCtl *c, int pair, const NormAcc<R> &acc, R mj) {
    double sd = acc.dsd, sp = acc.dsp;
    block_sum2(sd, sp);
    mj = block_extreme<R, false>(mj);
    const double vals[3] = {sd, sp, (double)mj};
    double *part = K.partials + (size_t)pair * K.pstride;
    if (publish_partials<3>(vals, part, &c->ticket[0], gridDim.x, blockIdx.x)) {
        double o3[3];
        reduce_partials<3>(part, gridDim.x, o3, 0u, 4u);
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            const int it = c->iter;
            c->sel ^= 1;
            c->prev_other = 0;
            finalize_logger<R>(c, K.tr, pair, o3[0], o3[1], (unsigned)K.n, K.n_active);
            const bool brk = (R)c->err < (R)0.001f && it > 1;
            const R minjac = (R)o3[2];
            int rg = 0;
            if (!brk && minjac < (R)0.5) rg = 1;
            c->regrid = rg;
            c->minjac = (double)minjac;
            if (it < K.tr.cap) { K.tr.regrid[(size_t)pair * K.tr.cap + it] = rg; K.tr.minjac[(size_t)pair * K.tr.cap + it] = (double)minjac; }
        }
    }
}

template <class R>
__global__ void __launch_bounds__(TX *TY, OF2D_INTEG_MINB) k_fl_integrate(EngK<R> K, const vec2_t<R> *__restrict__ incr_all) {
    pdl_enter();
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny, P = K.P;
    const vec2_t<R> *__restrict__ u = pick(K, B_EST_CUR, h, pair, true);
    vec2_t<R> *un = pick(K, B_EST_NEXT, h, pair, true);
    const vec2_t<R> *__restrict__ incr = incr_all + (size_t)pair * K.nT;
    const bool skip = h.skip != 0;
    const bool prev_other = h.prev_other != 0;
    const R dt = (R)__ldcg(&c->dt);
    __shared__ vec2_t<R> s_new[(TILE + 2) * (TILE + 2)], s_old[(TILE + 2) * (TILE + 2)];   // halo tile of an interior tile: new and old field
    const TileWalk T(ny, nx);
    NormAcc<R> acc;
    R mj = (R)INFINITY;
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int j0 = T.tx(tile) * TILE, i0 = T.ty(tile) * TILE;
        if (!skip && j0 >= 1 && j0 + TILE < ny && i0 >= 1 && i0 + TILE < nx) {
            // interior tile, two phases.  (1) The CTA evaluates the NEW field once per point of the 34 x 34 halo tile
            // (coalesced loads of u and of the increment, all of a thread's loads issued before the arithmetic) into
            // shared memory, next to the old values.  (2) A thread owns 4 consecutive i of one j and reads its line
            // i-1 .. i+4 and its j-1 / j+1 neighbours from shared memory; all differences are central.  Same
            // expressions as the general path below.
            constexpr int HT = TILE + 2, NE = HT * HT, NR = (NE + TX * TY - 1) / (TX * TY);
            const int tid = threadIdx.x + threadIdx.y * TX;
            __syncthreads();   // the previous tile's phase 2 is over
            {
                vec2_t<R> uo[NR], rr[NR];
#pragma unroll
                for (int k = 0; k < NR; k++) {
                    const int e = min(tid + k * TX * TY, NE - 1);
                    const int r = e / HT, cc = e - r * HT;
                    const size_t o = (size_t)(i0 - 1 + r) * P + (size_t)(j0 - 1 + cc);
                    uo[k] = u[o]; rr[k] = incr[o];
                }
#pragma unroll
                for (int k = 0; k < NR; k++) {
                    const int e = tid + k * TX * TY;
                    if (e < NE) {
                        s_old[e] = uo[k];
                        s_new[e] = mk2<R>(uo[k].x + rr[k].x * dt, uo[k].y + rr[k].y * dt);
                    }
                }
            }
            __syncthreads();
            const size_t o0 = (size_t)(i0 + 4 * threadIdx.y) * P + (size_t)(j0 + threadIdx.x);
            const int sb = (4 * threadIdx.y) * HT + threadIdx.x + 1;   // (line i-1 of the thread, its own j) in the halo tile
            vec2_t<R> ce[6], le[4], ri[4], pv[4];
#pragma unroll
            for (int r = 0; r < 6; r++) ce[r] = s_new[sb + r * HT];
#pragma unroll
            for (int q = 0; q < 4; q++) { le[q] = s_new[sb + (q + 1) * HT - 1]; ri[q] = s_new[sb + (q + 1) * HT + 1]; pv[q] = s_old[sb + (q + 1) * HT]; }
            if (prev_other) {
#pragma unroll
                for (int q = 0; q < 4; q++) pv[q] = un[o0 + (size_t)q * P];
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const vec2_t<R> nv = ce[q + 1];
                acc.add(nv, pv[q]);
                const vec2_t<R> dx = mk2<R>((ce[q + 2].x - ce[q].x) / (R)2.0f, (ce[q + 2].y - ce[q].y) / (R)2.0f);
                const vec2_t<R> dy = mk2<R>((ri[q].x - le[q].x) / (R)2.0f, (ri[q].y - le[q].y) / (R)2.0f);
                const R J = ((R)1.0f + dx.x) * ((R)1.0f + dy.y) - dx.y * dy.x;
                mj = J < mj ? J : mj;
                un[o0 + (size_t)q * P] = nv;
            }
            acc.flush();
            continue;
        }
        fl_integrate_general_tile<R>(u, un, incr, nx, ny, P, i0, j0, skip, prev_other, dt, acc, mj);
    }
    fl_integrate_epilogue<R>(K, c, pair, acc, mj);
}

// Fluid regrid, second half (ImageRegistrationFluid.cpp:116-124): Iaux = Imov o (id + level motion), derivatives of
// (Iref, Iaux) in the transposed layout, estimate <- 0 -- one kernel.  The warped image of a 32 x 32 tile (+1 halo) is
// evaluated once into shared memory (Image::warp2d, Image.cpp:119-182) and never reaches HBM; the derivatives
// (IterativeSolver.cpp:22-56) are turned through shared memory so that the transposed stores are coalesced.
template <class R>
__global__ void __launch_bounds__(TX *TY) k_fl_rewarp(EngK<R> K, int gate, const R *__restrict__ Iref_all, const R *__restrict__ Imov_all, int u_buf,
                                                      vec2_t<R> *__restrict__ gradI_all, R *__restrict__ It_all, int zero_buf) {
    pdl_enter();
    __shared__ R sw[TILE + 2][TILE + 2 + 1];
    __shared__ vec2_t<R> sg[TILE][TILE + 1];
    __shared__ R st[TILE][TILE + 1];
    const int pair = blockIdx.y;
    const CtlHot h = load_ctl(K.ctl + pair);
    if (!gate_open(h, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const R *__restrict__ Iref = Iref_all + (size_t)pair * K.n;
    const R *__restrict__ Imov = Imov_all + (size_t)pair * K.n;
    const vec2_t<R> *__restrict__ u = pick(K, u_buf, h, pair);
    vec2_t<R> *__restrict__ zero = pick(K, zero_buf, h, pair, true);
    const int tid = threadIdx.x + threadIdx.y * TX;
    const TileWalk T(nx, ny);
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        __syncthreads();
        for (int e = tid; e < (TILE + 2) * (TILE + 2); e += TX * TY) {
            const int r = e / (TILE + 2), cc = e - r * (TILE + 2);
            const int i = i0 + cc - 1, j = j0 + r - 1;
            R w = (R)0;
            if (i >= 0 && i < nx && j >= 0 && j < ny) { const int idx = i + j * nx; w = warp_pixel_lazy<R>(Imov, nx, ny, i, j, u[idx], idx); }
            sw[r][cc] = w;
        }
        __syncthreads();
        const int i = i0 + threadIdx.x;
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY, j = j0 + jl;
            if (i < nx && j < ny) {
                const int r = jl + 1, cc = threadIdx.x + 1;
                const R ce = sw[r][cc];
                R gx, gy;   // gradients.h:9-32 on the warped image
                if (i == 0) gx = sw[r][cc + 1] - ce;
                else if (i == nx - 1) gx = ce - sw[r][cc - 1];
                else gx = (sw[r][cc + 1] - sw[r][cc - 1]) / (R)2.0f;
                if (j == 0) gy = sw[r + 1][cc] - ce;
                else if (j == ny - 1) gy = ce - sw[r - 1][cc];
                else gy = (sw[r + 1][cc] - sw[r - 1][cc]) / (R)2.0f;
                sg[jl][threadIdx.x] = mk2<R>(gx, gy);
                st[jl][threadIdx.x] = ce - Iref[i + j * nx];
            }
        }
        __syncthreads();
        const int jt = j0 + threadIdx.x;   // fast thread index runs along j now
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int il = threadIdx.y + p * TY, it = i0 + il;
            if (it < nx && jt < ny) {
                const size_t o = (size_t)it * K.P + jt;
                gradI_all[(size_t)pair * K.nT + o] = sg[threadIdx.x][il];
                It_all[(size_t)pair * K.nT + o] = st[threadIdx.x][il];
                zero[o] = mk2<R>((R)0, (R)0);
            }
        }
    }
}

}  // namespace
