// sor_tile.cuh -- the lexicographic SOR sweep of OpticalFlowElastic::SOR_iteration /
// OpticalFlowFluid::SOR_iteration (OpticalFlowElastic.cpp:21-55, OpticalFlowFluid.cpp:7-41) as
// independent overlapped tiles.
//
// The sweep updates in place with x (i) outer and y (j) inner, so cell (i,j) sees NEW values at
// (i-1, j-1..j+1) and (i, j-1).  It is a lower-triangular solve whose off-diagonal weights are
//     a  = |c_relax| mu                 (south, same column)
//     aW = |c_relax| (2 mu + lambda)    (west)         aD = |c_relax| (mu + lambda) / 4   (north-/south-west)
// With the reference's under-relaxation (omega = 0.66) these sum to < 0.4, so the influence of a cell on
// cells further along the sweep decays geometrically: about 0.33 per column and 0.14 per row.  A tile
// that starts the same sequential sweep HW columns to the west / HS rows to the south / HN rows to the
// north of the cells it owns, taking OLD values on that outer ring, reproduces the exact sweep on its
// own cells up to eps = 2^-40 (fp32) / 2^-70 (fp64) of the iteration's step size (sor_plan) -- far below half an ulp,
// i.e. bit-identical in practice (tests/test_engine_gpu.py compares with the exact wavefront kernel).
// sor_plan() derives HW / HS / HN / M from the parameters and refuses (-> exact wavefront path) when
// they do not contract fast enough.
//
// Work layout (fields are in the transposed layout, element (i,j) at i*P + j, so a column is contiguous):
//   CTA = NT threads, thread t owns RPT consecutive rows; the CTA walks its columns west -> east.
//   Per column: TMA bulk copies (cp.async.bulk, one per field) stream the column segments of x_old,
//   gradI, It (and the motion the force is evaluated on) into an NS-stage shared-memory ring, each
//   stage guarded by an mbarrier.  A thread forms, for its rows, everything that does not depend on the
//   cell below (d), solves its RPT-row recurrence x_r = a x_{r-1} + d_r with a zero carry, publishes its
//   top value; after ONE __syncthreads per column it adds the carry from the threads below
//   (sum_q a^(RPT q) x_top[t-q], M terms).
#pragma once

#include <stdlib.h>

#include "device_math.cuh"
#include "engine_ctl.cuh"
#include "tma.cuh"

struct SorPlan {
    int nx, ny, P, batch;
    int NT, RPT, NS;
    int BX, BY, HW, HS, HN, M;
    int nbands, nstrips;
    double c_keep, c_relax, mu, mupl;   // double copies: halo / contraction estimates only
    double lambda, omega;               // the kernel's coefficients are formed in the field precision (sor_tile_launch)
    size_t nT;
    int supported;
};

namespace {

#ifndef OF2D_SOR_NS
#define OF2D_SOR_NS 6
#endif
constexpr int SOR_NS = OF2D_SOR_NS;

template <class R>
struct SorTileArgs {
    int nx, ny, P;
    size_t nT, n;
    int BX, BY, HW, HS, HN, M, LR;
    int which;                     // 0: x = estimate (sel, Logger epilogue); 1: x = velocity (vsel)
    vec2_t<R> *x[2];
    const vec2_t<R> *uf[2];        // fluid: the estimate the force is evaluated on (picked by sel)
    vec2_t<R> *incr;               // fluid: the increment R = v - du/dx v.x - du/dy v.y of the new velocity (OpticalFlowFluid.cpp:60-90), written here
    const vec2_t<R> *gradI;
    const R *It;
    R ck, cr, mu, mupl, a;
    float neg_zero;                // -0.0f, opaque to the compiler (PairOps<float>)
    PairCtl *ctl;
    int *n_active;
    double *partials;
    size_t pstride;
    TraceDev tr;
};

// Component-wise arithmetic on (x, y) pairs in the reference's rounding (a multiplication and an addition are two roundings).
// fp32: the packed sm_100a instructions (SASS FFMA2 / FADD2, both components in one issue slot); the product is
// fma(a, b, -0) with the -0 taken from a kernel argument, which rounds exactly like the multiplication (signed zeros
// included) and keeps ptxas from contracting it with the addition that follows (it does contract mul.f32x2 + add.f32x2
// even under -fmad=false).  fp64: plain component arithmetic.
template <class R> struct PairOps;
#if OF2D_RELAXED
// relaxed build: fused multiply-adds (fp32: packed, both components per instruction)
template <> struct PairOps<double> {
    __device__ __forceinline__ explicit PairOps(float) {}
    __device__ __forceinline__ double2 add(double2 a, double2 b) const { return make_double2(a.x + b.x, a.y + b.y); }
    __device__ __forceinline__ double2 sub(double2 a, double2 b) const { return make_double2(a.x - b.x, a.y - b.y); }
    __device__ __forceinline__ double2 mul(double2 a, double2 b) const { return make_double2(a.x * b.x, a.y * b.y); }
    __device__ __forceinline__ double2 scale(double s, double2 a) const { return make_double2(s * a.x, s * a.y); }
    __device__ __forceinline__ double2 fma(double s, double2 a, double2 c) const { return make_double2(::fma(s, a.x, c.x), ::fma(s, a.y, c.y)); }
};
template <> struct PairOps<float> {
    __device__ __forceinline__ explicit PairOps(float) {}
    __device__ __forceinline__ float2 add(float2 a, float2 b) const { return unpack_f32x2(add_f32x2(pack_f32x2(a.x, a.y), pack_f32x2(b.x, b.y))); }
    __device__ __forceinline__ float2 sub(float2 a, float2 b) const {
        unsigned long long r;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack_f32x2(a.x, a.y)), "l"(pack_f32x2(b.x, b.y)));
        return unpack_f32x2(r);
    }
    __device__ __forceinline__ float2 mul(float2 a, float2 b) const {
        unsigned long long r;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack_f32x2(a.x, a.y)), "l"(pack_f32x2(b.x, b.y)));
        return unpack_f32x2(r);
    }
    __device__ __forceinline__ float2 scale(float s, float2 a) const { return mul(make_float2(s, s), a); }
    __device__ __forceinline__ float2 fma(float s, float2 a, float2 c) const { return unpack_f32x2(fma_f32x2(pack_f32x2(s, s), pack_f32x2(a.x, a.y), pack_f32x2(c.x, c.y))); }
};
#else
template <> struct PairOps<double> {
    __device__ __forceinline__ explicit PairOps(float) {}
    __device__ __forceinline__ double2 add(double2 a, double2 b) const { return make_double2(a.x + b.x, a.y + b.y); }
    __device__ __forceinline__ double2 sub(double2 a, double2 b) const { return make_double2(a.x - b.x, a.y - b.y); }
    __device__ __forceinline__ double2 mul(double2 a, double2 b) const { return make_double2(a.x * b.x, a.y * b.y); }
    __device__ __forceinline__ double2 scale(double s, double2 a) const { return make_double2(s * a.x, s * a.y); }
};
template <> struct PairOps<float> {
    unsigned long long nz2;
    __device__ __forceinline__ explicit PairOps(float neg_zero) : nz2(pack_f32x2(neg_zero, neg_zero)) {}
    __device__ __forceinline__ float2 add(float2 a, float2 b) const { return unpack_f32x2(add_f32x2(pack_f32x2(a.x, a.y), pack_f32x2(b.x, b.y))); }
    __device__ __forceinline__ float2 sub(float2 a, float2 b) const {
        unsigned long long r;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack_f32x2(a.x, a.y)), "l"(pack_f32x2(b.x, b.y)));
        return unpack_f32x2(r);
    }
    __device__ __forceinline__ float2 mul(float2 a, float2 b) const { return unpack_f32x2(fma_f32x2(pack_f32x2(a.x, a.y), pack_f32x2(b.x, b.y), nz2)); }
    __device__ __forceinline__ float2 scale(float s, float2 a) const { return unpack_f32x2(fma_f32x2(pack_f32x2(s, s), pack_f32x2(a.x, a.y), nz2)); }
};
#endif

// the thread's RPT = 4 consecutive ring elements (index r0 = 4 t: 32-byte aligned for {x, y} floats, 16-byte for float scalars).
// fp32: 128-bit shared-memory loads -- the thread stride of 32 bytes makes 64-bit loads 4-way and 32-bit loads 4-way bank
// conflicted, 128-bit ones 2-way (pairs) / conflict-free (scalars); fp64 elements are 128 / 64 bits wide already.
__device__ __forceinline__ void ld_block4(const float2 *p, float2 (&o)[4]) {
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 2);
    o[0] = make_float2(a.x, a.y); o[1] = make_float2(a.z, a.w); o[2] = make_float2(b.x, b.y); o[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void ld_block4(const double2 *p, double2 (&o)[4]) {
#pragma unroll
    for (int r = 0; r < 4; r++) o[r] = p[r];
}
__device__ __forceinline__ void ld_block4(const float *p, float (&o)[4]) {
    const float4 a = *reinterpret_cast<const float4 *>(p);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}
__device__ __forceinline__ void ld_block4(const double *p, double (&o)[4]) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

template <class R, int RPT, bool FLUID, bool WARP>
__global__ void __launch_bounds__(WARP ? 32 : 128) k_sor_tile(SorTileArgs<R> A) {
    pdl_enter();
    using V = vec2_t<R>;
    static_assert(RPT == 4, "row blocks of 4: 32-byte aligned vector loads from the ring");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[SOR_NS], empty[SOR_NS];
    const int pair = blockIdx.z;
    PairCtl *c = A.ctl + pair;
    if (!__ldcg(&c->active)) return;

    // !WARP: the CTA is NT compute threads plus one PRODUCER warp (threads NT .. NT+31) that does nothing but stream
    // the column ring (TMA bulk copies), so the address arithmetic and the copy issue are off the sweep's critical path
    const int NT = WARP ? blockDim.x : blockDim.x - 32, t = threadIdx.x;
    const int nx = A.nx, ny = A.ny, P = A.P, LR = A.LR;
    const int is = 1 + blockIdx.x * A.BX, ie = min(is + A.BX, nx - 1);
    const int js = 1 + blockIdx.y * A.BY, je = min(js + A.BY, ny - 1);
    const int ic0 = max(1, is - A.HW);
    const int jc0 = max(1, js - A.HS), jc1 = min(ny - 1, je + A.HN);
    const int jl0 = (jc0 - 1) & ~3;              // first loaded row (multiple of 4: 16/32-byte aligned bulk copies)
    const int ncols = ie - ic0 + 2;              // loaded columns ic0-1 .. ie

    const int xsel = A.which ? __ldcg(&c->vsel) : __ldcg(&c->sel);
    const V *__restrict__ xin = A.x[xsel] + (size_t)pair * A.nT;
    V *__restrict__ xout = A.x[xsel ^ 1] + (size_t)pair * A.nT;
    const V *__restrict__ ufp = FLUID ? A.uf[__ldcg(&c->sel)] + (size_t)pair * A.nT : nullptr;
    const V *__restrict__ gp = A.gradI + (size_t)pair * A.nT;
    const R *__restrict__ tp = A.It + (size_t)pair * A.nT;
    V *__restrict__ incr = FLUID ? A.incr + (size_t)pair * A.nT : nullptr;

    // stage layout: [x LR + 4 (2 pad elements in front)][gradI LR][uf LR (fluid)][It LR]; row r of the tile at index r
    const size_t x_bytes = (size_t)(LR + 4) * sizeof(V);
    const size_t stage_bytes = x_bytes + (size_t)LR * (sizeof(V) * (FLUID ? 2 : 1) + sizeof(R));
    auto st_x = [&](int s) { return reinterpret_cast<V *>(smem_raw + s * stage_bytes) + 2; };
    auto st_g = [&](int s) { return reinterpret_cast<V *>(smem_raw + s * stage_bytes + x_bytes); };
    auto st_u = [&](int s) { return reinterpret_cast<V *>(smem_raw + s * stage_bytes + x_bytes) + LR; };
    auto st_t = [&](int s) { return reinterpret_cast<R *>(smem_raw + s * stage_bytes + x_bytes + (size_t)LR * sizeof(V) * (FLUID ? 2 : 1)); };
    // exchange between the threads of a column step, double-buffered over k: [2][top NT | d0 NT] (separate arrays: consecutive threads, consecutive words)
    V *pub = reinterpret_cast<V *>(smem_raw + SOR_NS * stage_bytes);
    const unsigned tx_bytes = (unsigned)((size_t)LR * (sizeof(V) * (FLUID ? 3 : 2) + sizeof(R)));

    auto issue = [&](int k) {   // thread 0: loaded column k -> stage k % NS (TMA bulk copies, one per field)
        const int s = k % SOR_NS;
        const size_t g = (size_t)(ic0 - 1 + k) * P + jl0;
        mbar_expect_tx(&full[s], tx_bytes);
        bulk_g2s(st_x(s), xin + g, (unsigned)(LR * sizeof(V)), &full[s]);
        bulk_g2s(st_g(s), gp + g, (unsigned)(LR * sizeof(V)), &full[s]);
        if (FLUID) bulk_g2s(st_u(s), ufp + g, (unsigned)(LR * sizeof(V)), &full[s]);
        bulk_g2s(st_t(s), tp + g, (unsigned)(LR * sizeof(R)), &full[s]);
    };

    if (t == 0) {
        for (int s = 0; s < SOR_NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    float sdf = 0.0f, spf = 0.0f;
    double sdd = 0.0, spd = 0.0;
    R mxr = (R)0;
    if (WARP ? t == 0 : t == NT) {
        const int pre = ncols < SOR_NS ? ncols : SOR_NS;
        for (int k = 0; k < pre; k++) issue(k);
        if (!WARP) {
            // column k goes into the stage column k - NS was in, once the compute threads have released it:
            // the (k / NS - 1)-th completion of that stage's `empty` barrier
            for (int k = SOR_NS; k < ncols; k++) {
                mbar_wait(&empty[k % SOR_NS], (unsigned)((k / SOR_NS) - 1) & 1u);
                proxy_fence_async();
                issue(k);
            }
        }
    }
    if (WARP || t < NT) {
    const int r0 = t * RPT;                      // ring index of the thread's first row
    const int j0 = jl0 + r0;                     // its image row
    bool comp[RPT];
    bool allcomp = true;
#pragma unroll
    for (int r = 0; r < RPT; r++) { comp[r] = (j0 + r >= jc0) && (j0 + r < jc1); allcomp = allcomp && comp[r]; }
    const bool above_comp = (j0 + RPT >= jc0) && (j0 + RPT < jc1);
    const R a = A.a;
    const R aR = (a * a) * (a * a);
    const PairOps<R> po(A.neg_zero);

    // rows j0-1 .. j0+RPT of the previous (new) column, rows j0-1 .. j0+RPT of the current (old) column
    V newW[RPT + 2], oldC[RPT + 2];
    mbar_wait(&full[0], 0);
    mbar_wait(&full[1 % SOR_NS], 0);
    {
        const V *x0 = st_x(0), *x1 = st_x(1 % SOR_NS);
#pragma unroll
        for (int r = 0; r < RPT + 2; r++) { newW[r] = x0[r0 - 1 + r]; oldC[r] = x1[r0 - 1 + r]; }   // (once per tile)
    }
    // fluid: the estimate u on the thread's rows of the column to the west (for du/dx of the fused increment)
    V uW[RPT];
    if (FLUID) {
        const V *u0 = st_u(0);
#pragma unroll
        for (int r = 0; r < RPT; r++) uW[r] = u0[r0 + r];
    }

    // carry = sum_q cq[q] top[t-q]: the new value of the row below the thread's block from the zero-carry tops of the
    // threads below (coefficients a^(RPT (q-1)) while those threads are fully computed blocks; constant over columns)
    constexpr int MQ = 8;
    R cq[MQ];
    int tqi[MQ];
    {
        R coef = (R)1;
        bool alive = t > 0;
#pragma unroll
        for (int q = 1; q <= MQ; q++) {
            const int tq = t - q;
            alive = alive && tq >= 0 && q <= A.M;
            cq[q - 1] = alive ? coef : (R)0;
            tqi[q - 1] = tq >= 0 ? tq : 0;
            const int jq = jl0 + tq * RPT;
            alive = alive && (jq >= jc0 && jq + RPT - 1 < jc1);
            coef *= aR;
        }
    }
    bool own[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) own[r] = (j0 + r >= js) && (j0 + r < je);
#if OF2D_RELAXED
    // the column recurrence is linear in the carry: x_r = xt_r + ap[r] * carry, ap[r] = a^(r+1) over the computed rows of the block
    R ap[RPT];
    {
        R coef = (R)1;
#pragma unroll
        for (int r = 0; r < RPT; r++) { coef = comp[r] ? coef * a : (R)0; ap[r] = coef; }
    }
    const R ncrmu = -(A.cr * A.mu), ncrmupl = -(A.cr * A.mupl);
#endif

    for (int k = 1; k <= ncols - 2; k++) {
        const int i = ic0 - 1 + k;
        const int sC = k % SOR_NS, sE = (k + 1) % SOR_NS;
        mbar_wait(&full[sE], (unsigned)(((k + 1) / SOR_NS) & 1));
        V oldE[RPT + 2];
        {
            const V *xe = st_x(sE);
            oldE[0] = xe[r0 - 1];
            {
                V blk[4];
                ld_block4(xe + r0, blk);
#pragma unroll
                for (int r = 0; r < RPT; r++) oldE[r + 1] = blk[r];
            }
            oldE[RPT + 1] = xe[r0 + RPT];
        }
        // everything of the reference's expression that does not involve the cell below (S):
        //   o = ck C + cr (b - mu (((E + W) + N) + S) - mupl (E + W + 0.25 (NE' - NW' - SE' + SW')))
        V ckC[RPT], bb[RPT], sum3[RPT], k2[RPT], mk2v[RPT], d[RPT], xt[RPT], uC[RPT];   // mk2v = mupl * k2
        V uS = mk2<R>((R)0, (R)0), uN = uS;   // fluid: u on the rows below / above the thread's block, this column (read before the stage is refilled)
        {
            const V *gs = st_g(sC);
            const R *ts = st_t(sC);
            const V *us = st_u(sC);
            if (FLUID && i >= is) { uS = us[r0 - 1]; uN = us[r0 + RPT]; }
            V gblk[4], ublk[4];
            R tblk[4];
            ld_block4(gs + r0, gblk);
            ld_block4(ts + r0, tblk);
            if (FLUID) ld_block4(us + r0, ublk);
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const V Cc = oldC[r + 1], N = oldC[r + 2];
                const V W = newW[r + 1], SW = newW[r], NW = newW[r + 2];
                const V E = oldE[r + 1], SE = oldE[r], NE = oldE[r + 2];
                uC[r] = FLUID ? ublk[r] : Cc;
#if OF2D_RELAXED
                {   // the same expression regrouped: d = ck C + (cr s) dI - (cr mu) sum3 - (cr mupl) k2, fused multiply-adds
                    const V dI = gblk[r];
                    const R sf = tblk[r] + uC[r].x * dI.x + uC[r].y * dI.y;
                    const V ew = po.add(E, W);
                    sum3[r] = po.add(ew, N);
                    const V cr4 = po.scale((R)0.25f, po.add(po.sub(po.sub(NE, NW), SE), SW));
                    k2[r] = po.add(ew, mk2<R>(cr4.y, cr4.x));
                    const V dd = po.fma(A.ck, Cc, po.fma(ncrmu, sum3[r], po.fma(ncrmupl, k2[r], po.scale(A.cr * sf, dI))));
                    d[r] = comp[r] ? dd : Cc;
                    continue;
                }
#endif
                {   // OpticalFlow::get_force (OpticalFlow.cpp:33): s = It + u.x dI.x + u.y dI.y; f = dI s
                    const V dI = gblk[r];
                    const V pr = po.mul(uC[r], dI);
                    const R sf = tblk[r] + pr.x + pr.y;
                    bb[r] = po.scale(sf, dI);
                }
                const V ew = po.add(E, W);
                sum3[r] = po.add(ew, N);
                const V cr4 = po.scale((R)0.25f, po.add(po.sub(po.sub(NE, NW), SE), SW));   // ((NE - NW) - SE) + SW per component
                k2[r] = po.add(ew, mk2<R>(cr4.y, cr4.x));                                    // the x equation takes the y differences and vice versa
                mk2v[r] = po.scale(A.mupl, k2[r]);
                ckC[r] = po.scale(A.ck, Cc);
                const V dd = po.add(ckC[r], po.scale(A.cr, po.sub(po.sub(bb[r], po.scale(A.mu, sum3[r])), mk2v[r])));
                d[r] = comp[r] ? dd : Cc;
            }
        }
        // zero-carry estimate of the column recurrence x_r = a x_{r-1} + d_r over the thread's rows
        xt[0] = d[0];
#pragma unroll
        for (int r = 1; r < RPT; r++) {
            const R ar = comp[r] ? a : (R)0;
#if OF2D_RELAXED
            xt[r] = po.fma(ar, xt[r - 1], d[r]);
#else
            xt[r] = po.add(po.scale(ar, xt[r - 1]), d[r]);
#endif
        }
        V carry = mk2<R>((R)0, (R)0);   // the new value of the row below the thread's block
        V d0_above = mk2<R>((R)0, (R)0);
        if (WARP) {   // one warp per tile: exchange through shuffles, no shared memory round trip and no block barrier
            __syncwarp();
            if (t == 0) {   // every lane has read stage sC: stream the next column(s) in
                proxy_fence_async();
                if (k == 1 && SOR_NS < ncols) issue(SOR_NS);
                if (k + SOR_NS < ncols) issue(k + SOR_NS);
            }
            const V top = xt[RPT - 1];
#pragma unroll
            for (int q = 0; q < MQ; q++) {
                if (q < A.M) {
                    const R tx_ = __shfl_up_sync(0xffffffffu, top.x, q + 1), ty_ = __shfl_up_sync(0xffffffffu, top.y, q + 1);
                    carry = po.add(carry, po.scale(cq[q], mk2<R>(tx_, ty_)));
                }
            }
            d0_above.x = __shfl_down_sync(0xffffffffu, d[0].x, 1);
            d0_above.y = __shfl_down_sync(0xffffffffu, d[0].y, 1);
        } else {
            V *pb_top = pub + (k & 1) * 2 * NT, *pb_d0 = pb_top + NT;
            pb_top[t] = xt[RPT - 1];
            pb_d0[t] = d[0];
            asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");   // the compute threads only
            if (t == 0) {   // every compute thread has read stage sC (and stage 0 after the first step): the producer may refill it
                if (k == 1) mbar_arrive(&empty[0]);
                mbar_arrive(&empty[sC]);
            }
#if OF2D_RELAXED
            // the tops are fetched first (independent loads), then summed in order: a load per term inside the chain cost one shared-memory
            // round trip per term on the critical path of every column step (Elastic 2.69 -> 2.37 ms).  cq[q] = 0 beyond M, so the surplus
            // terms add nothing.  (Issuing the NEXT step's ring loads here as well was measured: +-0.)
            if (A.M <= 4) {   // (M <= 3 at the relaxed truncation; uniform branch)
                V tv[4];
#pragma unroll
                for (int q = 0; q < 4; q++) tv[q] = pb_top[tqi[q]];
#pragma unroll
                for (int q = 0; q < 4; q++) carry = po.fma(cq[q], tv[q], carry);
            } else {
                V tv[MQ];
#pragma unroll
                for (int q = 0; q < MQ; q++) tv[q] = pb_top[tqi[q]];
#pragma unroll
                for (int q = 0; q < MQ; q++) carry = po.fma(cq[q], tv[q], carry);
            }
#else
#pragma unroll
            for (int q = 0; q < MQ; q++) {
                const V tv = pb_top[tqi[q]];
                carry = po.add(carry, po.scale(cq[q], tv));
            }
#endif
            if (t + 1 < NT) d0_above = pb_d0[t + 1];
        }
        // the reference's expression, literally, with S = the (estimated) new value of the cell below
        V xn[RPT];
#if OF2D_RELAXED
#pragma unroll
        for (int r = 0; r < RPT; r++) xn[r] = po.fma(ap[r], carry, xt[r]);   // linear carry correction (ap = 0 on rows that are not computed: xt = old value there)
        if (false)
#endif
        {
            V S = carry;
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                // ckC + cr (b - mu (sum3 + S) - mupl k2), OpticalFlowElastic.cpp:41-48 / OpticalFlowFluid.cpp:27-34
                const V lit = po.add(ckC[r], po.scale(A.cr, po.sub(po.sub(bb[r], po.scale(A.mu, po.add(sum3[r], S))), mk2v[r])));
                xn[r] = comp[r] ? lit : oldC[r + 1];
                S = xn[r];
            }
        }
        V ntop;
#if OF2D_RELAXED
        if (t + 1 < NT) ntop = above_comp ? po.fma(a, xn[RPT - 1], d0_above) : d0_above;
#else
        if (t + 1 < NT) ntop = above_comp ? po.add(po.scale(a, xn[RPT - 1]), d0_above) : d0_above;
#endif
        else ntop = oldC[RPT + 1];
        if (i >= is) {
            V uE[RPT];
            if (FLUID) {   // u on the thread's rows of the column to the east (its stage is refilled one step from now)
                const V *ue = st_u(sE);
#pragma unroll
                for (int r = 0; r < RPT; r++) uE[r] = ue[r0 + r];   // (only on owned columns)
            }
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const int j = j0 + r;
                if (own[r]) {
                    xout[(size_t)i * P + j] = xn[r];
                    if (FLUID) {
                        // the increment of the NEW velocity, OpticalFlowFluid.cpp:60-90 (own cells are interior: central differences,
                        // gradients.h:9-32); border cells keep the 0 the buffer was created with (their velocity is never updated)
                        // (x / 2 == x * 0.5 exactly)
                        const V dudx = po.scale((R)0.5f, po.sub(uE[r], uW[r]));
                        const V nn = r + 1 < RPT ? uC[r + 1 < RPT ? r + 1 : r] : uN, ss = r > 0 ? uC[r > 0 ? r - 1 : 0] : uS;
                        const V dudy = po.scale((R)0.5f, po.sub(nn, ss));
                        const V v = xn[r];
                        const V rr = po.sub(po.sub(v, po.scale(v.x, dudx)), po.scale(v.y, dudy));   // (v - du/dx v.x) - du/dy v.y
                        incr[(size_t)i * P + j] = rr;
                        const R sm = maxabs_term<R>(rr);
                        mxr = mxr < sm ? sm : mxr;
                    }
                    if (!FLUID) {
                        const V oc = oldC[r + 1];
                        const V df = po.sub(xn[r], oc), dsq = po.mul(df, df), osq = po.mul(oc, oc);
                        if (sizeof(R) == 4) { sdf += sqrt_approx((float)(dsq.x + dsq.y)); spf += sqrt_approx((float)(osq.x + osq.y)); }   // Logger addends: as NormAcc (engine_kernels.cuh)
                        else { sdd += sqrt((double)(dsq.x + dsq.y)); spd += sqrt((double)(osq.x + osq.y)); }
                    }
                }
            }
            if (sizeof(R) == 4 && !FLUID && (k & 15) == 0) { sdd += (double)sdf; spd += (double)spf; sdf = 0.0f; spf = 0.0f; }
        }
        if (FLUID) {
#pragma unroll
            for (int r = 0; r < RPT; r++) uW[r] = uC[r];
        }
        newW[0] = carry;
#pragma unroll
        for (int r = 0; r < RPT; r++) newW[r + 1] = xn[r];
        newW[RPT + 1] = ntop;
#pragma unroll
        for (int r = 0; r < RPT + 2; r++) oldC[r] = oldE[r];
    }

    }   // compute threads

    if (FLUID) {
        // time step from the maximum of the increment (OpticalFlowFluid.cpp:92-95, :135-137; Motion.cpp:51-58), taken by the last CTA
        mxr = block_extreme<R, true>(mxr);
        const double vals[1] = {(double)mxr};
        const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.x + blockIdx.y * gridDim.x;
        double *part = A.partials + (size_t)pair * A.pstride;
        if (publish_partials<1>(vals, part, &c->ticket[1], nblocks, bid)) {
            double o1[1];
            reduce_partials<1>(part, nblocks, o1, 1u, 0u);
            if (t == 0) {
                const R maxabs = sizeof(R) == 4 ? (R)sqrtf((float)o1[0]) : (R)sqrt(o1[0]);   // Motion.cpp:57
                const R dt = (R)0.65f / maxabs;                                              // OpticalFlowFluid.h:32, .cpp:93
                c->maxabs = (double)maxabs;
                c->dt = (double)dt;
                c->skip = dt >= (R)65.0f;                                                    // .cpp:135-137
                c->vsel ^= 1;
                const int it = c->iter;
                if (it < A.tr.cap) { A.tr.maxabs[(size_t)pair * A.tr.cap + it] = (double)maxabs; A.tr.dt[(size_t)pair * A.tr.cap + it] = (double)dt; }
            }
        }
        return;
    }
    double sd = sdd + (double)sdf, sp = spd + (double)spf;
    block_sum2(sd, sp);
    const double vals[2] = {sd, sp};
    const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.x + blockIdx.y * gridDim.x;
    double *part = A.partials + (size_t)pair * A.pstride;
    if (publish_partials<2>(vals, part, &c->ticket[0], nblocks, bid)) {
        double out[2];
        reduce_partials<2>(part, nblocks, out, 0u, 0u);
        if (t == 0) {
            c->sel ^= 1;
            finalize_logger<R>(c, A.tr, pair, out[0], out[1], (unsigned)A.n, A.n_active);
        }
    }
}

}  // namespace

// halo widths from the contraction of the sweep; supported = 0 when the parameters do not contract
static inline SorPlan sor_plan(int nx, int ny, int batch, double mu, double lambda, double omega, bool dbl, bool fluid, int sm_count = 148) {
    SorPlan S;
    memset(&S, 0, sizeof(S));
    S.nx = nx; S.ny = ny; S.batch = batch;
    S.P = (ny + 3) & ~3;
    S.NS = SOR_NS;
    const double den = -6.0 * mu - 2.0 * lambda;
    S.c_keep = 1.0 - omega; S.c_relax = den != 0 ? omega / den : 0.0; S.mu = mu; S.mupl = mu + lambda; S.lambda = lambda; S.omega = omega;
    if (den == 0 || nx < 3 || ny < 3) { S.supported = 0; return S; }
    const double cr = fabs(S.c_relax);
    const double a = cr * fabs(mu), aW = cr * (fabs(mu) + fabs(mu + lambda)), aD = 0.25 * cr * fabs(mu + lambda);
    // eps: halo truncation relative to the step of the iteration: 2^-40 (fp32) / 2^-70 (fp64) keep the halo error ~2^-16 ulp below the field values
    double eps = dbl ? ldexp(1.0, -70) : ldexp(1.0, -40);
#if OF2D_RELAXED
    // relaxed build: the truncation sits at the rounding level of the field instead of 2^-16 ulp below it (Elastic at 2048^2, 50
    // sweeps: 7e-9 px from the 2^-40 result); Fluid amplifies perturbations, so its velocity sweep keeps 4 more bits
    eps = dbl ? ldexp(1.0, -48) : (fluid ? ldexp(1.0, -30) : ldexp(1.0, -26));
#endif
    { const char *e = getenv("OF2D_SOR_EPS_LOG2"); if (e && atoi(e) <= -20 && atoi(e) >= -100) eps = ldexp(1.0, atoi(e)); }   // halo truncation (tuning / experiments)
    S.supported = 1;
    if (!(a < 0.6) || !(aW + 2 * aD < 0.6) || !(1.0 - a - aW - aD > 0.15)) { S.supported = 0; return S; }
    const double rx = (aW + 2 * aD) / (1.0 - a), rs = a / (1.0 - aW - 2 * aD), rn = aD / (1.0 - a - aW - aD);
    auto halo = [&](double rho) { return rho <= 1e-12 ? 1 : (int)ceil(log(eps) / log(rho)) + 1; };
    S.HW = halo(rx); S.HS = halo(rs); S.HN = halo(rn);
    S.RPT = 4;
    S.NT = 64;
    // + the producer warp: at most 128 threads (the kernel's launch bounds)
    { const char *e = getenv("OF2D_SOR_NT"); if (e && atoi(e) >= 32 && atoi(e) <= 96) S.NT = atoi(e) & ~31; }
    S.M = a <= 1e-12 ? 1 : (int)ceil(log(eps) / (S.RPT * log(a))) + 1;
    if (S.HW > 96 || S.HS + S.HN > S.NT * S.RPT / 2 || S.M > 8) { S.supported = 0; return S; }
    const int by_max = S.NT * S.RPT - S.HS - S.HN - 8;
    S.nstrips = ceil_div(ny - 2, by_max);
    S.BY = ceil_div(ny - 2, S.nstrips);
    // column bands: as many as keep the GPU about `waves` CTAs per SM deep, but not narrower than the west halo
    // one wave: as many CTAs per SM as the kernel's shared memory lets be resident (at most 4), so the grid is never a
    // full wave plus a remainder (fp64 stages are twice as large: 2-3 CTAs per SM)
    int per_sm = 4;
    {
        const size_t ev = dbl ? 16 : 8, es = dbl ? 8 : 4, LR = (size_t)S.NT * S.RPT + 8;
        const size_t stage = (LR + 4) * ev + LR * (ev * (fluid ? 2 : 1) + es);
        const size_t smem = (size_t)S.NS * stage + 2 * (size_t)S.NT * 2 * ev + 1024;   // + per-CTA reservation
        const int fit = (int)((227u * 1024u) / smem);
        if (fit < per_sm) per_sm = fit < 1 ? 1 : fit;
    }
    { const char *e = getenv("OF2D_SOR_PER_SM"); if (e && atoi(e) > 0) per_sm = atoi(e); }
    long want = (long)sm_count * per_sm / ((long)batch * S.nstrips);
    if (want < 1) want = 1;
    S.BX = ceil_div(nx - 2, (int)want);
    if (S.BX < S.HW) S.BX = S.HW;
    { const char *e = getenv("OF2D_SOR_BX"); if (e && atoi(e) > 0) S.BX = atoi(e); }
    S.nbands = ceil_div(nx - 2, S.BX);
    S.BX = ceil_div(nx - 2, S.nbands);
    S.nT = (size_t)nx * S.P + 1024;
    return S;
}

template <class R>
static int sor_tile_launch(of2d_ctx *ctx, const SorPlan &S, PairCtl *ctl, int *n_active, double *partials, size_t pstride, const TraceDev &tr, int which,
                           vec2_t<R> *x0, vec2_t<R> *x1, const vec2_t<R> *uf0, const vec2_t<R> *uf1, const vec2_t<R> *gradI, const R *It, vec2_t<R> *incr = nullptr) {
    SorTileArgs<R> A;
    A.nx = S.nx; A.ny = S.ny; A.P = S.P; A.nT = S.nT; A.n = (size_t)S.nx * S.ny;
    A.BX = S.BX; A.BY = S.BY; A.HW = S.HW; A.HS = S.HS; A.HN = S.HN; A.M = S.M;
    A.LR = S.NT * S.RPT + 8;
    A.which = which;
    A.x[0] = x0; A.x[1] = x1; A.uf[0] = uf0; A.uf[1] = uf1; A.gradI = gradI; A.It = It; A.incr = incr;
    {   // the coefficients exactly as the reference's expression rounds them in the field precision
        // (OpticalFlowElastic.cpp:41: (1.0f-omega), omega / (-6*mu-2*lambda), (mu+lambda)); same as the strict path, solvers.cu
        const R mu = (R)S.mu, lambda = (R)S.lambda, omega = (R)S.omega;
        A.ck = (R)1.0f - omega;
        A.cr = omega / ((R)-6 * mu - (R)2 * lambda);
        A.mu = mu;
        A.mupl = mu + lambda;
        A.a = -(A.cr * A.mu);
    }
    A.neg_zero = -0.0f;
    A.ctl = ctl; A.n_active = n_active; A.partials = partials; A.pstride = pstride; A.tr = tr;
    const bool fluid = which == 1;
    const size_t stage = (size_t)(A.LR + 4) * sizeof(vec2_t<R>) + (size_t)A.LR * (sizeof(vec2_t<R>) * (fluid ? 2 : 1) + sizeof(R));
    const size_t smem = SOR_NS * stage + 2 * (size_t)S.NT * 2 * sizeof(vec2_t<R>);
    const dim3 grid(S.nbands, S.nstrips, S.batch);
    auto go = [&](auto kernel, const char *name) -> int {
        int st = of2d_ensure_dynamic_smem((const void *)kernel, smem);
        if (st) return st;
        ProfScope _ps(ctx, name);
        // measured: with the attribute the next sweep's CTAs are placed as slots free up and Elastic (sweep after sweep) loses 12 %;
        // Fluid (the sweep follows other kernels) gains 1 %
        if (fluid) pdl_launch<1>(kernel, grid, S.NT == 32 ? 32 : S.NT + 32, smem, ctx->stream, A);
        else pdl_launch<2>(kernel, grid, S.NT == 32 ? 32 : S.NT + 32, smem, ctx->stream, A);
        return OF2D_SUCCESS;
    };
    int st;
    if (fluid) st = S.NT == 32 ? go(k_sor_tile<R, 4, true, true>, "sor_tile_fluid") : go(k_sor_tile<R, 4, true, false>, "sor_tile_fluid");
    else st = S.NT == 32 ? go(k_sor_tile<R, 4, false, true>, "sor_tile_elastic") : go(k_sor_tile<R, 4, false, false>, "sor_tile_elastic");
    if (st) return st;
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}
