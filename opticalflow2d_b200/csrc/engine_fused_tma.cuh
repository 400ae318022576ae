// engine_fused_tma.cuh -- relaxed build, fp32 fields: the two fused Demons kernels of engine_fused.cuh as a TWO-STAGE
// TMA PIPELINE.  Every input of an interior tile arrives through ONE tensor-map TMA load per array
// (cp.async.bulk.tensor.3d, SASS UTMALDG; the third coordinate is the pair of a batch) into a double-buffered stage, and
// the loads of a CTA's NEXT tile are issued before the current tile is computed, so no warp waits on global memory:
//
//   k_rt_force_conv    stage = { Imov window 48 x 48, u on the warped-image window, Iref on the correspondence window }
//   k_rt_compose_conv  stage = { window of the current estimate 48 x 48 float2, the correspondence on the composed window
//                                (overwritten in place by the composed field) }
//
// The arithmetic of an interior tile is the FAST instance of engine_fused.cuh (no validity / wrap / border tests); border
// tiles (about 6 % at 2048^2, 23 % at 512^2) run that file's general instance on the same shared arrays -- with the source window
// taken from the same prefetch (zeros outside the field, which no tap box reaches) -- so the flat-index semantics of the
// reference's convolution (Field.tpp:245-248) and Image::warp2d's border rules stay where they were.
// Reference computation: DemonsThirions.cpp:18-42, DemonsDiffeomorphic.cpp:15-30, Demons.cpp:34-63, Motion.cpp:113-178.
#pragma once

#if OF2D_RELAXED

#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

namespace {

#ifndef OF2D_RT_MINB
#define OF2D_RT_MINB 3   // resident CTAs per SM the register allocation aims at (4 = 64 registers: Thirion 4.86 -> 5.30 ms, measured)
#endif
struct TmaMaps4 { CUtensorMap m[4]; };

// one tensor-map TMA load of a {bx, by, 1} box at (x, y, z); out-of-range elements arrive as zeros
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

constexpr unsigned rt_round128(unsigned b) { return (b + 127u) & ~127u; }
constexpr int RT_TW = 40;   // columns of the u / Iref tiles of the force kernel (origin i0 - 4)
constexpr int RT_VP = 36;   // columns of the correspondence tile of the compose kernel (origin i0 - 2).  Box rows are kept multiples of
                            // 32 bytes: a {34, 34} box of 8-byte elements (272-byte rows) raised "illegal instruction" on sm_100a

template <int KW> struct RtGeom {
    using G = FusedGeom<KW>;
    static constexpr int CX = G::CX, HW = G::HW, WW = G::WW, WP = G::WP, CW = G::CW;
    // compose kernel: stage = window of u (float2 [FW][FW]) + v on the composed window (float2 [CW][CW])
    static constexpr int VOFF = 2 - CX;   // column of the composed window's first element in the [CW][RT_VP] tile
    static constexpr unsigned C_SU = FW * FW * 8, C_SV = rt_round128(CW * RT_VP * 8), C_STAGE = C_SU + C_SV;
    static constexpr unsigned C_TX = FW * FW * 8 + CW * RT_VP * 8;
    // force kernel: stage = Imov window (float [FW][FW]) + u (float2 [WW][RT_TW]) + Iref (float [CW][RT_TW])
    static constexpr unsigned F_SI = FW * FW * 4, F_SU = rt_round128(WW * RT_TW * 8), F_SR = rt_round128(CW * RT_TW * 4), F_STAGE = F_SI + F_SU + F_SR;
    static constexpr unsigned F_TX = FW * FW * 4 + WW * RT_TW * 8 + CW * RT_TW * 4;
    static constexpr unsigned F_SW = rt_round128(WW * WP * 4), F_SC = rt_round128(CW * CW * 8);
    static constexpr unsigned C_SMEM = 2 * C_STAGE, F_SMEM = 2 * F_STAGE + F_SW + F_SC;
};

// a tile whose 48 x 48 window lies inside the image: every tap box is the whole window, every flat index valid
__device__ __forceinline__ bool rt_tile_fast(int i0, int j0, int nx, int ny) {
    return i0 >= FO && j0 >= FO && i0 - FO + FW <= nx && j0 - FO + FW <= ny;
}

// Window elements -> threads, balanced over the 8 warps: the 32 x 32 core of a (32 + 2H)^2 window goes as "column H + tx,
// rows ty + 8k, k < 4" (aligned row segments, constant strides); the other elements (the rows under the core, then the two
// side strips) are dealt out as q = tid + 256 m, so every warp handles 4 + NX / 256 elements (+-1).
template <int H> struct Spread {
    static constexpr int W = TILE + 2 * H;
    static constexpr int NB = (W - TILE) * TILE;   // rows 32 .. W-1 of the core columns
    static constexpr int NX = NB + 2 * H * W;      // + the side strips
    static constexpr int NM = (NX + TX * TY - 1) / (TX * TY);
    int rc[NM];                                     // (row << 8 | column) of the thread's m-th extra element, -1: none
    __device__ __forceinline__ explicit Spread(int tid) {
#pragma unroll
        for (int m = 0; m < NM; m++) {
            int q = tid + TX * TY * m;
            if (q >= NX) rc[m] = -1;
            else if (q < NB) rc[m] = ((TILE + (q >> 5)) << 8) | (H + (q & 31));
            else { q -= NB; const int r = q / (2 * H), t = q - r * (2 * H); rc[m] = (r << 8) | (t < H ? t : t + TILE); }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// compose + smoothing + Logger
// ---------------------------------------------------------------------------------------------
// elements whose taps left the staged window (correspondences of several pixels): Motion::accumulate from global memory
template <int KW>
__device__ __noinline__ void rt_compose_fixup(unsigned bad, float2 *sV, const float2 *__restrict__ u, int nx, int ny, int i0, int j0, const Spread<RtGeom<KW>::CX> X) {
    constexpr int CX = RtGeom<KW>::CX;
    for (int bit = 0; bad >> bit; bit++) {
        if (!(bad >> bit & 1u)) continue;
        int r = threadIdx.y + TY * bit, cc = CX + threadIdx.x;
        if (bit >= 4) { r = X.rc[bit - 4 < Spread<CX>::NM ? bit - 4 : 0] >> 8; cc = X.rc[bit - 4 < Spread<CX>::NM ? bit - 4 : 0] & 255; }
        float2 *pv = sV + r * RT_VP + cc;
        const float2 v = *pv;
        *pv = compose_slow<float>(u, nx, ny, i0 - CX + cc, j0 - CX + r, v.x, v.y);
    }
}

template <int KW>
struct RtComposeTile {
    using V = float2;
    using RG = RtGeom<KW>;
    static constexpr int CX = RG::CX, CW = RG::CW;

    // v + u o (id + v) with the four taps from the staged window; returns true (and leaves v in place) when a tap is outside
    static __device__ __forceinline__ bool elem(V *sV, const V *sU, int i0, int j0, int r, int cc, float xf, float yf) {
        V *pv = sV + r * RT_VP + cc;
        const V v = *pv;
        const float px = xf + v.x, flx = floorf(px), py = yf + v.y, fly = floorf(py);
        const int lx = (int)flx - (i0 - FO), ly = (int)fly - (j0 - FO);
        const float fx = px - flx, fy = py - fly;
        const bool ok = (unsigned)lx < (unsigned)(FW - 1) && (unsigned)ly < (unsigned)(FW - 1);
        const V *p = sU + (ok ? ly * FW + lx : 0);
        const V s00 = p[0], s10 = p[1], s01 = p[FW], s11 = p[FW + 1];
        const float lx_ = s00.x + fx * (s10.x - s00.x), hx_ = s01.x + fx * (s11.x - s01.x);
        const float ly_ = s00.y + fx * (s10.y - s00.y), hy_ = s01.y + fx * (s11.y - s01.y);
        if (ok) *pv = make_float2(v.x + (lx_ + fy * (hx_ - lx_)), v.y + (ly_ + fy * (hy_ - ly_)));
        return !ok;
    }
    static __device__ __forceinline__ void elem_add(V *sV, const V *sU, int r, int cc) {
        V *pv = sV + r * RT_VP + cc;
        const V v = *pv, uc = sU[(r - CX + FO) * FW + (cc - CX + FO)];
        *pv = make_float2(uc.x + v.x, uc.y + v.y);
    }
    static __device__ __forceinline__ void run(const V *sU, V *sVtile, const Spread<CX> &X, const V *__restrict__ u, V *__restrict__ out, int nx, int ny, int i0, int j0,
                                               int add_only, const ConvW<float> &W, NormAcc<float> &acc) {
        const int tx = threadIdx.x, ty = threadIdx.y;
        V *sV = sVtile + RG::VOFF;   // element (r, cc) of the composed window at sV[r * RT_VP + cc]
        if (add_only) {
#pragma unroll
            for (int k = 0; k < 4; k++) elem_add(sV, sU, ty + TY * k, CX + tx);
#pragma unroll
            for (int m = 0; m < Spread<CX>::NM; m++) if (X.rc[m] >= 0) elem_add(sV, sU, X.rc[m] >> 8, X.rc[m] & 255);
        } else {
            const float xf = (float)(i0 + tx), yf = (float)(j0 - CX + ty);
            unsigned bad = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) bad |= elem(sV, sU, i0, j0, ty + TY * k, CX + tx, xf, yf + (float)(TY * k)) ? 1u << k : 0u;
#pragma unroll
            for (int m = 0; m < Spread<CX>::NM; m++)
                if (X.rc[m] >= 0) {
                    const int r = X.rc[m] >> 8, cc = X.rc[m] & 255;
                    bad |= elem(sV, sU, i0, j0, r, cc, (float)(i0 - CX + cc), (float)(j0 - CX + r)) ? 1u << (4 + m) : 0u;
                }
            if (bad) rt_compose_fixup<KW>(bad, sV, u, nx, ny, i0, j0, X);
        }
        __syncthreads();
        const int i = i0 + tx, jb = j0 + 4 * ty;
        V *op = out + (i + jb * nx);
        const V *prev = sU + (4 * ty + FO) * FW + tx + FO;   // Logger's prev: the current estimate at the pixel (in the window)
        fused_conv_p<float, KW, RT_VP>(sV, W, i0, j0, nx, ny, (long)nx * ny, true, [&](int q, V o) {
            op[q * nx] = o;
            acc.add(o, prev[q * FW]);
        });
        acc.flush();
    }
};

template <int KW>
__global__ void __launch_bounds__(TX *TY, OF2D_RT_MINB)
k_rt_compose_conv(EngK<float> K, int v_buf, int add_only, const __grid_constant__ ConvW<float> W, const __grid_constant__ TmaMaps4 M) {
    pdl_enter();
    using V = float2;
    using RG = RtGeom<KW>;
    using G = FusedGeom<KW>;
    static_assert(G::CP == G::CW && G::CW <= RT_VP, "the general instance works on a dense [CW][CW] array inside the stage's [CW][RT_VP] tile");
    extern __shared__ __align__(128) unsigned char smem_dynamic[];
    __shared__ uint64_t bars[3];   // full[0], full[1], the general instance's own barrier
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny, n = (int)K.n;
    const V *__restrict__ u = pick(K, B_EST_CUR, h, pair);
    const V *__restrict__ v = pick(K, v_buf, h, pair);
    V *__restrict__ out = pick(K, B_EST_NEXT, h, pair);
    const CUtensorMap *mu = &M.m[h.sel & 1];
    const CUtensorMap *mv = &M.m[2 + (v_buf == B_C0 ? 0 : v_buf == B_C1 ? 1 : (h.nsquares & 1) ? 0 : 1)];
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init_fence(); }
    unsigned phase = 0u, uses = 0u;   // bit s of phase: parity the next wait on full[s] expects
    const Strip<G::CX> SC_(tid);
    const Spread<G::CX> X(tid);
    const TileWalk T(nx, ny);
    NormAcc<float> acc;
    auto stage_u = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * RG::C_STAGE); };
    auto stage_v = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * RG::C_STAGE + RG::C_SU); };
    auto issue = [&](int tile, int s) {   // thread 0, after a barrier that ended every read of stage s
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;   // (border tiles too: the general instance takes the window from here)
        proxy_fence_async();
        mbar_expect_tx(&bars[s], RG::C_TX);
        tma_load_3d(stage_u(s), mu, i0 - FO, j0 - FO, pair, &bars[s]);
        tma_load_3d(stage_v(s), mv, i0 - 2, j0 - G::CX, pair, &bars[s]);
    };
    __syncthreads();   // the barriers are initialised
    int tile = blockIdx.x;
    if (tid == 0 && tile < T.ntiles) issue(tile, 0);
    for (int k = 0; tile < T.ntiles; tile += gridDim.x, k++) {
        const int s = k & 1;
        const int next = tile + gridDim.x;
        if (tid == 0 && next < T.ntiles) issue(next, s ^ 1);
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        mbar_wait(&bars[s], (phase >> s) & 1u);
        phase ^= 1u << s;
        if (rt_tile_fast(i0, j0, nx, ny)) RtComposeTile<KW>::run(stage_u(s), stage_v(s), X, u, out, nx, ny, i0, j0, add_only, W, acc);
        else ComposeConvTile<float, KW, false>::run(stage_u(s), stage_v(s), &bars[2], uses, SC_, u, v, out, nx, ny, n, i0, j0, add_only, W, acc, true);
        __syncthreads();   // every read of stage s is over: the tile after the next may land there
    }
    logger_epilogue<float>(K, c, pair, acc.dsd, acc.dsp);
}

// ---------------------------------------------------------------------------------------------
// warp -> derivatives -> demons force -> smoothing (EPI 2: + maxabs -> number of squarings)
// ---------------------------------------------------------------------------------------------
// warped-image elements whose taps left the staged window (estimates of several pixels): Image::warp2d from global memory
template <int KW>
__device__ __noinline__ void rt_warp_fixup(unsigned bad, float *sW, const float2 *sUt, const float *__restrict__ Imov, int nx, int ny, int i0, int j0, const Spread<RtGeom<KW>::HW> X) {
    constexpr int HW = RtGeom<KW>::HW, WP = RtGeom<KW>::WP;
    for (int bit = 0; bad >> bit; bit++) {
        if (!(bad >> bit & 1u)) continue;
        int r = threadIdx.y + TY * bit, cc = HW + threadIdx.x;
        if (bit >= 4) { r = X.rc[bit - 4 < Spread<HW>::NM ? bit - 4 : 0] >> 8; cc = X.rc[bit - 4 < Spread<HW>::NM ? bit - 4 : 0] & 255; }
        const float2 uu = sUt[r * RT_TW + cc + (4 - HW)];
        sW[r * WP + cc] = warp_slow<float>(Imov, nx, ny, i0 - HW + cc, j0 - HW + r, uu.x, uu.y);
    }
}

template <int KW>
struct RtForceTile {
    using V = float2;
    using RG = RtGeom<KW>;
    static constexpr int CX = RG::CX, HW = RG::HW, WW = RG::WW, WP = RG::WP, CW = RG::CW;

    static __device__ __forceinline__ bool warp_elem(float *sW, const float *sI, const V *sUt, int i0, int j0, int r, int cc, float xf, float yf) {
        const V uu = sUt[r * RT_TW + cc + (4 - HW)];
        const float px = xf + uu.x, flx = floorf(px), py = yf + uu.y, fly = floorf(py);
        const int lx = (int)flx - (i0 - FO), ly = (int)fly - (j0 - FO);
        const float fx = px - flx, fy = py - fly;
        const bool ok = (unsigned)lx < (unsigned)(FW - 1) && (unsigned)ly < (unsigned)(FW - 1);
        const float *p = sI + (ok ? ly * FW + lx : 0);
        const float s00 = p[0], s10 = p[1], s01 = p[FW], s11 = p[FW + 1];
        const float lo = s00 + fx * (s10 - s00), hi = s01 + fx * (s11 - s01);
        sW[r * WP + cc] = lo + fy * (hi - lo);
        return !ok;
    }
    static __device__ __forceinline__ void force_elem(V *sC, const float *sW, const float *sR, int r, int cc, float sratio, bool &divzero) {
        const float *w = sW + (r + 1) * WP + (cc + 1);
        const float ce = w[0];
        const float gx = (w[1] - w[-1]) * 0.5f, gy = (w[WP] - w[-WP]) * 0.5f;
        const float It = ce - sR[r * RT_TW + cc + (4 - CX)];
        const float den = gx * gx + gy * gy + (It * It) * sratio;
        V cv = make_float2(0.0f, 0.0f);
        if (den == 0) divzero = true;
        else { const float s = __fdividef(-It, den); cv = make_float2(gx * s, gy * s); }
        sC[r * CW + cc] = cv;
    }
    template <int EPI>
    static __device__ __forceinline__ void run(const float *sI, const V *sUt, const float *sR, float *sW, V *sC, const Spread<HW> &XW, const Spread<CX> &XC,
                                               const float *__restrict__ Imov, V *__restrict__ out, int nx, int ny, int i0, int j0, float sratio, const ConvW<float> &W,
                                               bool &divzero, float &mx) {
        const int tx = threadIdx.x, ty = threadIdx.y;
        {
            const float xf = (float)(i0 + tx), yf = (float)(j0 - HW + ty);
            unsigned bad = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) bad |= warp_elem(sW, sI, sUt, i0, j0, ty + TY * k, HW + tx, xf, yf + (float)(TY * k)) ? 1u << k : 0u;
#pragma unroll
            for (int m = 0; m < Spread<HW>::NM; m++)
                if (XW.rc[m] >= 0) {
                    const int r = XW.rc[m] >> 8, cc = XW.rc[m] & 255;
                    bad |= warp_elem(sW, sI, sUt, i0, j0, r, cc, (float)(i0 - HW + cc), (float)(j0 - HW + r)) ? 1u << (4 + m) : 0u;
                }
            if (bad) rt_warp_fixup<KW>(bad, sW, sUt, Imov, nx, ny, i0, j0, XW);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) force_elem(sC, sW, sR, ty + TY * k, CX + tx, sratio, divzero);
#pragma unroll
        for (int m = 0; m < Spread<CX>::NM; m++) if (XC.rc[m] >= 0) force_elem(sC, sW, sR, XC.rc[m] >> 8, XC.rc[m] & 255, sratio, divzero);
        __syncthreads();
        const int i = i0 + tx, jb = j0 + 4 * ty;
        V *op = out + (i + jb * nx);
        fused_conv<float, KW>(sC, W, i0, j0, nx, ny, (long)nx * ny, true, [&](int q, V o) {
            op[q * nx] = o;
            if (EPI == 2) { const float s = maxabs_term<float>(o); mx = mx < s ? s : mx; }
        });
    }
};

template <int EPI, int KW>
__global__ void __launch_bounds__(TX *TY, OF2D_RT_MINB)
k_rt_force_conv(EngK<float> K, const float *__restrict__ Iref_all, const float *__restrict__ Imov_all, float sratio, const __grid_constant__ ConvW<float> W, int dst_buf, int nsq_cap,
                const __grid_constant__ TmaMaps4 M) {
    pdl_enter();
    using V = float2;
    using RG = RtGeom<KW>;
    using G = FusedGeom<KW>;
    extern __shared__ __align__(128) unsigned char smem_dynamic[];
    __shared__ uint64_t bars[3];
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny, n = (int)K.n;
    const float *__restrict__ Iref = Iref_all + (size_t)pair * K.n;
    const float *__restrict__ Imov = Imov_all + (size_t)pair * K.n;
    const V *__restrict__ u = pick(K, B_EST_CUR, h, pair);
    V *__restrict__ out = pick(K, dst_buf, h, pair);
    const CUtensorMap *mu = &M.m[h.sel & 1], *mi = &M.m[2], *mr = &M.m[3];
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init_fence(); }
    unsigned phase = 0u, uses = 0u;
    const Strip<G::HW> SW_(tid);
    const Strip<G::CX> SC_(tid);
    const Spread<G::HW> XW(tid);
    const Spread<G::CX> XC(tid);
    const TileWalk T(nx, ny);
    bool divzero = false;
    float mx = 0.0f;
    auto stage_i = [&](int s) { return reinterpret_cast<float *>(smem_dynamic + (unsigned)s * RG::F_STAGE); };
    auto stage_u = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * RG::F_STAGE + RG::F_SI); };
    auto stage_r = [&](int s) { return reinterpret_cast<float *>(smem_dynamic + (unsigned)s * RG::F_STAGE + RG::F_SI + RG::F_SU); };
    float *sW = reinterpret_cast<float *>(smem_dynamic + 2 * RG::F_STAGE);
    V *sC = reinterpret_cast<V *>(smem_dynamic + 2 * RG::F_STAGE + RG::F_SW);
    auto issue = [&](int tile, int s) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        proxy_fence_async();
        mbar_expect_tx(&bars[s], RG::F_TX);
        tma_load_3d(stage_i(s), mi, i0 - FO, j0 - FO, pair, &bars[s]);
        tma_load_3d(stage_u(s), mu, i0 - 4, j0 - G::HW, pair, &bars[s]);
        tma_load_3d(stage_r(s), mr, i0 - 4, j0 - G::CX, pair, &bars[s]);
    };
    __syncthreads();
    int tile = blockIdx.x;
    if (tid == 0 && tile < T.ntiles) issue(tile, 0);
    for (int k = 0; tile < T.ntiles; tile += gridDim.x, k++) {
        const int s = k & 1;
        const int next = tile + gridDim.x;
        if (tid == 0 && next < T.ntiles) issue(next, s ^ 1);
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        mbar_wait(&bars[s], (phase >> s) & 1u);
        phase ^= 1u << s;
        if (rt_tile_fast(i0, j0, nx, ny)) RtForceTile<KW>::template run<EPI>(stage_i(s), stage_u(s), stage_r(s), sW, sC, XW, XC, Imov, out, nx, ny, i0, j0, sratio, W, divzero, mx);
        else   // the general instance of engine_fused.cuh on the staged window (its other inputs come through plain loads with the flat bounds test)
            ForceConvTile<float, KW, false>::template run<EPI>(stage_i(s), sC, sW, &bars[2], uses, SW_, SC_, u, Iref, Imov, out, nx, ny, n, i0, j0, sratio, W, divzero, mx, true);
        __syncthreads();
    }
    if (divzero) atomicOr(&c->flags, OF2D_FLAG_DIVZERO);
    if (EPI == 2) demons_nsquares_epilogue<float>(K, c, pair, mx, nsq_cap);
}

// ---- host: tensor maps -------------------------------------------------------------------------------------------
typedef CUresult (*of2d_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline of2d_encode_tiled_fn tensor_map_encoder() {
    static of2d_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (of2d_encode_tiled_fn)p;
        else cudaGetLastError();
    }
    return fn;
}
// `batch` images of `rows` lines of `line` elements of `esize` bytes (4: float, 8: float2), `img` elements from one image to the next,
// box {bx, by, 1}; false when the layout does not meet TMA's alignment rules (base and strides multiples of 16 bytes) or the
// driver has no encoder
inline bool make_field_map_pitched(CUtensorMap *m, const void *base, int esize, int line, int rows, size_t img_elems, int batch, int bx, int by) {
    const of2d_encode_tiled_fn enc = tensor_map_encoder();
    if (!enc || !base) return false;
    const size_t row = (size_t)line * esize, img = img_elems * esize;
    if (((uintptr_t)base & 15u) || (row & 15u) || (img & 15u) || ((size_t)bx * esize & 31u) || bx > 256 || by > 256) return false;   // (32-byte box rows: see RT_VP)
    const cuuint64_t dims[3] = {(cuuint64_t)line, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)row, (cuuint64_t)img};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(m, esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// field of `batch` images of nx x ny elements, image after image
inline bool make_field_map(CUtensorMap *m, const void *base, int esize, int nx, int ny, int batch, int bx, int by) {
    return make_field_map_pitched(m, base, esize, nx, ny, (size_t)nx * ny, batch, bx, by);
}

}  // namespace

#endif  // OF2D_RELAXED
