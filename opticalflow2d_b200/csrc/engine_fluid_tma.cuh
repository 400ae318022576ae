// engine_fluid_tma.cuh -- relaxed build, fp32 fields: k_fl_integrate as a tensor-map TMA pipeline of OF2D_FI_NST stages (the structure of
// engine_fused_tma.cuh).  The halo tiles of the estimate and of the increment of a CTA's NEXT TWO interior tiles land in shared
// memory (one cp.async.bulk.tensor.3d each, SASS UTMALDG) while the current tile is integrated, so no warp waits on global
// memory -- the plain kernel spent 10.8 stall cycles per issued instruction on its loads (profiles/r2_fluid__k_fl_integrate_float.txt).
// After a regrid Logger's prev (the pre-reset estimate = the buffer this kernel overwrites) is read by the thread that owns the
// points, before it waits for the stage.
// Reference computation: OpticalFlowFluid.cpp:97-121 (explicit Euler step), Image::jacobian / min Image.cpp:189-218, 96-104,
// Logger.cpp:32-51, regrid decision ImageRegistrationFluid.cpp:99-124.  Border tiles, partial tiles and skipped steps
// (dt >= 65, OpticalFlowFluid.cpp:135-137) go through the same staged tiles (zero fill outside the field, one-sided differences at its edge).
#pragma once

#if OF2D_RELAXED

namespace {

constexpr int FI_W = 36, FI_H = TILE + 2;                       // halo tile in shared memory: [FI_H rows i][FI_W columns j], origin (j0 - 2, i0 - 1)
constexpr unsigned FI_TILE = rt_round128(FI_W * FI_H * 8);      // one array (float2)
#ifndef OF2D_FI_NST
#define OF2D_FI_NST 3    // stages of the pipeline (tiles in flight per CTA = OF2D_FI_NST - 1)
#endif
#ifndef OF2D_FI_MINB
#define OF2D_FI_MINB 3   // resident CTAs per SM the register allocation aims at (3 stages of 19.7 KB each: 3 CTAs; measured at 2048^2, 40 iterations:
                         // 2 stages / 4 CTAs 5.53 ms, 3 / 3 5.30 ms, 4 / 2 5.32 ms; 2 stages with Logger's prev staged as well: 5.44 ms)
#endif
constexpr int FI_NST = OF2D_FI_NST;
constexpr unsigned FI_STAGE = 2 * FI_TILE, FI_SMEM = FI_NST * FI_STAGE, FI_TX = 2 * FI_W * FI_H * 8;
struct TmaMaps5 { CUtensorMap m[5]; };                           // estimate buffers 0 / 1 (halo tile), increment (m[3], m[4]: tile-only maps of the estimate, unused)

__global__ void __launch_bounds__(TX *TY, OF2D_FI_MINB)
k_rt_fl_integrate(EngK<float> K, const float2 *__restrict__ incr_all, const __grid_constant__ TmaMaps5 M) {
    pdl_enter();
    using V = float2;
    extern __shared__ __align__(128) unsigned char smem_dynamic[];
    __shared__ uint64_t bars[FI_NST];
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny, P = K.P;
    V *un = pick(K, B_EST_NEXT, h, pair, true);
    const V *__restrict__ incr = incr_all + (size_t)pair * K.nT;
    const bool skip = h.skip != 0;
    const bool prev_other = h.prev_other != 0;
    const float dt = (float)__ldcg(&c->dt);
    const CUtensorMap *mu = &M.m[h.sel & 1], *mr = &M.m[2];
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < FI_NST; s++) mbar_init(&bars[s], 1);
        mbar_init_fence();
    }
    unsigned phase = 0u;
    const TileWalk T(ny, nx);
    NormAcc<float> acc;
    float mj = INFINITY;
    auto stage_u = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * FI_STAGE); };
    auto stage_r = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * FI_STAGE + FI_TILE); };
    // every tile goes through the staged halo tiles: outside the field TMA delivers zeros, which the one-sided differences of
    // the field's edge (gradients.h:9-32) never read; a skipped step (dt >= 65) integrates with dt = 0
    const float dte = skip ? 0.0f : dt;
    auto issue = [&](int tile, int s) {
        const int j0 = T.tx(tile) * TILE, i0 = T.ty(tile) * TILE;
        proxy_fence_async();
        mbar_expect_tx(&bars[s], FI_TX);
        tma_load_3d(stage_u(s), mu, j0 - 2, i0 - 1, pair, &bars[s]);
        tma_load_3d(stage_r(s), mr, j0 - 2, i0 - 1, pair, &bars[s]);
    };
    __syncthreads();
    int tile = blockIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int a = 0; a < FI_NST - 1; a++) if (tile + a * (int)gridDim.x < T.ntiles) issue(tile + a * (int)gridDim.x, a);
    }
    for (int s = 0; tile < T.ntiles; tile += gridDim.x, s = s + 1 == FI_NST ? 0 : s + 1) {
        // the stage the tile FI_NST - 1 ahead lands in was read by the previous iteration, which ended with a block barrier
        const int next = tile + (FI_NST - 1) * (int)gridDim.x;
        if (tid == 0 && next < T.ntiles) issue(next, s == 0 ? FI_NST - 1 : s - 1);
        const int j0 = T.tx(tile) * TILE, i0 = T.ty(tile) * TILE;
        const int ib = i0 + 4 * threadIdx.y, j = j0 + threadIdx.x;
        const size_t o0 = (size_t)ib * P + (size_t)j;
        V pvo[4];
        if (prev_other) {   // after a regrid Logger's prev is the pre-reset estimate: the OTHER buffer, at the points this thread is about to overwrite
#pragma unroll
            for (int q = 0; q < 4; q++) pvo[q] = (ib + q < nx && j < ny) ? __ldcg(un + o0 + (size_t)q * P) : make_float2(0.0f, 0.0f);
        }
        mbar_wait(&bars[s], (phase >> s) & 1u);
        phase ^= 1u << s;
        // a thread owns 4 consecutive i of one j: the new field u + dt R on its line i-1 .. i+4 and on the j-1 / j+1 neighbours of
        // its own four points, evaluated from the staged tiles (the same expression wherever a point is needed: the same bits)
        const V *su = stage_u(s) + (4 * threadIdx.y) * FI_W + threadIdx.x + 2;   // (line i-1 of the thread, its own j)
        const V *sr = stage_r(s) + (4 * threadIdx.y) * FI_W + threadIdx.x + 2;
        auto nw = [&](int o) { const V a = su[o], r = sr[o]; return make_float2(fmaf(r.x, dte, a.x), fmaf(r.y, dte, a.y)); };
        V ce[6], le[4], ri[4], pv[4];
#pragma unroll
        for (int r = 0; r < 6; r++) ce[r] = nw(r * FI_W);
#pragma unroll
        for (int q = 0; q < 4; q++) { le[q] = nw((q + 1) * FI_W - 1); ri[q] = nw((q + 1) * FI_W + 1); pv[q] = su[(q + 1) * FI_W]; }
        if (prev_other) {
#pragma unroll
            for (int q = 0; q < 4; q++) pv[q] = pvo[q];
        }
        if (j0 >= 1 && j0 + TILE < ny && i0 >= 1 && i0 + TILE < nx) {   // interior tile: central differences everywhere
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const V nv = ce[q + 1];
                acc.add(nv, pv[q]);
                const V dx = make_float2((ce[q + 2].x - ce[q].x) * 0.5f, (ce[q + 2].y - ce[q].y) * 0.5f);
                const V dy = make_float2((ri[q].x - le[q].x) * 0.5f, (ri[q].y - le[q].y) * 0.5f);
                const float J = (1.0f + dx.x) * (1.0f + dy.y) - dx.y * dy.x;
                mj = J < mj ? J : mj;
                un[o0 + (size_t)q * P] = nv;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = ib + q;
                if (i >= nx || j >= ny) continue;
                const V nv = ce[q + 1];
                acc.add(nv, pv[q]);
                V dx, dy;
                if (i == 0) dx = make_float2(ce[q + 2].x - nv.x, ce[q + 2].y - nv.y);
                else if (i == nx - 1) dx = make_float2(nv.x - ce[q].x, nv.y - ce[q].y);
                else dx = make_float2((ce[q + 2].x - ce[q].x) * 0.5f, (ce[q + 2].y - ce[q].y) * 0.5f);
                if (j == 0) dy = make_float2(ri[q].x - nv.x, ri[q].y - nv.y);
                else if (j == ny - 1) dy = make_float2(nv.x - le[q].x, nv.y - le[q].y);
                else dy = make_float2((ri[q].x - le[q].x) * 0.5f, (ri[q].y - le[q].y) * 0.5f);
                const float J = (1.0f + dx.x) * (1.0f + dy.y) - dx.y * dy.x;
                mj = J < mj ? J : mj;
                un[o0 + (size_t)q * P] = nv;
            }
        }
        acc.flush();
        __syncthreads();   // every read of stage s is over
    }
    fl_integrate_epilogue<float>(K, c, pair, acc, mj);
}


// ---------------------------------------------------------------------------------------------
// Fluid regrid, second half (k_fl_rewarp; ImageRegistrationFluid.cpp:116-124): Iaux = Imov o (id + level motion), derivatives of
// (Iref, Iaux) into the transposed layout, estimate <- 0 -- from staged tiles: the new level motion on the 34 x 34 halo tile, the
// 48 x 48 window of Imov the bilinear taps read (displacements beyond ~6 px: Image::warp2d from global memory), Iref on the tile.
// ---------------------------------------------------------------------------------------------
constexpr int RW_H = TILE + 2, RW_LP = 36;                       // halo tile of the level motion: [RW_H][RW_LP] float2, origin (i0 - 2, j0 - 1)
constexpr unsigned RW_SL = rt_round128(RW_LP * RW_H * 8), RW_SI = FW * FW * 4, RW_SR = TILE * TILE * 4;
constexpr unsigned RW_STAGE = RW_SL + RW_SI + RW_SR, RW_TX = RW_LP * RW_H * 8 + FW * FW * 4 + TILE * TILE * 4;
constexpr unsigned RW_SW = rt_round128(RW_H * (RW_H + 1) * 4), RW_SG = rt_round128(TILE * (TILE + 1) * 8), RW_ST = rt_round128(TILE * (TILE + 1) * 4);
constexpr unsigned RW_SMEM = 2 * RW_STAGE + RW_SW + RW_SG + RW_ST;

__device__ __noinline__ float rt_rewarp_slow(const float *__restrict__ Imov, int nx, int ny, int i, int j, float ux, float uy) {
    return warp_pixel_lazy<float>(Imov, nx, ny, i, j, make_float2(ux, uy), i + j * nx);
}

__global__ void __launch_bounds__(TX *TY, 3)
k_rt_fl_rewarp(EngK<float> K, int gate, const float *__restrict__ Iref_all, const float *__restrict__ Imov_all, float2 *__restrict__ gradI_all, float *__restrict__ It_all, int zero_buf,
               const __grid_constant__ TmaMaps4 M) {   // M.m[0 / 1]: level-motion buffers, m[2]: Imov window, m[3]: Iref tile
    pdl_enter();
    using V = float2;
    extern __shared__ __align__(128) unsigned char smem_dynamic[];
    __shared__ uint64_t bars[2];
    const int pair = blockIdx.y;
    const CtlHot h = load_ctl(K.ctl + pair);
    if (!gate_open(h, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const float *__restrict__ Imov = Imov_all + (size_t)pair * K.n;
    V *__restrict__ zero = pick(K, zero_buf, h, pair, true);
    const CUtensorMap *ml = &M.m[(h.msel & 1) ^ 1], *mi = &M.m[2], *mr = &M.m[3];   // B_LVL_NEXT
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init_fence(); }
    unsigned phase = 0u;
    const TileWalk T(nx, ny);
    auto stage_l = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * RW_STAGE); };
    auto stage_i = [&](int s) { return reinterpret_cast<float *>(smem_dynamic + (unsigned)s * RW_STAGE + RW_SL); };
    auto stage_r = [&](int s) { return reinterpret_cast<float *>(smem_dynamic + (unsigned)s * RW_STAGE + RW_SL + RW_SI); };
    float (*sw)[RW_H + 1] = reinterpret_cast<float (*)[RW_H + 1]>(smem_dynamic + 2 * RW_STAGE);
    V (*sg)[TILE + 1] = reinterpret_cast<V (*)[TILE + 1]>(smem_dynamic + 2 * RW_STAGE + RW_SW);
    float (*st)[TILE + 1] = reinterpret_cast<float (*)[TILE + 1]>(smem_dynamic + 2 * RW_STAGE + RW_SW + RW_SG);
    auto issue = [&](int tile, int s) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        proxy_fence_async();
        mbar_expect_tx(&bars[s], RW_TX);
        tma_load_3d(stage_l(s), ml, i0 - 2, j0 - 1, pair, &bars[s]);
        tma_load_3d(stage_i(s), mi, i0 - FO, j0 - FO, pair, &bars[s]);
        tma_load_3d(stage_r(s), mr, i0, j0, pair, &bars[s]);
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
        const V *sl = stage_l(s);
        const float *si = stage_i(s), *sr = stage_r(s);
        // taps inside the window AND inside the image: window positions [lx0, lx0 + lxn) x [ly0, ly0 + lyn)
        const TapBox tb(i0 - FO, j0 - FO, nx, ny);
        for (int e = tid; e < RW_H * RW_H; e += TX * TY) {
            const int r = e / RW_H, cc = e - r * RW_H;
            const int i = i0 + cc - 1, j = j0 + r - 1;
            float w = 0.0f;
            if (i >= 0 && i < nx && j >= 0 && j < ny) {
                const V u = sl[r * RW_LP + cc + 1];
                const float px = (float)i + u.x, flx = floorf(px), py = (float)j + u.y, fly = floorf(py);
                const int lx = (int)flx - (i0 - FO), ly = (int)fly - (j0 - FO);
                const float fx = px - flx, fy = py - fly;
                if ((unsigned)(lx - tb.lx0) < tb.lxn && (unsigned)(ly - tb.ly0) < tb.lyn) {
                    const float *p = si + ly * FW + lx;
                    const float s00 = p[0], s10 = p[1], s01 = p[FW], s11 = p[FW + 1];
                    const float lo = s00 + fx * (s10 - s00), hi = s01 + fx * (s11 - s01);
                    w = lo + fy * (hi - lo);
                } else w = rt_rewarp_slow(Imov, nx, ny, i, j, u.x, u.y);
            }
            sw[r][cc] = w;
        }
        __syncthreads();
        const int i = i0 + threadIdx.x;
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY, j = j0 + jl;
            if (i < nx && j < ny) {
                const int r = jl + 1, cc = threadIdx.x + 1;
                const float ce = sw[r][cc];
                float gx, gy;   // gradients.h:9-32 on the warped image
                if (i == 0) gx = sw[r][cc + 1] - ce;
                else if (i == nx - 1) gx = ce - sw[r][cc - 1];
                else gx = (sw[r][cc + 1] - sw[r][cc - 1]) * 0.5f;
                if (j == 0) gy = sw[r + 1][cc] - ce;
                else if (j == ny - 1) gy = ce - sw[r - 1][cc];
                else gy = (sw[r + 1][cc] - sw[r - 1][cc]) * 0.5f;
                sg[jl][threadIdx.x] = make_float2(gx, gy);
                st[jl][threadIdx.x] = ce - sr[jl * TILE + threadIdx.x];
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
                zero[o] = make_float2(0.0f, 0.0f);
            }
        }
        __syncthreads();   // every read of stage s and of the tile arrays is over
    }
}

}  // namespace

#endif  // OF2D_RELAXED
