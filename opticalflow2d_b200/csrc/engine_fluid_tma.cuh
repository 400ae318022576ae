// engine_fluid_tma.cuh -- relaxed build, fp32 fields: k_fl_integrate as a two-stage tensor-map TMA pipeline (the structure of
// engine_fused_tma.cuh).  The halo tiles of the estimate and of the increment of a CTA's NEXT interior tile land in shared
// memory (one cp.async.bulk.tensor.3d each, SASS UTMALDG) while the current tile is integrated, so no warp waits on global
// memory -- the plain kernel spent 10.8 stall cycles per issued instruction on its loads (profiles/r2_fluid__k_fl_integrate_float.txt).
// Reference computation: OpticalFlowFluid.cpp:97-121 (explicit Euler step), Image::jacobian / min Image.cpp:189-218, 96-104,
// Logger.cpp:32-51, regrid decision ImageRegistrationFluid.cpp:99-124.  Border tiles, partial tiles and skipped steps
// (dt >= 65, OpticalFlowFluid.cpp:135-137) go through the same staged tiles (zero fill outside the field, one-sided differences at its edge).
#pragma once

#if OF2D_RELAXED

namespace {

constexpr int FI_W = 36, FI_H = TILE + 2;                       // halo tile in shared memory: [FI_H rows i][FI_W columns j], origin (j0 - 2, i0 - 1)
constexpr unsigned FI_TILE = rt_round128(FI_W * FI_H * 8);      // one array (float2)
constexpr unsigned FI_PREV = TILE * TILE * 8;                    // after a regrid: Logger's prev (the pre-reset estimate, the OTHER buffer) on the tile
constexpr unsigned FI_STAGE = 2 * FI_TILE + FI_PREV, FI_SMEM = 2 * FI_STAGE, FI_TX = 2 * FI_W * FI_H * 8;
struct TmaMaps5 { CUtensorMap m[5]; };                           // estimate buffers 0 / 1 (halo tile), increment, estimate buffers 0 / 1 (tile only)

__global__ void __launch_bounds__(TX *TY, 4)
k_rt_fl_integrate(EngK<float> K, const float2 *__restrict__ incr_all, const __grid_constant__ TmaMaps5 M) {
    pdl_enter();
    using V = float2;
    extern __shared__ __align__(128) unsigned char smem_dynamic[];
    __shared__ uint64_t bars[2];
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny, P = K.P;
    const V *__restrict__ u = pick(K, B_EST_CUR, h, pair, true);
    V *un = pick(K, B_EST_NEXT, h, pair, true);
    const V *__restrict__ incr = incr_all + (size_t)pair * K.nT;
    const bool skip = h.skip != 0;
    const bool prev_other = h.prev_other != 0;
    const float dt = (float)__ldcg(&c->dt);
    const CUtensorMap *mu = &M.m[h.sel & 1], *mr = &M.m[2], *mp = &M.m[3 + ((h.sel & 1) ^ 1)];
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init_fence(); }
    unsigned phase = 0u;
    const TileWalk T(ny, nx);
    NormAcc<float> acc;
    float mj = INFINITY;
    auto stage_u = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * FI_STAGE); };
    auto stage_r = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * FI_STAGE + FI_TILE); };
    auto stage_p = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * FI_STAGE + 2 * FI_TILE); };
    // every tile goes through the staged halo tiles: outside the field TMA delivers zeros, which the one-sided differences of
    // the field's edge (gradients.h:9-32) never read; a skipped step (dt >= 65) integrates with dt = 0
    const float dte = skip ? 0.0f : dt;
    auto issue = [&](int tile, int s) {
        const int j0 = T.tx(tile) * TILE, i0 = T.ty(tile) * TILE;
        proxy_fence_async();
        mbar_expect_tx(&bars[s], FI_TX + (prev_other ? FI_PREV : 0u));
        tma_load_3d(stage_u(s), mu, j0 - 2, i0 - 1, pair, &bars[s]);
        tma_load_3d(stage_r(s), mr, j0 - 2, i0 - 1, pair, &bars[s]);
        if (prev_other) tma_load_3d(stage_p(s), mp, j0, i0, pair, &bars[s]);
    };
    __syncthreads();
    int tile = blockIdx.x;
    if (tid == 0 && tile < T.ntiles) issue(tile, 0);
    for (int k = 0; tile < T.ntiles; tile += gridDim.x, k++) {
        const int s = k & 1;
        const int next = tile + gridDim.x;
        if (tid == 0 && next < T.ntiles) issue(next, s ^ 1);
        const int j0 = T.tx(tile) * TILE, i0 = T.ty(tile) * TILE;
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
        if (prev_other) {   // after a regrid Logger's prev is the pre-reset estimate
            const V *sp = stage_p(s) + (4 * threadIdx.y) * TILE + threadIdx.x;
#pragma unroll
            for (int q = 0; q < 4; q++) pv[q] = sp[q * TILE];
        }
        const int ib = i0 + 4 * threadIdx.y, j = j0 + threadIdx.x;
        const size_t o0 = (size_t)ib * P + (size_t)j;
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

}  // namespace

#endif  // OF2D_RELAXED
