// engine_hs_tma.cuh -- relaxed build, fp32 fields: the two-step Horn-Schunck kernel (k_hs_pair, engine_kernels.cuh) fed by a two-stage
// tensor-map TMA pipeline (the structure of engine_fused_tma.cuh): u^k on the 36 x 36 halo tile, gradI and It on the 34 x 34 halo
// tile of a CTA's NEXT tile land in shared memory (one cp.async.bulk.tensor.3d each) while the current tile takes its two Jacobi
// steps; outside the field TMA delivers zeros, which no in-field point reads (border points take q = 0, gradients.h:72-80).
// Reference computation: OpticalFlowDiffusion.cpp:19-84, gradients.h:72-80, Logger.cpp:32-51.
#pragma once

#if OF2D_RELAXED

namespace {

constexpr int HP_H0 = TILE + 4, HP_H1 = TILE + 2;
constexpr int HP_GP = 36, HP_TP = 40;                           // pitches of the gradI tile (origin i0 - 2, j0 - 1) and of the It tile (origin i0 - 4, j0 - 1)
constexpr unsigned HP_S0 = rt_round128(HP_H0 * HP_H0 * 8), HP_SG = rt_round128(HP_GP * HP_H1 * 8), HP_SI = rt_round128(HP_TP * HP_H1 * 4);
constexpr unsigned HP_STAGE = HP_S0 + HP_SG + HP_SI, HP_TX = HP_H0 * HP_H0 * 8 + HP_GP * HP_H1 * 8 + HP_TP * HP_H1 * 4;
constexpr unsigned HP_S1 = rt_round128(HP_H1 * HP_H1 * 8), HP_SMEM = 2 * HP_STAGE + HP_S1;

__global__ void __launch_bounds__(TX *TY, OF2D_HS_MINB)
k_rt_hs_pair(EngK<float> K, float alphasq, const __grid_constant__ TmaMaps4 M) {   // M.m[0 / 1]: estimate buffers, m[2]: gradI, m[3]: It
    pdl_enter();
    using V = float2;
    constexpr int H0 = HP_H0, H1 = HP_H1, NRING = 4 * H1 - 4;
    extern __shared__ __align__(128) unsigned char smem_dynamic[];
    __shared__ uint64_t bars[2];
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const bool single = h.redo != 0;   // redo of the first step of the previous two-step launch: one step only
    const int nx = K.nx, ny = K.ny;
    V *__restrict__ un = pick(K, B_EST_NEXT, h, pair);
    const CUtensorMap *mu = &M.m[h.sel & 1], *mg = &M.m[2], *mt = &M.m[3];
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init_fence(); }
    unsigned phase = 0u;
    // the ring point of this thread (threads 0 .. 131): coordinates in the 34 x 34 halo tile
    int rr = 0, rc = 0;
    if (tid < H1) { rr = 0; rc = tid; }
    else if (tid < 2 * H1) { rr = H1 - 1; rc = tid - H1; }
    else if (tid < 2 * H1 + TILE) { rr = tid - 2 * H1 + 1; rc = 0; }
    else { rr = tid - (2 * H1 + TILE) + 1; rc = H1 - 1; }
    const bool has_ring = tid < NRING;
    const TileWalk T(nx, ny);
    NormAcc<float> acc1, acc2;
    bool divzero = false;
    auto stage_0 = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * HP_STAGE); };
    auto stage_g = [&](int s) { return reinterpret_cast<V *>(smem_dynamic + (unsigned)s * HP_STAGE + HP_S0); };
    auto stage_t = [&](int s) { return reinterpret_cast<float *>(smem_dynamic + (unsigned)s * HP_STAGE + HP_S0 + HP_SG); };
    V *s1 = reinterpret_cast<V *>(smem_dynamic + 2 * HP_STAGE);   // u^(k+1) on the 34 x 34 halo tile
    auto issue = [&](int tile, int s) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        proxy_fence_async();
        mbar_expect_tx(&bars[s], HP_TX);
        tma_load_3d(stage_0(s), mu, i0 - 2, j0 - 2, pair, &bars[s]);
        tma_load_3d(stage_g(s), mg, i0 - 2, j0 - 1, pair, &bars[s]);
        tma_load_3d(stage_t(s), mt, i0 - 4, j0 - 1, pair, &bars[s]);
    };
    __syncthreads();
    int tile = blockIdx.x;
    if (tid == 0 && tile < T.ntiles) issue(tile, 0);
    for (int k = 0; tile < T.ntiles; tile += gridDim.x, k++) {
        const int s = k & 1;
        const int next = tile + gridDim.x;
        if (tid == 0 && next < T.ntiles) issue(next, s ^ 1);
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        const int i = i0 + threadIdx.x;
        mbar_wait(&bars[s], (phase >> s) & 1u);
        phase ^= 1u << s;
        const V *s0 = stage_0(s), *sg = stage_g(s);
        const float *sit = stage_t(s);
        // gradI / It of the thread's 4 pixels (kept for the second step) and of its ring point
        V dI[PY];
        float it[PY];
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY;
            dI[p] = sg[(jl + 1) * HP_GP + threadIdx.x + 2];
            it[p] = sit[(jl + 1) * HP_TP + threadIdx.x + 4];
        }
        // first step on the 34 x 34 halo: own pixels, then the ring point
        V u1[PY];
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY, j = j0 + jl;
            const int e0 = (jl + 2) * H0 + threadIdx.x + 2;
            const bool border = i == 0 || i >= nx - 1 || j == 0 || j >= ny - 1;
            bool dz = false;
            u1[p] = hs_point<float>(s0[e0 - 1], s0[e0 + 1], s0[e0 - H0], s0[e0 + H0], border, dI[p], it[p], alphasq, dz);
            s1[(jl + 1) * H1 + threadIdx.x + 1] = u1[p];
            if (i < nx && j < ny) {
                divzero = divzero || dz;   // (points outside the field see zeros from TMA: their test means nothing)
                acc1.add(u1[p], s0[e0]);
                if (single) un[i + j * nx] = u1[p];
            }
        }
        if (!single) {
            if (has_ring) {
                const int ri = i0 - 1 + rc, rj = j0 - 1 + rr;
                const bool ring_in = ri >= 0 && ri < nx && rj >= 0 && rj < ny;
                const int e0 = (rr + 1) * H0 + rc + 1;
                const bool border = ri <= 0 || ri >= nx - 1 || rj <= 0 || rj >= ny - 1;
                bool dz = false;
                const V v = hs_point<float>(s0[e0 - 1], s0[e0 + 1], s0[e0 - H0], s0[e0 + H0], border, sg[rr * HP_GP + rc + 1], sit[rr * HP_TP + rc + 3], alphasq, dz);
                s1[rr * H1 + rc] = ring_in ? v : make_float2(0.0f, 0.0f);
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
                const V o = hs_point<float>(s1[e1 - 1], s1[e1 + 1], s1[e1 - H1], s1[e1 + H1], border, dI[p], it[p], alphasq, dz);
                if (i < nx && j < ny) {
                    un[i + j * nx] = o;
                    acc2.add(o, u1[p]);
                }
            }
            acc2.flush();
        } else acc1.flush();
        __syncthreads();   // every read of stage s and of s1 is over
    }
    hs_pair_epilogue<float>(K, c, pair, acc1, acc2, single, divzero);
}

}  // namespace

#endif  // OF2D_RELAXED
