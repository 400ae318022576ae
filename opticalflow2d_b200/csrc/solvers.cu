// solvers.cu -- per-iteration solver kernels (layer L2 of SURVEY.md): Horn-Schunck Jacobi step,
// Demons force (fused warp + derivatives + force), lexicographic SOR sweep as a t = 2i + j
// wavefront (elastic and fluid), fluid increment.
#include <math.h>
#include <string.h>

#include "device_math.cuh"

namespace {

constexpr int TX = 32, TY = 8;
inline dim3 grid2d(int nx, int ny, int batch) { return dim3(ceil_div(nx, TX), ceil_div(ny, TY), batch); }

// ---------------------------------------------------------------------------------------------
// Horn-Schunck Jacobi (OpticalFlowDiffusion.cpp:19-84): ubar = qlaplacian(u); f = force(ubar);
// u' = ubar - f / (alpha^2 + |gradI|^2).  Reads u 8 + gradI 8 + It 4, writes 8 B/pixel.
// ---------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(TX *TY) k_diffusion_step(int nx, int ny, const vec2_t<R> *__restrict__ u, vec2_t<R> *__restrict__ unew,
                                                           const vec2_t<R> *__restrict__ gradI, const R *__restrict__ It, R alphasq,
                                                           unsigned *__restrict__ status) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t off = (size_t)blockIdx.z * nx * ny;
    u += off; unew += off; gradI += off; It += off;
    const int idx = i + j * nx;
    vec2_t<R> q;
    if (i == 0 || i == nx - 1 || j == 0 || j == ny - 1) {
        q = mk2<R>((R)0.0f, (R)0.0f);
    } else {   // gradients.h:78
        const vec2_t<R> a = u[idx - 1], b = u[idx + 1], c = u[idx - nx], d = u[idx + nx];
        q = mk2<R>((((a.x + b.x) + c.x) + d.x) / (R)4.0f, (((a.y + b.y) + c.y) + d.y) / (R)4.0f);
    }
    const vec2_t<R> dI = gradI[idx];
    const vec2_t<R> f = lssd_force<R>(dI, It[idx], q);
    const R den = alphasq + dI.x * dI.x + dI.y * dI.y;
    if (den == 0) { atomicOr(&status[blockIdx.z], OF2D_FLAG_DIVZERO); unew[idx] = q; return; }
    unew[idx] = mk2<R>(q.x - f.x / den, q.y - f.y / den);
}

// Demons::demons_iteration (Demons.cpp:34-63) as a stand-alone pass on stored derivatives
template <class R>
__global__ void k_demons_correspondence(size_t n, const vec2_t<R> *__restrict__ gradI, const R *__restrict__ It, vec2_t<R> *__restrict__ corr, R sigma_isq,
                                        R sigma_xsq, unsigned *__restrict__ status) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const vec2_t<R> g = gradI[k];
        const R t = It[k];
        const R den = g.x * g.x + g.y * g.y + t * t * sigma_isq / sigma_xsq;
        if (den == 0) { atomicOr(status, OF2D_FLAG_DIVZERO); corr[k] = mk2<R>((R)0, (R)0); continue; }
        corr[k] = mk2<R>(g.x * t / den * (R)-1, g.y * t / den * (R)-1);
    }
}

// OpticalFlow::get_force (OpticalFlow.cpp:15-39) as a stand-alone pass
template <class R>
__global__ void k_lssd_force(size_t n, const vec2_t<R> *__restrict__ gradI, const R *__restrict__ It, const vec2_t<R> *__restrict__ u, vec2_t<R> *__restrict__ f) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        f[k] = lssd_force<R>(gradI[k], It[k], u[k]);
}

// ---------------------------------------------------------------------------------------------
// Demons force (DemonsThirions.cpp:18-27 + Demons.cpp:34-63): the warped image is never stored; each
// pixel re-evaluates the bilinear warp at its 4 stencil neighbours (served from L1).
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ R warped_at(const R *__restrict__ Imov, const vec2_t<R> *__restrict__ u, int nx, int ny, int i, int j) {
    const int idx = i + j * nx;
    return warp_pixel<R>(Imov, nx, ny, i, j, u[idx], Imov[idx]);
}

template <class R>
__global__ void __launch_bounds__(TX *TY) k_demons_force(int nx, int ny, const R *__restrict__ Iref, const R *__restrict__ Imov,
                                                         const vec2_t<R> *__restrict__ u, vec2_t<R> *__restrict__ corr, R sigma_isq, R sigma_xsq,
                                                         unsigned *__restrict__ status) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t off = (size_t)blockIdx.z * nx * ny;
    Iref += off; Imov += off; u += off; corr += off;
    const int idx = i + j * nx;
    const R c = warped_at<R>(Imov, u, nx, ny, i, j);
    R gx, gy;
    if (i == 0) gx = warped_at<R>(Imov, u, nx, ny, i + 1, j) - c;
    else if (i == nx - 1) gx = c - warped_at<R>(Imov, u, nx, ny, i - 1, j);
    else gx = (warped_at<R>(Imov, u, nx, ny, i + 1, j) - warped_at<R>(Imov, u, nx, ny, i - 1, j)) / (R)2.0f;
    if (j == 0) gy = warped_at<R>(Imov, u, nx, ny, i, j + 1) - c;
    else if (j == ny - 1) gy = c - warped_at<R>(Imov, u, nx, ny, i, j - 1);
    else gy = (warped_at<R>(Imov, u, nx, ny, i, j + 1) - warped_at<R>(Imov, u, nx, ny, i, j - 1)) / (R)2.0f;
    const R It = c - Iref[idx];
    const R den = gx * gx + gy * gy + It * It * sigma_isq / sigma_xsq;
    if (den == 0) { atomicOr(&status[blockIdx.z], OF2D_FLAG_DIVZERO); corr[idx] = mk2<R>((R)0, (R)0); return; }
    corr[idx] = mk2<R>(gx * It / den * (R)-1, gy * It / den * (R)-1);
}

// ---------------------------------------------------------------------------------------------
// Lexicographic SOR sweep (OpticalFlowElastic.cpp:21-55 == OpticalFlowFluid.cpp:7-41).
//
// The reference updates in place with x (i) as the outer and y (j) as the inner loop, so cell (i,j)
// sees NEW values at (i-1, j-1..j+1) and (i, j-1) and OLD values elsewhere.  Cells on a hyperplane
// t = 2i + j are independent, and every dependence points to a smaller t, so executing hyperplanes in
// order reproduces the sequential sweep exactly.
//
// Mapping: the interior columns are cut into bands of 32; one warp owns a band, lane l owns column
// I0 + l and at step t updates row j = t - 2l.  The band's rows live in a shared-memory ring that is
// updated in place (so "new to the west/south, old to the east/north" holds by construction); rows
// are prefetched D at a time through registers, the force is evaluated row-parallel when a row
// enters the ring, and finished rows are written back coalesced.  Neighbouring bands hand over
// through progress counters in global memory: a band may load row r only after its western
// neighbour has published row r (its last column is then final).  All bands of a launch are
// co-resident (cooperative launch), and tasks are dealt round-robin in increasing band order, so
// the wait chain always ends at a running band.
// ---------------------------------------------------------------------------------------------
constexpr int SOR_D = 8;                       // rows per prefetch group
constexpr int SOR_RING = 64 + 2 * SOR_D + 8;   // ring rows
constexpr int SOR_PITCH = 34;                  // 32 columns + west/east halo
constexpr int SOR_PUBLISH = 16;                // publish progress every this many rows

template <class R>
struct SorArgs {
    int nx, ny, batch, nbands;
    vec2_t<R> *x;                // field swept in place (elastic: motion; fluid: velocity)
    const vec2_t<R> *uforce;     // field the force is evaluated on (elastic: NULL = x itself; fluid: motion)
    const vec2_t<R> *gradI;
    const R *It;
    R c_keep, c_relax, mu, mupl; // (1-omega), omega/(-6 mu - 2 lambda), mu, mu+lambda
    unsigned *progress;          // [batch*nbands]
    unsigned base;               // epoch offset: a band has finished row r when progress >= base + r
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <class V> __device__ __forceinline__ V ld_l2(const V *p) { return __ldcg(p); }

template <class R>
__global__ void __launch_bounds__(32) k_sor_wavefront(SorArgs<R> A) {
    using V = vec2_t<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V *ring_x = reinterpret_cast<V *>(smem_raw);                 // [SOR_RING][SOR_PITCH]
    V *ring_b = ring_x + SOR_RING * SOR_PITCH;                   // [SOR_RING][32]
    const int lane = threadIdx.x;
    const int nx = A.nx, ny = A.ny;
    const int ntasks = A.batch * A.nbands;

    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int pair = task / A.nbands, band = task % A.nbands;
        const size_t off = (size_t)pair * nx * ny;
        V *x = A.x + off;
        const V *uf = A.uforce ? A.uforce + off : nullptr;
        const V *gradI = A.gradI + off;
        const R *It = A.It + off;
        const int I0 = 1 + 32 * band;                            // first interior column of the band
        const int ncols = min(32, (nx - 1) - I0);                // interior columns are 1 .. nx-2
        const bool active = lane < ncols;
        const int i = I0 + lane;
        const unsigned *west = band > 0 ? A.progress + task - 1 : nullptr;
        unsigned *mine = A.progress + task;
        unsigned west_seen = 0;
        const int T = (ny - 2) + 2 * (ncols - 1);                // last hyperplane of this band

        V q_x[SOR_D], q_h[SOR_D], q_g[SOR_D], q_u[SOR_D];
        R q_it[SOR_D];

        // rows [r0, r0+D) -> registers.  Lane l loads its own column; lanes 0/1 also the two halos.
        auto issue = [&](int r0) {
            if (r0 > ny - 1) return;
            if (west) {
                const int rmax = min(r0 + SOR_D - 1, ny - 2);    // boundary rows are never written
                if (rmax >= 1) {
                    const unsigned need = A.base + (unsigned)rmax;
                    while (west_seen < need) {
                        unsigned v = 0;
                        if (lane == 0) v = ld_acquire(west);
                        west_seen = __shfl_sync(0xffffffffu, v, 0);
                        if (west_seen < need) __nanosleep(64);
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < SOR_D; d++) {
                const int r = r0 + d;
                if (r > ny - 1) continue;
                const size_t row = (size_t)r * nx;
                if (i <= nx - 1) q_x[d] = ld_l2(&x[row + i]);
                if (lane == 0) q_h[d] = ld_l2(&x[row + I0 - 1]);
                if (lane == 1 && I0 + 32 <= nx - 1) q_h[d] = ld_l2(&x[row + I0 + 32]);
                if (active && r >= 1 && r <= ny - 2) {
                    q_g[d] = gradI[row + i];
                    q_it[d] = It[row + i];
                    if (uf) q_u[d] = uf[row + i];
                }
            }
        };
        // registers -> ring, force evaluated on the OLD value of the cell (OpticalFlowElastic.cpp:15)
        auto commit = [&](int r0) {
#pragma unroll
            for (int d = 0; d < SOR_D; d++) {
                const int r = r0 + d;
                if (r > ny - 1) continue;
                const int rr = r % SOR_RING;
                if (i <= nx - 1) ring_x[rr * SOR_PITCH + lane + 1] = q_x[d];
                if (lane == 0) ring_x[rr * SOR_PITCH] = q_h[d];
                if (lane == 1 && I0 + 32 <= nx - 1) ring_x[rr * SOR_PITCH + 33] = q_h[d];
                if (active && r >= 1 && r <= ny - 2)
                    ring_b[rr * 32 + lane] = lssd_force<R>(q_g[d], q_it[d], uf ? q_u[d] : q_x[d]);
            }
        };

        issue(0);
        commit(0);
        issue(SOR_D);
        __syncwarp();
        int published = 0;
        for (int g = 0; g * SOR_D + 1 <= T; g++) {
            commit((g + 1) * SOR_D);
            issue((g + 2) * SOR_D);
            __syncwarp();
#pragma unroll 1
            for (int s = 1; s <= SOR_D; s++) {
                const int t = g * SOR_D + s;
                if (t > T) break;
                const int j = t - 2 * lane;
                if (active && j >= 1 && j <= ny - 2) {
                    const int c = lane + 1;
                    const V *rm = ring_x + ((j - 1) % SOR_RING) * SOR_PITCH + c;
                    const V *r0 = ring_x + (j % SOR_RING) * SOR_PITCH + c;
                    const V *rp = ring_x + ((j + 1) % SOR_RING) * SOR_PITCH + c;
                    const V C = r0[0], E = r0[1], W = r0[-1];
                    const V N = rp[0], NE = rp[1], NW = rp[-1];
                    const V S = rm[0], SE = rm[1], SW = rm[-1];
                    const V b = ring_b[(j % SOR_RING) * 32 + lane];
                    V o;
                    o.x = A.c_keep * C.x + A.c_relax * (b.x - A.mu * (E.x + W.x + N.x + S.x) -
                                                        A.mupl * (E.x + W.x + (R)0.25f * (NE.y - NW.y - SE.y + SW.y)));
                    o.y = A.c_keep * C.y + A.c_relax * (b.y - A.mu * (E.y + W.y + N.y + S.y) -
                                                        A.mupl * (E.y + W.y + (R)0.25f * (NE.x - NW.x - SE.x + SW.x)));
                    ring_x[(j % SOR_RING) * SOR_PITCH + c] = o;
                }
                __syncwarp();
                // row finished by the last column of the band: write it back, coalesced
                const int rdone = t - 2 * (ncols - 1);
                if (rdone >= 1 && rdone <= ny - 2) {
                    if (active) x[(size_t)rdone * nx + i] = ring_x[(rdone % SOR_RING) * SOR_PITCH + lane + 1];
                    if (rdone - published >= SOR_PUBLISH || rdone == ny - 2) {
                        __threadfence();
                        __syncwarp();
                        if (lane == 0) st_release(mine, A.base + (unsigned)rdone);
                        published = rdone;
                    }
                }
            }
        }
        __syncwarp();
    }
}

// fluid increment (OpticalFlowFluid.cpp:60-90): R = v - du/dx * v.x - du/dy * v.y, plus the per-block
// maximum of Motion::maxabs's term
template <class R>
__global__ void __launch_bounds__(TX *TY) k_fluid_increment(int nx, int ny, const vec2_t<R> *__restrict__ u, const vec2_t<R> *__restrict__ vel,
                                                            vec2_t<R> *__restrict__ incr, double *__restrict__ partial_max) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    R m = (R)0;
    if (i < nx && j < ny) {
        const int idx = i + j * nx;
        const vec2_t<R> v = vel[idx];
        const vec2_t<R> dudx = partial_x_v<R>(u, idx, i, nx);
        const vec2_t<R> dudy = partial_y_v<R>(u, idx, j, nx, ny);
        const vec2_t<R> r = mk2<R>(v.x - dudx.x * v.x - dudy.x * v.y, v.y - dudx.y * v.x - dudy.y * v.y);
        incr[idx] = r;
        m = maxabs_term<R>(r);
    }
    m = block_extreme<R, true>(m);
    if (threadIdx.x == 0 && threadIdx.y == 0) partial_max[blockIdx.x + blockIdx.y * gridDim.x] = (double)m;
}
__global__ void k_finalize_max2(int nblocks, const double *__restrict__ partial, double *__restrict__ out) {
    double m = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) m = partial[k] > m ? partial[k] : m;
    m = block_extreme<double, true>(m);
    if (threadIdx.x == 0) out[0] = m;
}
// OpticalFlowFluid.cpp:97-121: u += R * dt
template <class R>
__global__ void k_integrate(size_t n, R dt, const vec2_t<R> *__restrict__ incr, vec2_t<R> *__restrict__ u) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const vec2_t<R> r = incr[k];
        vec2_t<R> v = u[k];
        v.x += r.x * dt; v.y += r.y * dt;
        u[k] = v;
    }
}

template <class R>
int poll_status(of2d_ctx *ctx, int batch, unsigned *h_status) {
    if (!h_status) return OF2D_SUCCESS;
    OF2D_CUDA_TRY(cudaMemcpyAsync(ctx->h_mailbox, ctx->d_status, sizeof(unsigned) * batch, cudaMemcpyDeviceToHost, ctx->stream));
    OF2D_CUDA_TRY(cudaMemsetAsync(ctx->d_status, 0, sizeof(unsigned) * batch, ctx->stream));
    OF2D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(h_status, ctx->h_mailbox, sizeof(unsigned) * batch);
    return OF2D_SUCCESS;
}

template <class R>
int diffusion_impl(of2d_ctx *ctx, int nx, int ny, int batch, const R *u, R *unew, const R *gradI, const R *It, R alpha, unsigned *h_status) {
    OF2D_REQUIRE(nx > 0 && ny > 0 && batch > 0 && batch <= kMaxBatchStatus && batch * sizeof(unsigned) <= 4096, "bad dimensions");
    OF2D_REQUIRE(u != unew, "diffusion step is out of place");
    k_diffusion_step<R><<<grid2d(nx, ny, batch), dim3(TX, TY), 0, ctx->stream>>>(nx, ny, (const vec2_t<R> *)u, (vec2_t<R> *)unew, (const vec2_t<R> *)gradI, It,
                                                                              alpha * alpha, ctx->d_status);
    OF2D_LAUNCH_CHECK(ctx);
    return poll_status<R>(ctx, batch, h_status);
}

template <class R>
int demons_force_impl(of2d_ctx *ctx, int nx, int ny, int batch, const R *Iref, const R *Imov, const R *u, R *corr, R sigma_i, R sigma_x, unsigned *h_status) {
    OF2D_REQUIRE(nx > 1 && ny > 1 && batch > 0 && batch * sizeof(unsigned) <= 4096, "bad dimensions");
    k_demons_force<R><<<grid2d(nx, ny, batch), dim3(TX, TY), 0, ctx->stream>>>(nx, ny, Iref, Imov, (const vec2_t<R> *)u, (vec2_t<R> *)corr, sigma_i * sigma_i,
                                                                            sigma_x * sigma_x, ctx->d_status);
    OF2D_LAUNCH_CHECK(ctx);
    return poll_status<R>(ctx, batch, h_status);
}

template <class R>
int sor_sweep_impl(of2d_ctx *ctx, int nx, int ny, int batch, R *x, const R *uforce, const R *gradI, const R *It, R mu, R lambda, R omega) {
    OF2D_REQUIRE(nx > 0 && ny > 0 && batch > 0, "bad dimensions");
    if (nx < 3 || ny < 3) return OF2D_SUCCESS;   // no interior cell: the reference's loops do not execute
    SorArgs<R> A;
    A.nx = nx; A.ny = ny; A.batch = batch;
    A.nbands = ceil_div(nx - 2, 32);
    A.x = (vec2_t<R> *)x;
    A.uforce = (const vec2_t<R> *)uforce;
    A.gradI = (const vec2_t<R> *)gradI;
    A.It = It;
    A.c_keep = (R)1.0f - omega;
    A.c_relax = omega / ((R)-6 * mu - (R)2 * lambda);
    A.mu = mu;
    A.mupl = mu + lambda;
    const size_t ntasks = (size_t)batch * A.nbands;
    if (ctx->progress_cap < ntasks) {
        if (ctx->d_progress) { OF2D_CUDA_TRY(cudaStreamSynchronize(ctx->stream)); OF2D_CUDA_TRY(cudaFree(ctx->d_progress)); }
        OF2D_CUDA_TRY(cudaMalloc(&ctx->d_progress, sizeof(unsigned) * ntasks));
        ctx->progress_cap = ntasks;
        ctx->progress_epoch = 0;
        OF2D_CUDA_TRY(cudaMemsetAsync(ctx->d_progress, 0, sizeof(unsigned) * ntasks, ctx->stream));
    }
    // Counters are monotone across sweeps: every launch gets a fresh window [base, base + ny + 2) strictly
    // above everything published before (whatever the previous grid size was), so no reset is needed
    // between launches -- only on wrap-around.
    const unsigned span = (unsigned)ny + 2u;
    if ((uint64_t)ctx->progress_epoch + 2ull * span >= 0x7fffffffull) {
        OF2D_CUDA_TRY(cudaMemsetAsync(ctx->d_progress, 0, sizeof(unsigned) * ctx->progress_cap, ctx->stream));
        ctx->progress_epoch = 0;
    }
    A.progress = ctx->d_progress;
    A.base = ctx->progress_epoch + 1;
    ctx->progress_epoch += span;

    const size_t smem = sizeof(vec2_t<R>) * (size_t)SOR_RING * (SOR_PITCH + 32);
    static bool configured[2][64] = {};
    const int slot = sizeof(R) == 8;
    if (!configured[slot][ctx->device & 63]) {
        OF2D_CUDA_TRY(cudaFuncSetAttribute(k_sor_wavefront<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[slot][ctx->device & 63] = true;
    }
    int per_sm = 0;
    OF2D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sor_wavefront<R>, 32, smem));
    OF2D_REQUIRE(per_sm > 0, "SOR kernel does not fit on an SM");
    const size_t resident = (size_t)per_sm * ctx->sm_count;
    const int grid = (int)(ntasks < resident ? ntasks : resident);
    void *args[] = {&A};
    OF2D_CUDA_TRY(cudaLaunchCooperativeKernel((void *)k_sor_wavefront<R>, dim3(grid), dim3(32), args, smem, ctx->stream));
    ctx->launches++;
    return OF2D_SUCCESS;
}

template <class R>
int fluid_step_impl(of2d_ctx *ctx, int nx, int ny, R *u, R *vel, R *incr, const R *gradI, const R *It, R mu, R lambda, R omega, R *h_maxabs, R *h_dt) {
    OF2D_REQUIRE(nx > 1 && ny > 1, "bad dimensions");
    int st = sor_sweep_impl<R>(ctx, nx, ny, 1, vel, u, gradI, It, mu, lambda, omega);
    if (st) return st;
    const dim3 g = grid2d(nx, ny, 1);
    const int nb = g.x * g.y;
    OF2D_REQUIRE(nb <= kMaxPartialBlocks * 4, "image too large for the reduction scratch");
    k_fluid_increment<R><<<g, dim3(TX, TY), 0, ctx->stream>>>(nx, ny, (const vec2_t<R> *)u, (const vec2_t<R> *)vel, (vec2_t<R> *)incr, ctx->d_partials);
    OF2D_LAUNCH_CHECK(ctx);
    k_finalize_max2<<<1, 256, 0, ctx->stream>>>(nb, ctx->d_partials, (double *)ctx->d_mailbox);
    OF2D_LAUNCH_CHECK(ctx);
    OF2D_CUDA_TRY(cudaMemcpyAsync(ctx->h_mailbox, ctx->d_mailbox, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    OF2D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const double msq = *(const double *)ctx->h_mailbox;
    const R maxabs = sizeof(R) == 4 ? (R)sqrtf((float)msq) : (R)sqrt(msq);   // Motion.cpp:57
    const R dumax = (R)0.65f;                                                // OpticalFlowFluid.h:32
    const R dt = dumax / maxabs;                                             // OpticalFlowFluid.cpp:93
    if (h_maxabs) *h_maxabs = maxabs;
    if (h_dt) *h_dt = dt;
    if (dt >= (R)65.0f) return OF2D_SUCCESS;                                 // OpticalFlowFluid.cpp:135-137
    const size_t n = (size_t)nx * ny;
    const size_t want = (n + 255) / 256, cap = (size_t)ctx->sm_count * 8;
    k_integrate<R><<<(int)(want < cap ? want : cap), 256, 0, ctx->stream>>>(n, dt, (const vec2_t<R> *)incr, (vec2_t<R> *)u);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}

}  // namespace

extern "C" {

int of2d_poll_status(of2d_ctx *ctx, int batch, unsigned *h_status) {
    OF2D_REQUIRE(batch > 0 && batch * sizeof(unsigned) <= 4096 && h_status, "bad arguments");
    return poll_status<float>(ctx, batch, h_status);
}

int of2d_demons_correspondence_f32(of2d_ctx *ctx, size_t n, const float *g, const float *It, float *c, float si, float sx) {
    const size_t want = (n + 255) / 256, cap = (size_t)ctx->sm_count * 8;
    k_demons_correspondence<float><<<(int)(want < cap ? want : cap), 256, 0, ctx->stream>>>(n, (const float2 *)g, It, (float2 *)c, si * si, sx * sx, ctx->d_status);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}
int of2d_demons_correspondence_f64(of2d_ctx *ctx, size_t n, const double *g, const double *It, double *c, double si, double sx) {
    const size_t want = (n + 255) / 256, cap = (size_t)ctx->sm_count * 8;
    k_demons_correspondence<double><<<(int)(want < cap ? want : cap), 256, 0, ctx->stream>>>(n, (const double2 *)g, It, (double2 *)c, si * si, sx * sx, ctx->d_status);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}
int of2d_lssd_force_f32(of2d_ctx *ctx, int nx, int ny, int batch, const float *g, const float *It, const float *u, float *f) {
    const size_t n = (size_t)nx * ny * batch, want = (n + 255) / 256, cap = (size_t)ctx->sm_count * 8;
    k_lssd_force<float><<<(int)(want < cap ? want : cap), 256, 0, ctx->stream>>>(n, (const float2 *)g, It, (const float2 *)u, (float2 *)f);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}
int of2d_lssd_force_f64(of2d_ctx *ctx, int nx, int ny, int batch, const double *g, const double *It, const double *u, double *f) {
    const size_t n = (size_t)nx * ny * batch, want = (n + 255) / 256, cap = (size_t)ctx->sm_count * 8;
    k_lssd_force<double><<<(int)(want < cap ? want : cap), 256, 0, ctx->stream>>>(n, (const double2 *)g, It, (const double2 *)u, (double2 *)f);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}
int of2d_diffusion_step_f32(of2d_ctx *ctx, int nx, int ny, int batch, const float *u, float *unew, const float *g, const float *It, float alpha, unsigned *hs) { return diffusion_impl<float>(ctx, nx, ny, batch, u, unew, g, It, alpha, hs); }
int of2d_diffusion_step_f64(of2d_ctx *ctx, int nx, int ny, int batch, const double *u, double *unew, const double *g, const double *It, double alpha, unsigned *hs) { return diffusion_impl<double>(ctx, nx, ny, batch, u, unew, g, It, alpha, hs); }
int of2d_elastic_step_f32(of2d_ctx *ctx, int nx, int ny, int batch, float *u, const float *g, const float *It, float mu, float lambda, float omega) { return sor_sweep_impl<float>(ctx, nx, ny, batch, u, nullptr, g, It, mu, lambda, omega); }
int of2d_elastic_step_f64(of2d_ctx *ctx, int nx, int ny, int batch, double *u, const double *g, const double *It, double mu, double lambda, double omega) { return sor_sweep_impl<double>(ctx, nx, ny, batch, u, nullptr, g, It, mu, lambda, omega); }
int of2d_fluid_step_f32(of2d_ctx *ctx, int nx, int ny, float *u, float *v, float *incr, const float *g, const float *It, float mu, float lambda, float omega, float *hm, float *hd) { return fluid_step_impl<float>(ctx, nx, ny, u, v, incr, g, It, mu, lambda, omega, hm, hd); }
int of2d_fluid_step_f64(of2d_ctx *ctx, int nx, int ny, double *u, double *v, double *incr, const double *g, const double *It, double mu, double lambda, double omega, double *hm, double *hd) { return fluid_step_impl<double>(ctx, nx, ny, u, v, incr, g, It, mu, lambda, omega, hm, hd); }
int of2d_demons_force_f32(of2d_ctx *ctx, int nx, int ny, int batch, const float *r, const float *m, const float *u, float *c, float si, float sx, unsigned *hs) { return demons_force_impl<float>(ctx, nx, ny, batch, r, m, u, c, si, sx, hs); }
int of2d_demons_force_f64(of2d_ctx *ctx, int nx, int ny, int batch, const double *r, const double *m, const double *u, double *c, double si, double sx, unsigned *hs) { return demons_force_impl<double>(ctx, nx, ny, batch, r, m, u, c, si, sx, hs); }

}  // extern "C"
