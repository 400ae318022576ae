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
#include <math.h>
#include <string.h>

#include <new>
#include <vector>

#include "device_math.cuh"
#include "engine_ctl.cuh"
#include "engine_internal.cuh"

namespace {

constexpr int TX = 32, TY = 8, PY = 4;   // CTA = 32 x 8 threads, PY rows per thread: 32 x 32 pixel tile
constexpr int TILE = 32;

__device__ __forceinline__ int ld_int(const int *p) { return __ldcg(p); }

enum Buf { B_C0 = 0, B_C1 = 1, B_EST_CUR = 2, B_EST_NEXT = 3, B_CRES = 4, B_CTMP = 5, B_LVL_CUR = 6, B_LVL_NEXT = 7, B_ESTN = 8, B_EXT = 9 };
enum Gate { G_NONE = 0, G_ACTIVE = 1, G_REGRID = 2 };

// pointers of one pair family, by value in every kernel
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
    vec2_t<R> *estN;
    vec2_t<R> *ext;                // caller-provided field (set per launch)
};

template <class R>
__device__ __forceinline__ vec2_t<R> *pick(const EngK<R> &K, int which, const PairCtl *c, int pair, bool transposed = false) {
    const size_t off = (size_t)pair * (transposed ? K.nT : K.n);
    switch (which) {
        case B_C0: return K.c[0] + off;
        case B_C1: return K.c[1] + off;
        case B_EST_CUR: return K.est[ld_int(&c->sel)] + off;
        case B_EST_NEXT: return K.est[ld_int(&c->sel) ^ 1] + off;
        case B_CRES: return K.c[(ld_int(&c->nsquares) & 1) ? 0 : 1] + off;
        case B_CTMP: return K.c[(ld_int(&c->nsquares) & 1) ? 1 : 0] + off;
        case B_LVL_CUR: return K.lvl[ld_int(&c->msel)] + off;
        case B_LVL_NEXT: return K.lvl[ld_int(&c->msel) ^ 1] + off;
        case B_ESTN: return K.estN + off;
        default: return K.ext + off;
    }
}

__device__ __forceinline__ bool gate_open(const PairCtl *c, int gate) {
    if (gate == G_ACTIVE) return ld_int(&c->active) != 0;
    if (gate == G_REGRID) return ld_int(&c->regrid) != 0;
    return true;
}

// ---------------------------------------------------------------------------------------------
// control
// ---------------------------------------------------------------------------------------------
__global__ void k_ctl_begin(PairCtl *ctl, int batch, int niter, int *n_active) {
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
    for (int k = 0; k < 4; k++) c->ticket[k] = 0u;
    c->err = 0.0;
}

__global__ void k_regrid_commit(PairCtl *ctl, int batch) {
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
template <class R>
__global__ void __launch_bounds__(TX *TY) k_hs_iter(EngK<R> K, const vec2_t<R> *__restrict__ gradI_all, const R *__restrict__ It_all, R alphasq) {
    const int pair = blockIdx.z;
    PairCtl *c = K.ctl + pair;
    if (!ld_int(&c->active)) return;
    const int nx = K.nx, ny = K.ny;
    const vec2_t<R> *__restrict__ u = pick(K, B_EST_CUR, c, pair);
    vec2_t<R> *__restrict__ un = pick(K, B_EST_NEXT, c, pair);
    const vec2_t<R> *__restrict__ gradI = gradI_all + (size_t)pair * K.n;
    const R *__restrict__ It = It_all + (size_t)pair * K.n;
    const int i = blockIdx.x * TILE + threadIdx.x;
    double sd = 0.0, sp = 0.0;
    if (i < nx) {
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int j = blockIdx.y * TILE + threadIdx.y + p * TY;
            if (j >= ny) break;
            const int idx = i + j * nx;
            vec2_t<R> q;
            if (i == 0 || i == nx - 1 || j == 0 || j == ny - 1) {
                q = mk2<R>((R)0.0f, (R)0.0f);
            } else {   // gradients.h:78
                const vec2_t<R> a = u[idx - 1], b = u[idx + 1], cc = u[idx - nx], d = u[idx + nx];
                q = mk2<R>((((a.x + b.x) + cc.x) + d.x) / (R)4.0f, (((a.y + b.y) + cc.y) + d.y) / (R)4.0f);
            }
            const vec2_t<R> dI = gradI[idx];
            const vec2_t<R> f = lssd_force<R>(dI, It[idx], q);
            const R den = alphasq + dI.x * dI.x + dI.y * dI.y;
            vec2_t<R> o;
            if (den == 0) { atomicOr(&c->flags, OF2D_FLAG_DIVZERO); o = q; }
            else o = mk2<R>(q.x - f.x / den, q.y - f.y / den);
            const vec2_t<R> old = u[idx];
            un[idx] = o;
            sd += vec_norm_d<R>(mk2<R>(o.x - old.x, o.y - old.y));
            sp += vec_norm_d<R>(old);
        }
    }
    block_sum2(sd, sp);
    const double vals[2] = {sd, sp};
    const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.x + blockIdx.y * gridDim.x;
    double *part = K.partials + (size_t)pair * K.pstride;
    if (publish_partials<2>(vals, part, &c->ticket[0], nblocks, bid)) {
        double out[2];
        reduce_partials<2>(part, nblocks, out, 0u, 0u);
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            c->sel ^= 1;
            finalize_logger<R>(c, K.tr, pair, out[0], out[1], (unsigned)K.n, K.n_active);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// generic gated primitives of the loops
// ---------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(TX *TY) k_e_warp(EngK<R> K, int gate, const R *__restrict__ src_all, int u_buf, R *__restrict__ dst_all) {
    const int pair = blockIdx.z;
    const PairCtl *c = K.ctl + pair;
    if (!gate_open(c, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const R *src = src_all + (size_t)pair * K.n;
    R *dst = dst_all + (size_t)pair * K.n;
    const vec2_t<R> *u = pick(K, u_buf, c, pair);
    const int i = blockIdx.x * TILE + threadIdx.x;
    if (i >= nx) return;
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int j = blockIdx.y * TILE + threadIdx.y + p * TY;
        if (j >= ny) break;
        const int idx = i + j * nx;
        dst[idx] = warp_pixel<R>(src, nx, ny, i, j, u[idx], src[idx]);
    }
}

// out = v + u o (id + v)   (Motion::accumulate, Motion.cpp:113-178); mode 1: out = u + v (Field::operator+=)
template <class R>
__global__ void __launch_bounds__(TX *TY) k_e_compose(EngK<R> K, int gate, int u_buf, int v_buf, int out_buf, int add_only) {
    const int pair = blockIdx.z;
    const PairCtl *c = K.ctl + pair;
    if (!gate_open(c, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const vec2_t<R> *__restrict__ u = pick(K, u_buf, c, pair);
    const vec2_t<R> *__restrict__ v = pick(K, v_buf, c, pair);
    vec2_t<R> *__restrict__ out = pick(K, out_buf, c, pair);
    const int i = blockIdx.x * TILE + threadIdx.x;
    if (i >= nx) return;
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int j = blockIdx.y * TILE + threadIdx.y + p * TY;
        if (j >= ny) break;
        const int idx = i + j * nx;
        const vec2_t<R> vv = v[idx], uu = u[idx];
        out[idx] = add_only ? mk2<R>(uu.x + vv.x, uu.y + vv.y) : compose_pixel<R>(u, nx, ny, i, j, vv, uu);
    }
}

// one squaring of Motion::exp (Motion.cpp:262-274): dst = w + w o (id + w), w = scale * src (scale only at s == 0)
template <class R>
__global__ void __launch_bounds__(TX *TY) k_e_square(EngK<R> K, int s) {
    const int pair = blockIdx.z;
    const PairCtl *c = K.ctl + pair;
    if (!ld_int(&c->active) || s >= ld_int(&c->nsquares)) return;
    const int nx = K.nx, ny = K.ny;
    const size_t off = (size_t)pair * K.n;
    const vec2_t<R> *__restrict__ src = K.c[(s & 1) ? 0 : 1] + off;
    vec2_t<R> *__restrict__ dst = K.c[(s & 1) ? 1 : 0] + off;
    const R sc = s == 0 ? (R)__ldcg(&c->scale) : (R)1;
    const int i = blockIdx.x * TILE + threadIdx.x;
    if (i >= nx) return;
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int j = blockIdx.y * TILE + threadIdx.y + p * TY;
        if (j >= ny) break;
        const int idx = i + j * nx;
        vec2_t<R> v = src[idx];
        v.x *= sc; v.y *= sc;
        // compose_pixel on the scaled field: taps are scaled on the fly (a power of two: exact)
        const Bilin<R> b = bilin_setup<R>(i, j, v.x, v.y, nx, ny);
        vec2_t<R> o = v;   // out of bounds: keeps the (scaled) value
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
        dst[idx] = o;
    }
}

// Demons force: warp + derivatives + demons_iteration (DemonsThirions.cpp:18-27, Demons.cpp:34-63)
template <class R>
__device__ __forceinline__ R warped_at(const R *__restrict__ Imov, const vec2_t<R> *__restrict__ u, int nx, int ny, int i, int j) {
    const int idx = i + j * nx;
    return warp_pixel<R>(Imov, nx, ny, i, j, u[idx], Imov[idx]);
}

template <class R>
__global__ void __launch_bounds__(TX *TY) k_e_demons_force(EngK<R> K, const R *__restrict__ Iref_all, const R *__restrict__ Imov_all, R sigma_isq, R sigma_xsq) {
    const int pair = blockIdx.z;
    PairCtl *c = K.ctl + pair;
    if (!ld_int(&c->active)) return;
    const int nx = K.nx, ny = K.ny;
    const R *__restrict__ Iref = Iref_all + (size_t)pair * K.n;
    const R *__restrict__ Imov = Imov_all + (size_t)pair * K.n;
    const vec2_t<R> *__restrict__ u = pick(K, B_EST_CUR, c, pair);
    vec2_t<R> *__restrict__ corr = K.c[0] + (size_t)pair * K.n;
    const int i = blockIdx.x * TILE + threadIdx.x;
    if (i >= nx) return;
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int j = blockIdx.y * TILE + threadIdx.y + p * TY;
        if (j >= ny) break;
        const int idx = i + j * nx;
        const R cc = warped_at<R>(Imov, u, nx, ny, i, j);
        R gx, gy;
        if (i == 0) gx = warped_at<R>(Imov, u, nx, ny, i + 1, j) - cc;
        else if (i == nx - 1) gx = cc - warped_at<R>(Imov, u, nx, ny, i - 1, j);
        else gx = (warped_at<R>(Imov, u, nx, ny, i + 1, j) - warped_at<R>(Imov, u, nx, ny, i - 1, j)) / (R)2.0f;
        if (j == 0) gy = warped_at<R>(Imov, u, nx, ny, i, j + 1) - cc;
        else if (j == ny - 1) gy = cc - warped_at<R>(Imov, u, nx, ny, i, j - 1);
        else gy = (warped_at<R>(Imov, u, nx, ny, i, j + 1) - warped_at<R>(Imov, u, nx, ny, i, j - 1)) / (R)2.0f;
        const R It = cc - Iref[idx];
        const R den = gx * gx + gy * gy + It * It * sigma_isq / sigma_xsq;
        if (den == 0) { atomicOr(&c->flags, OF2D_FLAG_DIVZERO); corr[idx] = mk2<R>((R)0, (R)0); continue; }
        corr[idx] = mk2<R>(gx * It / den * (R)-1, gy * It / den * (R)-1);
    }
}

// ---------------------------------------------------------------------------------------------
// convolution (Field.tpp:210-269) on a shared-memory tile.  The bounds test of the reference is on
// the LINEAR index, so a tile is simply rows of the flattened array: element (row r, column c) of the
// halo is in[r*nx + c] whenever that flat index is inside [0, n) -- columns outside [0, nx) land in
// the neighbouring row exactly as in the reference.
//   EPI 0: none   EPI 1: Logger epilogue (result is the next estimate)   EPI 2: maxabs epilogue
// ---------------------------------------------------------------------------------------------
constexpr int kConvMaxHalf = 7;      // kernel widths up to 15
template <class R>
struct ConvW {
    const R *taps;         // (real) weights, [kw*kh] column-major as Kernel::get_kernel()
    const double *taps_d;  // double weights
    double full_weight;    // sum over the visiting order
    int kw, kh, cx, cy;
};

template <class R, int EPI>
__global__ void __launch_bounds__(TX *TY) k_e_conv(EngK<R> K, int src_buf, int dst_buf, ConvW<R> W, int nsq_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int pair = blockIdx.z;
    PairCtl *c = K.ctl + pair;
    if (!ld_int(&c->active)) return;
    const int nx = K.nx, ny = K.ny;
    const long n = (long)K.n;
    const vec2_t<R> *__restrict__ in = pick(K, src_buf, c, pair);
    vec2_t<R> *__restrict__ out = pick(K, dst_buf, c, pair);
    const int cx = W.cx, cy = W.cy, kw = W.kw;
    const int SW = TILE + 2 * cx, SH = TILE + 2 * cy;
    vec2_t<R> *tile = reinterpret_cast<vec2_t<R> *>(smem_raw);        // [SH][SW]
    R *wts = reinterpret_cast<R *>(tile + SH * SW);                    // [kw*kh]
    const int tid = threadIdx.x + threadIdx.y * TX;
    const int i0 = blockIdx.x * TILE, j0 = blockIdx.y * TILE;
    for (int e = tid; e < SH * SW; e += TX * TY) {
        const int r = e / SW, cc = e - r * SW;
        const long lin = (long)(j0 + r - cy) * nx + (i0 + cc - cx);
        vec2_t<R> v = mk2<R>((R)0, (R)0);
        if (lin >= 0 && lin < n) v = in[lin];
        tile[e] = v;
    }
    for (int e = tid; e < kw * W.kh; e += TX * TY) wts[e] = W.taps[e];
    __syncthreads();
    const int i = i0 + threadIdx.x;
    double sd = 0.0, sp = 0.0;
    R mx = (R)0;
    if (i < nx) {
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int jl = threadIdx.y + p * TY, j = j0 + jl;
            if (j >= ny) break;
            const long idx = i + (long)j * nx;
            const bool interior = (idx - cx - (long)cy * nx >= 0) && (idx + cx + (long)cy * nx < n);
            R ax = (R)0, ay = (R)0;
            double weight = 0.0;
            // visiting order of the reference: ii outer, jj inner (Field.tpp:242-243)
            for (int ii = -cx; ii <= cx; ii++) {
                for (int jj = -cy; jj <= cy; jj++) {
                    const int ik = (ii + cx) + (jj + cy) * kw;
                    if (!interior) {
                        const long lin = idx + ii + (long)jj * nx;
                        if (lin < 0 || lin >= n) continue;
                        weight += W.taps_d[ik];
                    }
                    const vec2_t<R> f = tile[(jl + jj + cy) * SW + (threadIdx.x + ii + cx)];
                    const R t = wts[ik];
                    ax = r_fma(f.x, t, ax);
                    ay = r_fma(f.y, t, ay);
                }
            }
            if (interior) weight = W.full_weight;
            vec2_t<R> o;
            if (weight != 0) { const R w = (R)weight; o = mk2<R>(ax / w, ay / w); }
            else o = tile[(jl + cy) * SW + threadIdx.x + cx];
            out[idx] = o;
            if (EPI == 1) {
                const vec2_t<R> old = pick(K, B_EST_CUR, c, pair)[idx];
                sd += vec_norm_d<R>(mk2<R>(o.x - old.x, o.y - old.y));
                sp += vec_norm_d<R>(old);
            } else if (EPI == 2) {
                const R s = maxabs_term<R>(o);
                mx = mx < s ? s : mx;
            }
        }
    }
    if (EPI == 0) return;
    const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.x + blockIdx.y * gridDim.x;
    double *part = K.partials + (size_t)pair * K.pstride;
    if (EPI == 1) {
        block_sum2(sd, sp);
        const double vals[2] = {sd, sp};
        if (publish_partials<2>(vals, part, &c->ticket[0], nblocks, bid)) {
            double o2[2];
            reduce_partials<2>(part, nblocks, o2, 0u, 0u);
            if (tid == 0) {
                c->sel ^= 1;
                finalize_logger<R>(c, K.tr, pair, o2[0], o2[1], (unsigned)K.n, K.n_active);
            }
        }
    } else {
        mx = block_extreme<R, true>(mx);
        const double vals[1] = {(double)mx};
        if (publish_partials<1>(vals, part, &c->ticket[1], nblocks, bid)) {
            double o1[1];
            reduce_partials<1>(part, nblocks, o1, 1u, 0u);
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
    __shared__ vec2_t<R> sg[TILE][TILE + 1];
    __shared__ R st[TILE][TILE + 1];
    const int pair = blockIdx.z;
    const PairCtl *c = K.ctl + pair;
    if (!gate_open(c, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const R *__restrict__ Iref = Iref_all + (size_t)pair * K.n;
    const R *__restrict__ Imov = Imov_all + (size_t)pair * K.n;
    const int i = blockIdx.x * TILE + threadIdx.x;
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int jl = threadIdx.y + p * TY, j = blockIdx.y * TILE + jl;
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
    if (!transposed) return;
    __syncthreads();
    const int jt = blockIdx.y * TILE + threadIdx.x;   // fast thread index runs along j now
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int il = threadIdx.y + p * TY, it = blockIdx.x * TILE + il;
        if (it < nx && jt < ny) {
            const size_t o = (size_t)pair * K.nT + (size_t)it * K.P + jt;
            gradI_all[o] = sg[threadIdx.x][il];
            It_all[o] = st[threadIdx.x][il];
        }
    }
}

// transposed (i*P + j) -> normal (i + j*nx) copy of a vec2 field
template <class R>
__global__ void __launch_bounds__(TX *TY) k_e_untranspose(EngK<R> K, int gate, int src_buf, int dst_buf) {
    __shared__ vec2_t<R> s[TILE][TILE + 1];
    const int pair = blockIdx.z;
    const PairCtl *c = K.ctl + pair;
    if (!gate_open(c, gate)) return;
    const int nx = K.nx, ny = K.ny;
    const vec2_t<R> *__restrict__ src = pick(K, src_buf, c, pair, true);
    vec2_t<R> *__restrict__ dst = pick(K, dst_buf, c, pair, false);
    const int jt = blockIdx.y * TILE + threadIdx.x;
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int il = threadIdx.y + p * TY, it = blockIdx.x * TILE + il;
        if (it < nx && jt < ny) s[il][threadIdx.x] = src[(size_t)it * K.P + jt];
    }
    __syncthreads();
    const int i = blockIdx.x * TILE + threadIdx.x;
#pragma unroll
    for (int p = 0; p < PY; p++) {
        const int jl = threadIdx.y + p * TY, j = blockIdx.y * TILE + jl;
        if (i < nx && j < ny) dst[(size_t)i + (size_t)j * nx] = s[threadIdx.x][jl];
    }
}

template <class R>
__global__ void k_e_zero(EngK<R> K, int gate, int buf, int transposed) {
    const int pair = blockIdx.y;
    const PairCtl *c = K.ctl + pair;
    if (!gate_open(c, gate)) return;
    vec2_t<R> *p = pick(K, buf, c, pair, transposed != 0);
    const size_t cnt = transposed ? K.nT : K.n;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < cnt; k += (size_t)gridDim.x * blockDim.x) p[k] = mk2<R>((R)0, (R)0);
}

// ---------------------------------------------------------------------------------------------
// Fluid in the transposed layout (element (i,j) at i*P + j; threadIdx.x runs along j)
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ vec2_t<R> tdx(const vec2_t<R> *__restrict__ f, size_t o, int i, int nx, int P) {   // d/dx, gradients.h:9-19
    if (i == 0) { const vec2_t<R> a = f[o + P], b = f[o]; return mk2<R>(a.x - b.x, a.y - b.y); }
    if (i == nx - 1) { const vec2_t<R> a = f[o], b = f[o - P]; return mk2<R>(a.x - b.x, a.y - b.y); }
    const vec2_t<R> a = f[o + P], b = f[o - P];
    return mk2<R>((a.x - b.x) / (R)2.0f, (a.y - b.y) / (R)2.0f);
}
template <class R>
__device__ __forceinline__ vec2_t<R> tdy(const vec2_t<R> *__restrict__ f, size_t o, int j, int ny) {          // d/dy, gradients.h:22-32
    if (j == 0) { const vec2_t<R> a = f[o + 1], b = f[o]; return mk2<R>(a.x - b.x, a.y - b.y); }
    if (j == ny - 1) { const vec2_t<R> a = f[o], b = f[o - 1]; return mk2<R>(a.x - b.x, a.y - b.y); }
    const vec2_t<R> a = f[o + 1], b = f[o - 1];
    return mk2<R>((a.x - b.x) / (R)2.0f, (a.y - b.y) / (R)2.0f);
}

// increment R = v - du/dx v.x - du/dy v.y (OpticalFlowFluid.cpp:60-90) + time step (:92-95, :135-137)
template <class R>
__global__ void __launch_bounds__(TX *TY) k_fl_increment(EngK<R> K, vec2_t<R> *const vel0, vec2_t<R> *const vel1, vec2_t<R> *__restrict__ incr_all) {
    const int pair = blockIdx.z;
    PairCtl *c = K.ctl + pair;
    if (!ld_int(&c->active)) return;
    const int nx = K.nx, ny = K.ny, P = K.P;
    const vec2_t<R> *__restrict__ u = pick(K, B_EST_CUR, c, pair, true);
    const vec2_t<R> *__restrict__ vel = (ld_int(&c->vsel) ? vel0 : vel1) + (size_t)pair * K.nT;   // the buffer the sweep just wrote
    vec2_t<R> *__restrict__ incr = incr_all + (size_t)pair * K.nT;
    const int j = blockIdx.x * TILE + threadIdx.x;
    R m = (R)0;
    if (j < ny) {
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int i = blockIdx.y * TILE + threadIdx.y + p * TY;
            if (i >= nx) break;
            const size_t o = (size_t)i * P + j;
            const vec2_t<R> v = vel[o];
            const vec2_t<R> dudx = tdx<R>(u, o, i, nx, P);
            const vec2_t<R> dudy = tdy<R>(u, o, j, ny);
            const vec2_t<R> r = mk2<R>(v.x - dudx.x * v.x - dudy.x * v.y, v.y - dudx.y * v.x - dudy.y * v.y);
            incr[o] = r;
            const R s = maxabs_term<R>(r);
            m = m < s ? s : m;
        }
    }
    m = block_extreme<R, true>(m);
    const double vals[1] = {(double)m};
    const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.x + blockIdx.y * gridDim.x;
    double *part = K.partials + (size_t)pair * K.pstride;
    if (publish_partials<1>(vals, part, &c->ticket[1], nblocks, bid)) {
        double o1[1];
        reduce_partials<1>(part, nblocks, o1, 1u, 0u);
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            const R maxabs = sizeof(R) == 4 ? (R)sqrtf((float)o1[0]) : (R)sqrt(o1[0]);   // Motion.cpp:57
            const R dt = (R)0.65f / maxabs;                                              // OpticalFlowFluid.h:32, .cpp:93
            c->maxabs = (double)maxabs;
            c->dt = (double)dt;
            c->skip = dt >= (R)65.0f;                                                    // .cpp:135-137
            c->vsel ^= 1;
            const int it = c->iter;
            if (it < K.tr.cap) { K.tr.maxabs[(size_t)pair * K.tr.cap + it] = (double)maxabs; K.tr.dt[(size_t)pair * K.tr.cap + it] = (double)dt; }
        }
    }
}

// integrate u += dt R (OpticalFlowFluid.cpp:97-121) + Logger + Jacobian minimum of the new field
// (Image.cpp:189-218, :96-104) + break / regrid decisions (ImageRegistrationFluid.cpp:99-124)
template <class R>
__global__ void __launch_bounds__(TX *TY) k_fl_integrate(EngK<R> K, const vec2_t<R> *__restrict__ incr_all) {
    const int pair = blockIdx.z;
    PairCtl *c = K.ctl + pair;
    if (!ld_int(&c->active)) return;
    const int nx = K.nx, ny = K.ny, P = K.P;
    const vec2_t<R> *__restrict__ u = pick(K, B_EST_CUR, c, pair, true);
    vec2_t<R> *un = pick(K, B_EST_NEXT, c, pair, true);
    const vec2_t<R> *__restrict__ incr = incr_all + (size_t)pair * K.nT;
    const bool skip = ld_int(&c->skip) != 0;
    const bool prev_other = ld_int(&c->prev_other) != 0;
    const R dt = (R)__ldcg(&c->dt);
    auto unew_at = [&](size_t o) -> vec2_t<R> {
        vec2_t<R> v = u[o];
        if (!skip) { const vec2_t<R> r = incr[o]; v.x += r.x * dt; v.y += r.y * dt; }
        return v;
    };
    const int j = blockIdx.x * TILE + threadIdx.x;
    double sd = 0.0, sp = 0.0;
    R mj = (R)INFINITY;
    if (j < ny) {
#pragma unroll
        for (int p = 0; p < PY; p++) {
            const int i = blockIdx.y * TILE + threadIdx.y + p * TY;
            if (i >= nx) break;
            const size_t o = (size_t)i * P + j;
            const vec2_t<R> nv = unew_at(o);
            const vec2_t<R> prev = prev_other ? un[o] : u[o];   // after a regrid Logger's prev is the pre-reset estimate
            sd += vec_norm_d<R>(mk2<R>(nv.x - prev.x, nv.y - prev.y));
            sp += vec_norm_d<R>(prev);
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
    }
    block_sum2(sd, sp);
    mj = block_extreme<R, false>(mj);
    const double vals[3] = {sd, sp, (double)mj};
    const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.x + blockIdx.y * gridDim.x;
    double *part = K.partials + (size_t)pair * K.pstride;
    if (publish_partials<3>(vals, part, &c->ticket[0], nblocks, bid)) {
        double o3[3];
        reduce_partials<3>(part, nblocks, o3, 0u, 4u);
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

}  // namespace

#include "sor_tile.cuh"

// =================================================================================================
// host side
// =================================================================================================
struct of2d_engine {
    of2d_ctx *ctx;
    of2d_engine_desc d;
    bool dbl, transposed;
    int P;
    size_t n, nT, elem;          // elem = sizeof(real)
    // device buffers (real-typed; vec2 fields have 2*elem per element)
    void *aux, *gradI, *It, *est[2], *c[2], *lvl[2], *estN, *vel[2], *incr;
    PairCtl *d_ctl;
    int *d_nactive;
    double *d_partials;
    size_t pstride;
    TraceDev tr;
    void *d_taps[2];             // [fluid kernel, diffusion kernel]: doubles followed by reals
    double full_weight[2];
    int nsq_cap;
    of2d_curvature_plan *plan;
    SorPlan sor;
    // host
    int *h_snap;                 // pinned [2]
    cudaEvent_t ev[2];
    std::vector<PairCtl> h_ctl;
    uint64_t iterations_enqueued;
    const void *cur_Imov;
    int last_niter;
};

namespace {

template <class R>
EngK<R> make_k(of2d_engine *E) {
    EngK<R> K;
    K.nx = E->d.dimx; K.ny = E->d.dimy; K.batch = E->d.batch; K.P = E->P;
    K.n = E->n; K.nT = E->nT;
    K.ctl = E->d_ctl; K.n_active = E->d_nactive; K.partials = E->d_partials; K.pstride = E->pstride; K.tr = E->tr;
    for (int k = 0; k < 2; k++) { K.est[k] = (vec2_t<R> *)E->est[k]; K.c[k] = (vec2_t<R> *)E->c[k]; K.lvl[k] = (vec2_t<R> *)E->lvl[k]; }
    K.estN = (vec2_t<R> *)E->estN;
    K.ext = nullptr;
    return K;
}

inline dim3 grid_tiles(int nx, int ny, int batch) { return dim3(ceil_div(nx, TILE), ceil_div(ny, TILE), batch); }

template <class R>
ConvW<R> conv_weights(of2d_engine *E, int which) {
    ConvW<R> W;
    const int kw = E->d.kernel_w;
    W.kw = kw; W.kh = kw; W.cx = (kw - 1) / 2; W.cy = (kw - 1) / 2;
    W.taps_d = (const double *)E->d_taps[which];
    W.taps = (const R *)((const char *)E->d_taps[which] + sizeof(double) * kw * kw);
    W.full_weight = E->full_weight[which];
    return W;
}

template <class R, int EPI>
int launch_conv(of2d_engine *E, const EngK<R> &K, int src, int dst, int which) {
    const ConvW<R> W = conv_weights<R>(E, which);
    const size_t smem = sizeof(vec2_t<R>) * (size_t)(TILE + 2 * W.cx) * (TILE + 2 * W.cy) + sizeof(R) * W.kw * W.kh;
    k_e_conv<R, EPI><<<grid_tiles(K.nx, K.ny, K.batch), dim3(TX, TY), smem, E->ctx->stream>>>(K, src, dst, W, E->nsq_cap);
    OF2D_LAUNCH_CHECK(E->ctx);
    return OF2D_SUCCESS;
}

#define TRY(x) do { int _s = (x); if (_s) return _s; } while (0)

// one iteration of method `m`, enqueued on the context's stream
template <class R>
int enqueue_iteration(of2d_engine *E, const EngK<R> &K, const R *d_Iref) {
    cudaStream_t s = E->ctx->stream;
    const dim3 g = grid_tiles(K.nx, K.ny, K.batch), b(TX, TY);
    const of2d_engine_desc &d = E->d;
    switch (d.method) {
        case 0: {
            const R alpha = (R)d.alpha;
            k_hs_iter<R><<<g, b, 0, s>>>(K, (const vec2_t<R> *)E->gradI, (const R *)E->It, alpha * alpha);
            OF2D_LAUNCH_CHECK(E->ctx);
            break;
        }
        case 1:
            TRY(of2d_curvature_engine_step(E->plan, K.ctl, K.n_active, K.partials, K.pstride, K.tr, E->est[0], E->est[1], E->gradI, E->It));
            break;
        case 2:
            TRY(sor_tile_launch<R>(E->ctx, E->sor, K.ctl, K.n_active, K.partials, K.pstride, K.tr, 0, (vec2_t<R> *)E->est[0], (vec2_t<R> *)E->est[1], nullptr, nullptr,
                                   (const vec2_t<R> *)E->gradI, (const R *)E->It));
            break;
        case 3:
        case 4: {
            const R si = (R)d.sigma_i, sx = (R)d.sigma_x;
            k_e_demons_force<R><<<g, b, 0, s>>>(K, d_Iref, (const R *)E->aux, si * si, sx * sx);
            OF2D_LAUNCH_CHECK(E->ctx);
            if (d.method == 3) {
                TRY((launch_conv<R, 0>(E, K, B_C0, B_C1, 0)));
                k_e_compose<R><<<g, b, 0, s>>>(K, G_ACTIVE, B_EST_CUR, B_C1, B_C0, d.accumulation == 1);
                OF2D_LAUNCH_CHECK(E->ctx);
                TRY((launch_conv<R, 1>(E, K, B_C0, B_EST_NEXT, 1)));
            } else {
                TRY((launch_conv<R, 2>(E, K, B_C0, B_C1, 0)));
                for (int q = 0; q < E->nsq_cap; q++) {
                    k_e_square<R><<<g, b, 0, s>>>(K, q);
                    OF2D_LAUNCH_CHECK(E->ctx);
                }
                k_e_compose<R><<<g, b, 0, s>>>(K, G_ACTIVE, B_EST_CUR, B_CRES, B_CTMP, 0);
                OF2D_LAUNCH_CHECK(E->ctx);
                TRY((launch_conv<R, 1>(E, K, B_CTMP, B_EST_NEXT, 1)));
            }
            break;
        }
        case 5: {
            TRY(sor_tile_launch<R>(E->ctx, E->sor, K.ctl, K.n_active, K.partials, K.pstride, K.tr, 1, (vec2_t<R> *)E->vel[0], (vec2_t<R> *)E->vel[1],
                                   (const vec2_t<R> *)E->est[0], (const vec2_t<R> *)E->est[1], (const vec2_t<R> *)E->gradI, (const R *)E->It));
            const dim3 gt(ceil_div(K.ny, TILE), ceil_div(K.nx, TILE), K.batch);
            k_fl_increment<R><<<gt, b, 0, s>>>(K, (vec2_t<R> *)E->vel[0], (vec2_t<R> *)E->vel[1], (vec2_t<R> *)E->incr);
            OF2D_LAUNCH_CHECK(E->ctx);
            k_fl_integrate<R><<<gt, b, 0, s>>>(K, (const vec2_t<R> *)E->incr);
            OF2D_LAUNCH_CHECK(E->ctx);
            // regrid (ImageRegistrationFluid.cpp:108-124): level <- est + level o (id + est); est <- 0; re-warp; derivatives
            k_e_untranspose<R><<<g, b, 0, s>>>(K, G_REGRID, B_EST_CUR, B_ESTN);
            OF2D_LAUNCH_CHECK(E->ctx);
            k_e_compose<R><<<g, b, 0, s>>>(K, G_REGRID, B_LVL_CUR, B_ESTN, B_LVL_NEXT, 0);
            OF2D_LAUNCH_CHECK(E->ctx);
            k_e_zero<R><<<dim3(E->ctx->sm_count * 2, K.batch), 256, 0, s>>>(K, G_REGRID, B_EST_NEXT, 1);
            OF2D_LAUNCH_CHECK(E->ctx);
            k_e_warp<R><<<g, b, 0, s>>>(K, G_REGRID, (const R *)E->cur_Imov, B_LVL_NEXT, (R *)E->aux);
            OF2D_LAUNCH_CHECK(E->ctx);
            k_e_derivatives<R><<<g, b, 0, s>>>(K, G_REGRID, d_Iref, (const R *)E->aux, (vec2_t<R> *)E->gradI, (R *)E->It, 1);
            OF2D_LAUNCH_CHECK(E->ctx);
            k_regrid_commit<<<ceil_div(K.batch, 128), 128, 0, s>>>(K.ctl, K.batch);
            OF2D_LAUNCH_CHECK(E->ctx);
            break;
        }
        default:
            of2d_set_error("engine: unknown method %d", d.method);
            return OF2D_ERR_INVALID;
    }
    return OF2D_SUCCESS;
}


template <class R>
int refine_impl(of2d_engine *E, const R *d_Iref, const R *d_Imov, R *d_motion, int niter) {
    of2d_ctx *ctx = E->ctx;
    cudaStream_t s = ctx->stream;
    EngK<R> K = make_k<R>(E);
    K.ext = (vec2_t<R> *)d_motion;
    const of2d_engine_desc &d = E->d;
    const dim3 g = grid_tiles(K.nx, K.ny, K.batch), b(TX, TY);
    const size_t vbytes = sizeof(vec2_t<R>) * E->n * K.batch, vbytesT = sizeof(vec2_t<R>) * E->nT * K.batch;
    if (niter > E->tr.cap) { of2d_set_error("engine: niter %d above the trace capacity %d", niter, E->tr.cap); return OF2D_ERR_INVALID; }
    E->cur_Imov = d_Imov;
    E->last_niter = niter;

    // ---- set-up: Iaux = Imov o (id + motion); derivatives; estimate = 0 (e.g. ImageRegistrationDemons.cpp:97-106)
    OF2D_CUDA_TRY(cudaMemcpyAsync(E->lvl[0], d_motion, vbytes, cudaMemcpyDeviceToDevice, s));
    k_ctl_begin<<<ceil_div(K.batch, 128), 128, 0, s>>>(K.ctl, K.batch, niter, K.n_active);
    OF2D_LAUNCH_CHECK(ctx);
    k_e_warp<R><<<g, b, 0, s>>>(K, G_NONE, d_Imov, B_LVL_CUR, (R *)E->aux);
    OF2D_LAUNCH_CHECK(ctx);
    if (d.method != 3 && d.method != 4) {
        k_e_derivatives<R><<<g, b, 0, s>>>(K, G_NONE, d_Iref, (const R *)E->aux, (vec2_t<R> *)E->gradI, (R *)E->It, E->transposed ? 1 : 0);
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
        for (int q = 0; q < m; q++) TRY(enqueue_iteration<R>(E, K, d_Iref));
        enq += m;
        E->iterations_enqueued += (uint64_t)m;
        OF2D_CUDA_TRY(cudaMemcpyAsync(&E->h_snap[slot], E->d_nactive, sizeof(int), cudaMemcpyDeviceToHost, s));
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

    // ---- tear-down: motion <- estimate + motion o (id + estimate); the estimate is dropped (:136-137)
    if (E->transposed) {
        k_e_untranspose<R><<<g, b, 0, s>>>(K, G_NONE, B_EST_CUR, B_ESTN);
        OF2D_LAUNCH_CHECK(ctx);
        k_e_compose<R><<<g, b, 0, s>>>(K, G_NONE, B_LVL_CUR, B_ESTN, B_EXT, 0);
    } else {
        k_e_compose<R><<<g, b, 0, s>>>(K, G_NONE, B_LVL_CUR, B_EST_CUR, B_EXT, 0);
    }
    OF2D_LAUNCH_CHECK(ctx);
    OF2D_CUDA_TRY(cudaMemcpyAsync(E->h_ctl.data(), E->d_ctl, sizeof(PairCtl) * K.batch, cudaMemcpyDeviceToHost, s));
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
    TRY(alloc(&E->d_taps[which], buf.size()));
    OF2D_CUDA_TRY(cudaMemcpy(E->d_taps[which], buf.data(), buf.size(), cudaMemcpyHostToDevice));
    return OF2D_SUCCESS;
}

}  // namespace

extern "C" {

int of2d_engine_create(of2d_ctx *ctx, const of2d_engine_desc *desc, of2d_engine **out) {
    *out = nullptr;
    OF2D_REQUIRE(desc && desc->dimx > 1 && desc->dimy > 1 && desc->batch > 0 && desc->method >= 0 && desc->method <= 5 && desc->max_iter >= 0, "bad engine description");
    OF2D_CUDA_TRY(cudaSetDevice(ctx->device));
    of2d_engine *E = new (std::nothrow) of2d_engine();
    OF2D_REQUIRE(E, "out of host memory");
    E->ctx = ctx; E->d = *desc; E->dbl = desc->real_is_double != 0;
    E->elem = E->dbl ? 8 : 4;
    const int nx = desc->dimx, ny = desc->dimy, B = desc->batch, m = desc->method;
    E->n = (size_t)nx * ny;
    E->transposed = (m == 2 || m == 5);
    int st = OF2D_SUCCESS;
    auto fail = [&](int code) { of2d_engine_destroy(E); return code; };
    if (E->transposed) {
        E->sor = sor_plan(nx, ny, B, desc->mu, desc->lambda, desc->omega, E->dbl);
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
    if (E->transposed || m == 0 || m == 1) { if ((st = alloc(&E->estN, v * nN))) return fail(st); }
    if (m == 5) {
        for (int k = 0; k < 2; k++) if ((st = alloc(&E->vel[k], v * nTt))) return fail(st);
        if ((st = alloc(&E->incr, v * nTt))) return fail(st);
    }
    if ((st = alloc((void **)&E->d_ctl, sizeof(PairCtl) * B))) return fail(st);
    if ((st = alloc((void **)&E->d_nactive, sizeof(int)))) return fail(st);
    // partials: up to 3 values per CTA of the widest grid (32 x 32 tiles, rows of the curvature pass, SOR tiles)
    size_t nblk = (size_t)ceil_div(nx, TILE) * ceil_div(ny, TILE);
    if ((size_t)ny > nblk) nblk = ny;
    if (E->transposed && (size_t)E->sor.nbands * E->sor.nstrips > nblk) nblk = (size_t)E->sor.nbands * E->sor.nstrips;
    E->pstride = nblk * 3 + 8;
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
        if (!(desc->kernel_w > 0 && desc->kernel_w <= 2 * kConvMaxHalf + 1 && desc->kernel_fluid && desc->kernel_diffusion)) {
            of2d_set_error("engine: demons needs two Gaussian kernels of width <= %d", 2 * kConvMaxHalf + 1);
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
    if (m == 1) {
        if ((st = of2d_curvature_plan_create(ctx, nx, ny, desc->alpha, desc->tau, E->dbl ? 1 : 0, &E->plan))) return fail(st);
        if ((st = of2d_curvature_plan_set_batch(E->plan, B))) return fail(st);
    }
    if (cudaHostAlloc((void **)&E->h_snap, sizeof(int) * 2, cudaHostAllocDefault) != cudaSuccess) { of2d_set_error("engine: pinned allocation failed"); return fail(OF2D_ERR_CUDA); }
    for (int k = 0; k < 2; k++)
        if (cudaEventCreateWithFlags(&E->ev[k], cudaEventDisableTiming) != cudaSuccess) { of2d_set_error("engine: event creation failed"); return fail(OF2D_ERR_CUDA); }
    E->h_ctl.resize((size_t)B);
    *out = E;
    return OF2D_SUCCESS;
}

void of2d_engine_destroy(of2d_engine *E) {
    if (!E) return;
    cudaSetDevice(E->ctx->device);
    cudaStreamSynchronize(E->ctx->stream);
    void *bufs[] = {E->aux, E->gradI, E->It, E->est[0], E->est[1], E->c[0], E->c[1], E->lvl[0], E->lvl[1], E->estN, E->vel[0], E->vel[1], E->incr,
                    E->d_ctl, E->d_nactive, E->d_partials, E->tr.err, E->tr.maxabs, E->tr.dt, E->tr.minjac, E->tr.regrid, E->tr.nsq, E->d_taps[0], E->d_taps[1]};
    for (void *p : bufs) if (p) cudaFree(p);
    if (E->plan) of2d_curvature_plan_destroy(E->plan);
    if (E->h_snap) cudaFreeHost(E->h_snap);
    for (int k = 0; k < 2; k++) if (E->ev[k]) cudaEventDestroy(E->ev[k]);
    delete E;
}

int of2d_engine_reset_state(of2d_engine *E) {
    if (E->d.method != 5) return OF2D_SUCCESS;
    const size_t bytes = 2 * E->elem * E->nT * E->d.batch;
    OF2D_CUDA_TRY(cudaMemsetAsync(E->vel[0], 0, bytes, E->ctx->stream));
    OF2D_CUDA_TRY(cudaMemsetAsync(E->vel[1], 0, bytes, E->ctx->stream));
    return OF2D_SUCCESS;
}

int of2d_engine_refine_f32(of2d_engine *E, const float *d_Iref, const float *d_Imov, float *d_motion, int niter) {
    OF2D_REQUIRE(!E->dbl, "engine was created for double fields");
    return refine_impl<float>(E, d_Iref, d_Imov, d_motion, niter);
}
int of2d_engine_refine_f64(of2d_engine *E, const double *d_Iref, const double *d_Imov, double *d_motion, int niter) {
    OF2D_REQUIRE(E->dbl, "engine was created for float fields");
    return refine_impl<double>(E, d_Iref, d_Imov, d_motion, niter);
}

int of2d_engine_pair_result(of2d_engine *E, int pair, int *iterations, int *nregrid, double *last_err) {
    OF2D_REQUIRE(pair >= 0 && pair < E->d.batch, "pair out of range");
    const PairCtl &c = E->h_ctl[(size_t)pair];
    if (iterations) *iterations = c.iter;
    if (nregrid) *nregrid = c.nregrid;
    if (last_err) *last_err = c.err;
    return OF2D_SUCCESS;
}

int of2d_engine_trace(of2d_engine *E, int pair, int which, double *h_out, int count) {
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

uint64_t of2d_engine_iterations_enqueued(of2d_engine *E) { return E->iterations_enqueued; }

}  // extern "C"
