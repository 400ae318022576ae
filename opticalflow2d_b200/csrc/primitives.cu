// primitives.cu -- field primitives of the registration solve (layer L1 of SURVEY.md):
// casts, bilinear warp, field composition, wrap-aware convolution, image derivatives, Jacobian,
// pointwise algebra, reductions (norm / maxabs / Logger), scaling-and-squaring, pyramid resampling.
//
// All kernels are HBM-bound streaming or short-stencil kernels: one thread per pixel on 32x8 tiles
// (x fastest => each warp touches one 128/256-byte row segment), neighbour reuse served by L1/L2.
#include <math.h>
#include <string.h>

#include "device_math.cuh"

namespace {

constexpr int TX = 32, TY = 8;

inline dim3 grid2d(int nx, int ny, int batch) { return dim3(ceil_div(nx, TX), ceil_div(ny, TY), batch); }

// ---------------------------------------------------------------------------------------------
// casts
// ---------------------------------------------------------------------------------------------
template <class Tin, class Tout>
__global__ void k_cast(size_t n, const Tin *__restrict__ in, Tout *__restrict__ out) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        out[k] = (Tout)in[k];
}

template <class R>
__global__ void k_motion_to_planar(size_t n, const vec2_t<R> *__restrict__ u, double *__restrict__ out) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const vec2_t<R> v = u[k];
        out[k] = (double)v.x;
        out[k + n] = (double)v.y;
    }
}

// the same for `batch` fields of n pixels in ONE launch (grid.y = field): out = [field][2][n]
template <class R>
__global__ void k_motion_to_planar_batch(size_t n, const vec2_t<R> *__restrict__ u, double *__restrict__ out) {
    u += (size_t)blockIdx.y * n;
    out += (size_t)blockIdx.y * 2 * n;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const vec2_t<R> v = u[k];
        out[k] = (double)v.x;
        out[k + n] = (double)v.y;
    }
}

// ---------------------------------------------------------------------------------------------
// warp / compose
// ---------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(TX *TY) k_warp(int nx, int ny, const R *__restrict__ src, const vec2_t<R> *__restrict__ u, R *__restrict__ dst) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t off = (size_t)blockIdx.z * nx * ny;
    src += off; u += off; dst += off;
    const int idx = i + j * nx;
    dst[idx] = warp_pixel<R>(src, nx, ny, i, j, u[idx], src[idx]);
}

template <class R>
__global__ void __launch_bounds__(TX *TY) k_compose(int nx, int ny, const vec2_t<R> *__restrict__ u, const vec2_t<R> *__restrict__ v, vec2_t<R> *__restrict__ out) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t off = (size_t)blockIdx.z * nx * ny;
    u += off; v += off; out += off;
    const int idx = i + j * nx;
    out[idx] = compose_pixel<R>(u, nx, ny, i, j, v[idx], u[idx]);
}

// ---------------------------------------------------------------------------------------------
// convolution with the reference's linear-index bounds test (Field.tpp:245-248)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 31 * 31;

template <class R>
struct ConvWeights {
    const R *taps;        // (real)k[idxkernel]
    const double *taps_d; // k[idxkernel]
    double full_weight;   // sum of all visited taps in visiting order
    int kw, cx, cy;
};

// NC = number of components (1 image, 2 motion)
template <class R, int NC, bool FAST>
__global__ void __launch_bounds__(TX *TY) k_convolute(int nx, int ny, const R *__restrict__ in, R *__restrict__ out, ConvWeights<R> W) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const long n = (long)nx * ny;
    const size_t off = (size_t)blockIdx.z * n * NC;
    in += off; out += off;
    const long idx = i + (long)j * nx;
    const int cx = W.cx, cy = W.cy, kw = W.kw;
    // every tap is inside [0, n) iff the first and the last visited linear index are
    const bool interior = (idx - cx - (long)cy * nx >= 0) && (idx + cx + (long)cy * nx < n);
    R acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = (R)0;
    double weight = 0.0;
    for (int ii = -cx; ii <= cx; ii++) {
        for (int jj = -cy; jj <= cy; jj++) {
            const long lin = idx + ii + (long)jj * nx;
            if (!interior && (lin < 0 || lin >= n)) continue;
            const int ik = (ii + cx) + (jj + cy) * kw;
            const R t = W.taps[ik];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const R f = in[lin * NC + c];
                acc[c] = FAST ? r_fma(f, t, acc[c]) : acc[c] + f * t;
            }
            if (!interior) weight += W.taps_d[ik];
        }
    }
    if (interior) weight = W.full_weight;
    if (weight != 0) {
        const R w = (R)weight;
#pragma unroll
        for (int c = 0; c < NC; c++) out[idx * NC + c] = acc[c] / w;
    } else {
#pragma unroll
        for (int c = 0; c < NC; c++) out[idx * NC + c] = in[idx * NC + c];
    }
}

// ---------------------------------------------------------------------------------------------
// derivatives, jacobian
// ---------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(TX *TY) k_derivatives(int nx, int ny, const R *__restrict__ Iref, const R *__restrict__ Imov, vec2_t<R> *__restrict__ gradI, R *__restrict__ It) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t off = (size_t)blockIdx.z * nx * ny;
    Iref += off; Imov += off; gradI += off; It += off;
    const int idx = i + j * nx;
    gradI[idx] = mk2<R>(partial_x<R>(Imov, idx, i, nx), partial_y<R>(Imov, idx, j, nx, ny));
    It[idx] = Imov[idx] - Iref[idx];
}

template <class R>
__global__ void __launch_bounds__(TX *TY) k_jacobian(int nx, int ny, const vec2_t<R> *__restrict__ u, R *__restrict__ jac, double *__restrict__ partial_min) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    R J = (R)INFINITY;
    if (i < nx && j < ny) {
        const int idx = i + j * nx;
        J = jacobian_pixel<R>(u, idx, i, j, nx, ny);
        if (jac) jac[idx] = J;
    }
    J = block_extreme<R, false>(J);
    if (threadIdx.x == 0 && threadIdx.y == 0) partial_min[blockIdx.x + blockIdx.y * gridDim.x] = (double)J;
}

__global__ void k_finalize_min(int nblocks, const double *__restrict__ partial, double *__restrict__ out) {
    double m = INFINITY;
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) m = partial[k] < m ? partial[k] : m;
    m = block_extreme<double, false>(m);
    if (threadIdx.x == 0) out[0] = m;
}

// ---------------------------------------------------------------------------------------------
// pointwise algebra
// ---------------------------------------------------------------------------------------------
template <class R>
__global__ void k_axpy(size_t n, R a, const R *__restrict__ x, R *__restrict__ y) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        y[k] = a == (R)1 ? y[k] + x[k] : (a == (R)-1 ? y[k] - x[k] : y[k] + a * x[k]);
}
template <class R>
__global__ void k_scale(size_t n, R a, R *__restrict__ x) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) x[k] *= a;
}
template <class R>
__global__ void k_scale_xy(size_t n, R ax, R ay, vec2_t<R> *__restrict__ u) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        vec2_t<R> v = u[k];
        v.x *= ax; v.y *= ay;
        u[k] = v;
    }
}
template <class R>
__global__ void k_normalize(size_t n, R lo, R hi, R *__restrict__ x) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        x[k] = (x[k] - lo) / (hi - lo);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
// mode 0: sum of |u|; mode 1: Logger (sum |u - prev|, sum |prev|, prev <- u)
template <class R, int MODE>
__global__ void __launch_bounds__(256) k_norm_partials(size_t n, const vec2_t<R> *__restrict__ u, vec2_t<R> *__restrict__ prev, double *__restrict__ partials) {
    double a = 0.0, b = 0.0;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const vec2_t<R> v = u[k];
        if (MODE == 0) {
            a += vec_norm_d<R>(v);
        } else {
            const vec2_t<R> p = prev[k];
            a += vec_norm_d<R>(mk2<R>(v.x - p.x, v.y - p.y));
            b += vec_norm_d<R>(p);
            prev[k] = v;
        }
    }
    block_sum2(a, b);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = a; partials[2 * blockIdx.x + 1] = b; }
}
__global__ void k_finalize_sum2(int nblocks, const double *__restrict__ partials, double *__restrict__ out) {
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) { a += partials[2 * k]; b += partials[2 * k + 1]; }
    block_sum2(a, b);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

// Logger finalize: the two sums plus the sticky kernel flags (read and cleared) in one mailbox write
__global__ void k_finalize_logger(int nblocks, const double *__restrict__ partials, unsigned *__restrict__ status, double *__restrict__ out) {
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) { a += partials[2 * k]; b += partials[2 * k + 1]; }
    block_sum2(a, b);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; out[2] = (double)status[0]; status[0] = 0u; }
}

template <class R>
__global__ void __launch_bounds__(256) k_maxabs_partials(size_t n, const vec2_t<R> *__restrict__ u, double *__restrict__ partials) {
    R m = (R)0;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const R s = maxabs_term<R>(u[k]);
        m = m < s ? s : m;
    }
    m = block_extreme<R, true>(m);
    if (threadIdx.x == 0) partials[blockIdx.x] = (double)m;
}
__global__ void k_finalize_max(int nblocks, const double *__restrict__ partial, double *__restrict__ out) {
    double m = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) m = partial[k] > m ? partial[k] : m;
    m = block_extreme<double, true>(m);
    if (threadIdx.x == 0) out[0] = m;
}

// image statistics: sum (float accumulate in the reference; here double partials), max (starts at 0), min
template <class R>
__global__ void __launch_bounds__(256) k_image_stats(size_t n, const R *__restrict__ x, double *__restrict__ partials) {
    double s = 0.0, dummy = 0.0;
    R mx = (R)0, mn = (R)INFINITY;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const R v = x[k];
        s += (double)v;
        mx = v > mx ? v : mx;
        mn = v < mn ? v : mn;
    }
    block_sum2(s, dummy);
    mx = block_extreme<R, true>(mx);
    mn = block_extreme<R, false>(mn);
    if (threadIdx.x == 0) { partials[3 * blockIdx.x] = s; partials[3 * blockIdx.x + 1] = (double)mx; partials[3 * blockIdx.x + 2] = (double)mn; }
}
__global__ void k_finalize_stats(int nblocks, const double *__restrict__ p, double *__restrict__ out) {
    double s = 0.0, dummy = 0.0, mx = 0.0, mn = INFINITY;
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) {
        s += p[3 * k];
        mx = p[3 * k + 1] > mx ? p[3 * k + 1] : mx;
        mn = p[3 * k + 2] < mn ? p[3 * k + 2] : mn;
    }
    block_sum2(s, dummy);
    mx = block_extreme<double, true>(mx);
    mn = block_extreme<double, false>(mn);
    if (threadIdx.x == 0) { out[0] = s; out[1] = mx; out[2] = mn; }
}

// ---------------------------------------------------------------------------------------------
// pyramid resampling (Field.tpp:76-206)
// ---------------------------------------------------------------------------------------------
template <class R, int NC>
__global__ void __launch_bounds__(TX *TY) k_downsample(int ix, int iy, const R *__restrict__ in, int ox, int oy, R *__restrict__ out) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    if (i >= ox || j >= oy) return;
    const unsigned sizein = (unsigned)ix * iy;
    const unsigned fx = ix / ox, fy = iy / oy;
    const unsigned idxin = (unsigned)i * fx + (unsigned)j * fy * ix;
    R val[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) val[c] = (R)0;
    int p = 0;
    for (unsigned ii = 0; ii < fx; ii++) {
        for (unsigned jj = 0; jj < fy; jj++) {
            const unsigned q = idxin + ii + jj * ix;
            if (q >= sizein) continue;
#pragma unroll
            for (int c = 0; c < NC; c++) val[c] += in[(size_t)q * NC + c];
            p++;
        }
    }
    if (p != 0) {
#pragma unroll
        for (int c = 0; c < NC; c++) out[((size_t)i + (size_t)j * ox) * NC + c] = val[c] / (R)p;
    }
}

template <class R, int NC>
__global__ void __launch_bounds__(TX *TY) k_upsample(int ix, int iy, const R *__restrict__ in, int ox, int oy, R *__restrict__ out) {
    const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
    if (i >= ox || j >= oy) return;
    const unsigned sizein = (unsigned)ix * iy;
    const R px = (R)i * (R)ix / (R)ox; const int dx = (int)r_floor(px); const R fx = px - (R)dx;
    const R py = (R)j * (R)iy / (R)oy; const int dy = (int)r_floor(py); const R fy = py - (R)dy;
    const unsigned idxO = (unsigned)dx + (unsigned)dy * ix;
    if (idxO >= sizein) return;
    const bool hx = (unsigned)dx < (unsigned)ix - 1, hy = (unsigned)dy < (unsigned)iy - 1;
    R weight = ((R)1 - fx) * ((R)1 - fy);
    if (hx) weight += fx * ((R)1 - fy);
    if (hy) weight += ((R)1 - fx) * fy;
    if (hx && hy) weight += fx * fy;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        R val = in[(size_t)idxO * NC + c] * ((R)1 - fx) * ((R)1 - fy);
        if (hx) val += in[(size_t)(idxO + 1) * NC + c] * fx * ((R)1 - fy);
        if (hy) val += in[(size_t)(idxO + ix) * NC + c] * ((R)1 - fx) * fy;
        if (hx && hy) val += in[(size_t)(idxO + 1 + ix) * NC + c] * fx * fy;
        if (weight != 0) out[((size_t)i + (size_t)j * ox) * NC + c] = val / weight;
    }
}

// Motion::Neumann_/Dirichlet_boundaryconditions (Motion.cpp:181-251). The reference applies its
// assignments sequentially; edges only read interior cells (or write zero), corners read interior
// cells, so the order does not matter -- except the (dimx-1, 0) corner, which reads
// u[(dimy-2) + 1*dimx] (Motion.cpp:213, y extent used for an x index), reproduced literally.
template <class R>
__global__ void k_boundary(int nx, int ny, int kind, int phase, vec2_t<R> *__restrict__ u) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const vec2_t<R> zero = mk2<R>((R)0, (R)0);
    if (phase == 0) {
        if (t >= 1 && t < nx - 1) {
            u[t] = kind ? zero : u[t + nx];
            u[t + (ny - 1) * nx] = kind ? zero : u[t + (ny - 2) * nx];
        }
        if (t >= 1 && t < ny - 1) {
            u[t * nx] = kind ? zero : u[t * nx + 1];
            u[t * nx + nx - 1] = kind ? zero : u[t * nx + nx - 2];
        }
    } else if (t == 0) {   // corners, after the edges (the reference assigns them last and in this order)
        u[0] = kind ? zero : u[1 + nx];
        u[(ny - 1) * nx] = kind ? zero : u[1 + (ny - 2) * nx];
        u[nx - 1] = kind ? zero : u[(ny - 2) + nx];
        u[(nx - 1) + (ny - 1) * nx] = kind ? zero : u[(nx - 2) + (ny - 2) * nx];
    }
}

inline int grid1d(size_t n, int sm_count) {
    const size_t want = (n + 255) / 256;
    const size_t cap = (size_t)sm_count * 8;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

template <class R>
int read_mailbox(of2d_ctx *ctx, double *vals, int count) {
    OF2D_CUDA_TRY(cudaMemcpyAsync(ctx->h_mailbox, ctx->d_mailbox, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    OF2D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(vals, ctx->h_mailbox, sizeof(double) * count);
    return OF2D_SUCCESS;
}

// ---- templated host bodies -------------------------------------------------------------------
template <class R>
int warp2d_impl(of2d_ctx *ctx, int nx, int ny, int batch, const R *src, const R *u, R *dst) {
    OF2D_REQUIRE(nx > 0 && ny > 0 && batch > 0, "bad dimensions");
    OF2D_REQUIRE(src != dst, "warp2d is out of place");
    k_warp<R><<<grid2d(nx, ny, batch), dim3(TX, TY), 0, ctx->stream>>>(nx, ny, src, (const vec2_t<R> *)u, dst);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}

template <class R>
int compose_impl(of2d_ctx *ctx, int nx, int ny, int batch, const R *u, const R *v, R *out) {
    OF2D_REQUIRE(nx > 0 && ny > 0 && batch > 0, "bad dimensions");
    OF2D_REQUIRE(u != out && v != out, "compose is out of place");
    k_compose<R><<<grid2d(nx, ny, batch), dim3(TX, TY), 0, ctx->stream>>>(nx, ny, (const vec2_t<R> *)u, (const vec2_t<R> *)v, (vec2_t<R> *)out);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}

template <class R, int NC>
int convolute_impl(of2d_ctx *ctx, int nx, int ny, int batch, const R *in, R *out, const double *h_kernel, int kw, int kh) {
    OF2D_REQUIRE(nx > 0 && ny > 0 && batch > 0, "bad dimensions");
    OF2D_REQUIRE(kw > 0 && kh > 0 && kw * kh <= kMaxTaps, "kernel too large");
    OF2D_REQUIRE(in != out, "convolute is out of place");
    const int ntaps = kw * kh;
    const size_t need = (sizeof(double) + sizeof(R)) * (size_t)ntaps;
    if (ctx->kernel_cap < need) {
        if (ctx->d_kernel) { OF2D_CUDA_TRY(cudaStreamSynchronize(ctx->stream)); OF2D_CUDA_TRY(cudaFree(ctx->d_kernel)); }
        OF2D_CUDA_TRY(cudaMalloc(&ctx->d_kernel, need));
        ctx->kernel_cap = need;
    }
    // host staging: doubles followed by reals; the visiting-order weight sum for interior pixels
    double hbuf_d[kMaxTaps];
    R hbuf_r[kMaxTaps];
    ConvWeights<R> W;
    W.kw = kw; W.cx = (kw - 1) / 2; W.cy = (kh - 1) / 2;
    for (int t = 0; t < ntaps; t++) { hbuf_d[t] = h_kernel[t]; hbuf_r[t] = (R)h_kernel[t]; }
    double full = 0.0;
    for (int ii = -W.cx; ii <= W.cx; ii++)
        for (int jj = -W.cy; jj <= W.cy; jj++) full += h_kernel[(ii + W.cx) + (jj + W.cy) * kw];
    W.full_weight = full;
    // weights are tiny: synchronous copies keep the host buffers' lifetime trivial
    OF2D_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    OF2D_CUDA_TRY(cudaMemcpy(ctx->d_kernel, hbuf_d, sizeof(double) * ntaps, cudaMemcpyHostToDevice));
    OF2D_CUDA_TRY(cudaMemcpy((char *)ctx->d_kernel + sizeof(double) * ntaps, hbuf_r, sizeof(R) * ntaps, cudaMemcpyHostToDevice));
    W.taps_d = (const double *)ctx->d_kernel;
    W.taps = (const R *)((char *)ctx->d_kernel + sizeof(double) * ntaps);
    if (ctx->fast_math)
        k_convolute<R, NC, true><<<grid2d(nx, ny, batch), dim3(TX, TY), 0, ctx->stream>>>(nx, ny, in, out, W);
    else
        k_convolute<R, NC, false><<<grid2d(nx, ny, batch), dim3(TX, TY), 0, ctx->stream>>>(nx, ny, in, out, W);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}

template <class R>
int derivatives_impl(of2d_ctx *ctx, int nx, int ny, int batch, const R *Iref, const R *Imov, R *gradI, R *It) {
    OF2D_REQUIRE(nx > 1 && ny > 1 && batch > 0, "derivatives need at least 2x2 pixels");
    k_derivatives<R><<<grid2d(nx, ny, batch), dim3(TX, TY), 0, ctx->stream>>>(nx, ny, Iref, Imov, (vec2_t<R> *)gradI, It);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}

template <class R>
int jacobian_impl(of2d_ctx *ctx, int nx, int ny, const R *u, R *jac, R *h_min) {
    OF2D_REQUIRE(nx > 1 && ny > 1, "jacobian needs at least 2x2 pixels");
    const dim3 g = grid2d(nx, ny, 1);
    const int nb = g.x * g.y;
    OF2D_REQUIRE(nb <= kMaxPartialBlocks * 4, "image too large for the reduction scratch");
    k_jacobian<R><<<g, dim3(TX, TY), 0, ctx->stream>>>(nx, ny, (const vec2_t<R> *)u, jac, ctx->d_partials);
    OF2D_LAUNCH_CHECK(ctx);
    if (h_min) {
        k_finalize_min<<<1, 256, 0, ctx->stream>>>(nb, ctx->d_partials, (double *)ctx->d_mailbox);
        OF2D_LAUNCH_CHECK(ctx);
        double v;
        int st = read_mailbox<R>(ctx, &v, 1);
        if (st) return st;
        *h_min = (R)v;
    }
    return OF2D_SUCCESS;
}

template <class R>
int norm_impl(of2d_ctx *ctx, size_t n, const R *u, R *h_norm) {
    OF2D_REQUIRE(n > 0, "empty field");
    const int nb = grid1d(n, ctx->sm_count);
    k_norm_partials<R, 0><<<nb, 256, 0, ctx->stream>>>(n, (const vec2_t<R> *)u, nullptr, ctx->d_partials);
    OF2D_LAUNCH_CHECK(ctx);
    k_finalize_sum2<<<1, 256, 0, ctx->stream>>>(nb, ctx->d_partials, (double *)ctx->d_mailbox);
    OF2D_LAUNCH_CHECK(ctx);
    double v[2];
    int st = read_mailbox<R>(ctx, v, 2);
    if (st) return st;
    *h_norm = (R)v[0] / (R)(unsigned)n;   // Motion.cpp:47: (float) sum / sizein
    return OF2D_SUCCESS;
}

template <class R>
int logger_impl(of2d_ctx *ctx, size_t n, const R *u, R *prev, R *h_err) {
    OF2D_REQUIRE(n > 0, "empty field");
    const int nb = grid1d(n, ctx->sm_count);
    k_norm_partials<R, 1><<<nb, 256, 0, ctx->stream>>>(n, (const vec2_t<R> *)u, (vec2_t<R> *)prev, ctx->d_partials);
    OF2D_LAUNCH_CHECK(ctx);
    k_finalize_logger<<<1, 256, 0, ctx->stream>>>(nb, ctx->d_partials, ctx->d_status, (double *)ctx->d_mailbox);
    OF2D_LAUNCH_CHECK(ctx);
    double v[3];
    int st = read_mailbox<R>(ctx, v, 3);
    if (st) return st;
    if (((unsigned)v[2]) & OF2D_FLAG_DIVZERO) {
        of2d_set_error("Divide by zero exception");
        return OF2D_ERR_DIVZERO;
    }
    const R diffnorm = (R)v[0] / (R)(unsigned)n, prevnorm = (R)v[1] / (R)(unsigned)n;
    *h_err = prevnorm == 0 ? (R)0.0f : diffnorm / prevnorm;   // Logger.cpp:39
    return OF2D_SUCCESS;
}

template <class R>
int maxabs_impl(of2d_ctx *ctx, size_t n, const R *u, R *h_maxabs) {
    OF2D_REQUIRE(n > 0, "empty field");
    const int nb = grid1d(n, ctx->sm_count);
    k_maxabs_partials<R><<<nb, 256, 0, ctx->stream>>>(n, (const vec2_t<R> *)u, ctx->d_partials);
    OF2D_LAUNCH_CHECK(ctx);
    k_finalize_max<<<1, 256, 0, ctx->stream>>>(nb, ctx->d_partials, (double *)ctx->d_mailbox);
    OF2D_LAUNCH_CHECK(ctx);
    double v;
    int st = read_mailbox<R>(ctx, &v, 1);
    if (st) return st;
    *h_maxabs = sizeof(R) == 4 ? (R)sqrtf((float)v) : (R)sqrt(v);   // Motion.cpp:57
    return OF2D_SUCCESS;
}

template <class R>
int stats_impl(of2d_ctx *ctx, size_t n, const R *x, R *h_sum, R *h_max, R *h_min) {
    OF2D_REQUIRE(n > 0, "empty field");
    const int nb = grid1d(n, ctx->sm_count);
    k_image_stats<R><<<nb, 256, 0, ctx->stream>>>(n, x, ctx->d_partials);
    OF2D_LAUNCH_CHECK(ctx);
    k_finalize_stats<<<1, 256, 0, ctx->stream>>>(nb, ctx->d_partials, (double *)ctx->d_mailbox);
    OF2D_LAUNCH_CHECK(ctx);
    double v[3];
    int st = read_mailbox<R>(ctx, v, 3);
    if (st) return st;
    if (h_sum) *h_sum = (R)v[0];
    if (h_max) *h_max = (R)v[1];
    if (h_min) *h_min = (R)v[2];
    return OF2D_SUCCESS;
}

template <class R>
int exp_impl(of2d_ctx *ctx, int nx, int ny, R *u, R *tmp, int *h_nsquares) {
    const size_t n = (size_t)nx * ny;
    R ma;
    int st = maxabs_impl<R>(ctx, n, u, &ma);
    if (st) return st;
    int nsquares = 0;
    if (ma != 0) {   // log2(0) = -inf: the reference's int cast is UB and lands on 0 after the clamp
        nsquares = sizeof(R) == 4 ? (int)ceilf(1 + log2f((float)ma)) : (int)ceil(1 + log2((double)ma));
        if (nsquares < 0) nsquares = 0;
    }
    if (h_nsquares) *h_nsquares = nsquares;
    if (nsquares == 0) return OF2D_SUCCESS;
    const R scale = (R)pow(2, -nsquares);
    k_scale<R><<<grid1d(2 * n, ctx->sm_count), 256, 0, ctx->stream>>>(2 * n, scale, u);
    OF2D_LAUNCH_CHECK(ctx);
    R *cur = u, *nxt = tmp;
    for (int s = 0; s < nsquares; s++) {
        st = compose_impl<R>(ctx, nx, ny, 1, cur, cur, nxt);
        if (st) return st;
        R *t = cur; cur = nxt; nxt = t;
    }
    if (cur != u) OF2D_CUDA_TRY(cudaMemcpyAsync(u, cur, sizeof(R) * 2 * n, cudaMemcpyDeviceToDevice, ctx->stream));
    return OF2D_SUCCESS;
}

template <class R>
int resample_impl(of2d_ctx *ctx, int ncomp, int ix, int iy, const R *in, int ox, int oy, R *out, bool up) {
    OF2D_REQUIRE(ncomp == 1 || ncomp == 2, "ncomp must be 1 or 2");
    OF2D_REQUIRE(ix > 0 && iy > 0 && ox > 0 && oy > 0, "bad dimensions");
    if (up) OF2D_REQUIRE(ox >= ix && oy >= iy, "input has to have same dimensions as target");
    else OF2D_REQUIRE(ox <= ix && oy <= iy, "input has to have same dimensions as target");
    const dim3 g = grid2d(ox, oy, 1), b(TX, TY);
    if (up) {
        if (ncomp == 1) k_upsample<R, 1><<<g, b, 0, ctx->stream>>>(ix, iy, in, ox, oy, out);
        else k_upsample<R, 2><<<g, b, 0, ctx->stream>>>(ix, iy, in, ox, oy, out);
    } else {
        if (ncomp == 1) k_downsample<R, 1><<<g, b, 0, ctx->stream>>>(ix, iy, in, ox, oy, out);
        else k_downsample<R, 2><<<g, b, 0, ctx->stream>>>(ix, iy, in, ox, oy, out);
    }
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------------
#define G1(n) grid1d((n), ctx->sm_count), 256, 0, ctx->stream

template <class R>
static int planar_batch_impl(of2d_ctx *ctx, size_t n, int batch, const R *u, double *out) {
    OF2D_REQUIRE(batch > 0 && batch <= 65535, "batch out of range");
    const unsigned gx = (unsigned)((n + 1023) / 1024 < 64 ? (n + 1023) / 1024 : 64);   // few, long CTAs per field: the launch shares the device with a running solve
    k_motion_to_planar_batch<R><<<dim3(gx ? gx : 1, (unsigned)batch), 256, 0, ctx->stream>>>(n, (const vec2_t<R> *)u, out);
    OF2D_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" {

int of2d_image_from_double_f32(of2d_ctx *ctx, size_t n, const double *in, float *out) { k_cast<double, float><<<G1(n)>>>(n, in, out); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_image_from_double_f64(of2d_ctx *ctx, size_t n, const double *in, double *out) { k_cast<double, double><<<G1(n)>>>(n, in, out); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_image_to_double_f32(of2d_ctx *ctx, size_t n, const float *in, double *out) { k_cast<float, double><<<G1(n)>>>(n, in, out); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_image_to_double_f64(of2d_ctx *ctx, size_t n, const double *in, double *out) { k_cast<double, double><<<G1(n)>>>(n, in, out); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_motion_to_planar_double_f32(of2d_ctx *ctx, size_t n, const float *u, double *out) { k_motion_to_planar<float><<<G1(n)>>>(n, (const float2 *)u, out); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_motion_to_planar_double_f64(of2d_ctx *ctx, size_t n, const double *u, double *out) { k_motion_to_planar<double><<<G1(n)>>>(n, (const double2 *)u, out); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_motion_to_planar_double_batch_f32(of2d_ctx *ctx, size_t n, int batch, const float *u, double *out) { return planar_batch_impl<float>(ctx, n, batch, u, out); }
int of2d_motion_to_planar_double_batch_f64(of2d_ctx *ctx, size_t n, int batch, const double *u, double *out) { return planar_batch_impl<double>(ctx, n, batch, u, out); }

int of2d_warp2d_f32(of2d_ctx *ctx, int nx, int ny, int batch, const float *s, const float *u, float *d) { return warp2d_impl<float>(ctx, nx, ny, batch, s, u, d); }
int of2d_warp2d_f64(of2d_ctx *ctx, int nx, int ny, int batch, const double *s, const double *u, double *d) { return warp2d_impl<double>(ctx, nx, ny, batch, s, u, d); }
int of2d_compose_f32(of2d_ctx *ctx, int nx, int ny, int batch, const float *u, const float *v, float *o) { return compose_impl<float>(ctx, nx, ny, batch, u, v, o); }
int of2d_compose_f64(of2d_ctx *ctx, int nx, int ny, int batch, const double *u, const double *v, double *o) { return compose_impl<double>(ctx, nx, ny, batch, u, v, o); }
int of2d_convolute_motion_f32(of2d_ctx *ctx, int nx, int ny, int batch, const float *in, float *out, const double *k, int kw, int kh) { return convolute_impl<float, 2>(ctx, nx, ny, batch, in, out, k, kw, kh); }
int of2d_convolute_motion_f64(of2d_ctx *ctx, int nx, int ny, int batch, const double *in, double *out, const double *k, int kw, int kh) { return convolute_impl<double, 2>(ctx, nx, ny, batch, in, out, k, kw, kh); }
int of2d_convolute_image_f32(of2d_ctx *ctx, int nx, int ny, int batch, const float *in, float *out, const double *k, int kw, int kh) { return convolute_impl<float, 1>(ctx, nx, ny, batch, in, out, k, kw, kh); }
int of2d_convolute_image_f64(of2d_ctx *ctx, int nx, int ny, int batch, const double *in, double *out, const double *k, int kw, int kh) { return convolute_impl<double, 1>(ctx, nx, ny, batch, in, out, k, kw, kh); }
int of2d_derivatives_f32(of2d_ctx *ctx, int nx, int ny, int batch, const float *r, const float *m, float *g, float *t) { return derivatives_impl<float>(ctx, nx, ny, batch, r, m, g, t); }
int of2d_derivatives_f64(of2d_ctx *ctx, int nx, int ny, int batch, const double *r, const double *m, double *g, double *t) { return derivatives_impl<double>(ctx, nx, ny, batch, r, m, g, t); }
int of2d_jacobian_f32(of2d_ctx *ctx, int nx, int ny, const float *u, float *jac, float *h_min) { return jacobian_impl<float>(ctx, nx, ny, u, jac, h_min); }
int of2d_jacobian_f64(of2d_ctx *ctx, int nx, int ny, const double *u, double *jac, double *h_min) { return jacobian_impl<double>(ctx, nx, ny, u, jac, h_min); }

int of2d_axpy_f32(of2d_ctx *ctx, size_t n, float a, const float *x, float *y) { k_axpy<float><<<G1(n)>>>(n, a, x, y); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_axpy_f64(of2d_ctx *ctx, size_t n, double a, const double *x, double *y) { k_axpy<double><<<G1(n)>>>(n, a, x, y); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_scale_f32(of2d_ctx *ctx, size_t n, float a, float *x) { k_scale<float><<<G1(n)>>>(n, a, x); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_scale_f64(of2d_ctx *ctx, size_t n, double a, double *x) { k_scale<double><<<G1(n)>>>(n, a, x); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_scale_xy_f32(of2d_ctx *ctx, size_t n, float ax, float ay, float *u) { k_scale_xy<float><<<G1(n)>>>(n, ax, ay, (float2 *)u); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_scale_xy_f64(of2d_ctx *ctx, size_t n, double ax, double ay, double *u) { k_scale_xy<double><<<G1(n)>>>(n, ax, ay, (double2 *)u); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_motion_norm_f32(of2d_ctx *ctx, size_t n, const float *u, float *h) { return norm_impl<float>(ctx, n, u, h); }
int of2d_motion_norm_f64(of2d_ctx *ctx, size_t n, const double *u, double *h) { return norm_impl<double>(ctx, n, u, h); }
int of2d_motion_maxabs_f32(of2d_ctx *ctx, size_t n, const float *u, float *h) { return maxabs_impl<float>(ctx, n, u, h); }
int of2d_motion_maxabs_f64(of2d_ctx *ctx, size_t n, const double *u, double *h) { return maxabs_impl<double>(ctx, n, u, h); }
int of2d_image_stats_f32(of2d_ctx *ctx, size_t n, const float *x, float *s, float *mx, float *mn) { return stats_impl<float>(ctx, n, x, s, mx, mn); }
int of2d_image_stats_f64(of2d_ctx *ctx, size_t n, const double *x, double *s, double *mx, double *mn) { return stats_impl<double>(ctx, n, x, s, mx, mn); }
int of2d_image_normalize_f32(of2d_ctx *ctx, size_t n, float lo, float hi, float *x) { k_normalize<float><<<G1(n)>>>(n, lo, hi, x); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_image_normalize_f64(of2d_ctx *ctx, size_t n, double lo, double hi, double *x) { k_normalize<double><<<G1(n)>>>(n, lo, hi, x); OF2D_LAUNCH_CHECK(ctx); return 0; }
int of2d_motion_exp_f32(of2d_ctx *ctx, int nx, int ny, float *u, float *tmp, int *h) { return exp_impl<float>(ctx, nx, ny, u, tmp, h); }
int of2d_motion_exp_f64(of2d_ctx *ctx, int nx, int ny, double *u, double *tmp, int *h) { return exp_impl<double>(ctx, nx, ny, u, tmp, h); }
int of2d_downsample_f32(of2d_ctx *ctx, int nc, int ix, int iy, const float *in, int ox, int oy, float *out) { return resample_impl<float>(ctx, nc, ix, iy, in, ox, oy, out, false); }
int of2d_downsample_f64(of2d_ctx *ctx, int nc, int ix, int iy, const double *in, int ox, int oy, double *out) { return resample_impl<double>(ctx, nc, ix, iy, in, ox, oy, out, false); }
int of2d_upsample_f32(of2d_ctx *ctx, int nc, int ix, int iy, const float *in, int ox, int oy, float *out) { return resample_impl<float>(ctx, nc, ix, iy, in, ox, oy, out, true); }
int of2d_upsample_f64(of2d_ctx *ctx, int nc, int ix, int iy, const double *in, int ox, int oy, double *out) { return resample_impl<double>(ctx, nc, ix, iy, in, ox, oy, out, true); }
int of2d_boundary_conditions_f32(of2d_ctx *ctx, int nx, int ny, int kind, float *u) {
    const int m = nx > ny ? nx : ny;
    k_boundary<float><<<ceil_div(m, 256), 256, 0, ctx->stream>>>(nx, ny, kind, 0, (float2 *)u); OF2D_LAUNCH_CHECK(ctx);
    k_boundary<float><<<1, 32, 0, ctx->stream>>>(nx, ny, kind, 1, (float2 *)u); OF2D_LAUNCH_CHECK(ctx); return 0;
}
int of2d_boundary_conditions_f64(of2d_ctx *ctx, int nx, int ny, int kind, double *u) {
    const int m = nx > ny ? nx : ny;
    k_boundary<double><<<ceil_div(m, 256), 256, 0, ctx->stream>>>(nx, ny, kind, 0, (double2 *)u); OF2D_LAUNCH_CHECK(ctx);
    k_boundary<double><<<1, 32, 0, ctx->stream>>>(nx, ny, kind, 1, (double2 *)u); OF2D_LAUNCH_CHECK(ctx); return 0;
}
int of2d_logger_update_f32(of2d_ctx *ctx, size_t n, const float *u, float *prev, float *h) { return logger_impl<float>(ctx, n, u, prev, h); }
int of2d_logger_update_f64(of2d_ctx *ctx, size_t n, const double *u, double *prev, double *h) { return logger_impl<double>(ctx, n, u, prev, h); }

}  // extern "C"
