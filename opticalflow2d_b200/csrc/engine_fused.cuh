// engine_fused.cuh -- relaxed build only (OF2D_RELAXED, arithmetic level 2): a Demons iteration in TWO kernels.
//
//   k_rx_force_conv    warp of the moving image -> derivatives -> demons force -> Gaussian smoothing of the correspondence
//                      (DemonsThirions.cpp:18-32 / DemonsDiffeomorphic.cpp:15-24: Image::warp2d, IterativeSolver::set_derivatives,
//                      Demons::demons_iteration Demons.cpp:34-63, Motion::convolute with K_fluid)
//   k_rx_compose_conv  u <- c + u o (id + c) (or u + c) -> Gaussian smoothing of the motion -> Logger norms
//                      (DemonsThirions.cpp:33-42: Motion::accumulate Motion.cpp:113-178, Motion::convolute with K_diffusion, Logger.cpp:32-51)
//
// Per 32 x 32 tile a CTA keeps every intermediate in shared memory: the source window the bilinear taps read (48 x 48, one TMA
// bulk copy per row), the warped image on the tile + (cx + 1) halo, the unsmoothed field on the tile + cx halo; only the
// smoothed result reaches HBM: 24 + 24 B/px per iteration, the algorithmic minimum of SURVEY 8(d), instead of 88.
//
// Addressing is by FLAT index throughout, which is what makes the reference's convolution semantics (Field.tpp:245-248: the
// bounds test is on the linear index, so x taps wrap into the neighbouring row) fall out without special cases: window
// element (r, c) of a tile stands for flat index lin = (j0 - h + r) nx + (i0 - h + c), i.e. for the ACTUAL pixel
// (lin mod nx, lin div nx), and everything a pixel needs from its neighbours (central / one-sided differences) sits at
// lin +- 1, lin +- nx, i.e. at (r, c +- 1), (r +- 1, c) of the same window.  Interior tiles take a path without any of the
// validity / wrap / border tests.
#pragma once

#if OF2D_RELAXED

namespace {

// resident CTAs per SM the register allocation of the fp32 kernels aims at.  Measured (Thirion / Diffeomorphic, 50 iterations at 2048^2):
// 4 CTAs (64 registers): 6.78 / 7.35 ms; 3 CTAs (80 registers): 5.73 / 6.83 ms; 2 CTAs (118 registers): 7.03 / 8.15 ms
#ifndef OF2D_FUSED_MINB
#define OF2D_FUSED_MINB 3
#endif
#ifndef OF2D_FUSED_MINB_K1
#define OF2D_FUSED_MINB_K1 OF2D_FUSED_MINB
#endif
#ifndef OF2D_FUSED_MINB_K2
#define OF2D_FUSED_MINB_K2 OF2D_FUSED_MINB
#endif
constexpr int FW = 48;   // side of the staged source window
constexpr int FO = 8;    // the tile starts at (FO, FO) of the window: origin (i0 - 8, j0 - 8) keeps every row 32-byte aligned

template <int KW> struct FusedGeom {
    static constexpr int CX = (KW - 1) / 2;      // convolution halo
    static constexpr int HW = CX + 1;            // halo of the warped image (differences)
    static constexpr int WW = TILE + 2 * HW, WP = WW + 1;
    static constexpr int CW = TILE + 2 * CX, CP = CW;
    static constexpr int NWE = (WW * WW + TX * TY - 1) / (TX * TY);
    static constexpr int NCE = (CW * CW + TX * TY - 1) / (TX * TY);
};

// stages the FW x FW window with origin (wx0, wy0) of `src` (flat addressing): TMA row copies when every row is 16-byte
// aligned and inside the array, else plain loads with the flat bounds test.  Returns true when the caller has to wait on `bar`.
template <class E>
__device__ __forceinline__ bool stage_window(E *sI, const E *__restrict__ src, int nx, long n, int wx0, int wy0, uint64_t *bar, int tid) {
    const long first = (long)wy0 * nx + wx0, last = (long)(wy0 + FW - 1) * nx + (wx0 + FW - 1);
    const bool tma = ((size_t)nx * sizeof(E)) % 16 == 0 && first >= 0 && last < n;
    if (tma) {
        if (tid < 32) {
            if (tid == 0) { proxy_fence_async(); mbar_expect_tx(bar, (unsigned)(FW * FW * sizeof(E))); }
            __syncwarp();
            for (int r = tid; r < FW; r += 32) bulk_g2s(sI + r * FW, src + first + (long)r * nx, (unsigned)(FW * sizeof(E)), bar);
        }
    } else {
        for (int e = tid; e < FW * FW; e += TX * TY) {
            const int r = e / FW, cc = e - r * FW;
            const long lin = first + (long)r * nx + cc;
            E v;
            memset(&v, 0, sizeof(E));
            if (lin >= 0 && lin < n) v = src[lin];
            sI[e] = v;
        }
    }
    return tma;
}

// flat window position -> actual pixel; returns false when the flat index is outside the field
__device__ __forceinline__ bool flat_pixel(int xw, int yw, int nx, long n, int &x, int &y, long &lin) {   // (fields of < 2^31 elements: engine_create)
    lin = (long)yw * nx + xw;
    x = xw; y = yw;
    if (xw < 0) { x = xw + nx; y = yw - 1; }
    else if (xw >= nx) { x = xw - nx; y = yw + 1; }
    return lin >= 0 && lin < n;
}

// the window positions whose four bilinear taps are inside the window AND inside the image: lx in [lx0, lx0 + lxn), ly likewise
struct TapBox {
    int lx0, ly0;
    unsigned lxn, lyn;
    __device__ __forceinline__ TapBox(int wx0, int wy0, int nx, int ny) {
        lx0 = max(0, -wx0); ly0 = max(0, -wy0);
        const int lx1 = min(FW - 2, nx - 2 - wx0), ly1 = min(FW - 2, ny - 2 - wy0);
        lxn = lx1 >= lx0 ? (unsigned)(lx1 - lx0 + 1) : 0u;
        lyn = ly1 >= ly0 ? (unsigned)(ly1 - ly0 + 1) : 0u;
    }
};

// Image::warp2d for pixel (x, y) (Image.cpp:144-173), every other case than "four taps inside the window": taps from global
// memory, the reference's renormalisation on the last row / column, the kept value outside the image.  Recomputes the sampling
// position from (x, y, u), so the hot path keeps nothing alive for it.
template <class R>
__device__ __noinline__ R warp_slow(const R *__restrict__ src, int nx, int ny, int x, int y, R ux, R uy) {
    const R px = (R)x + ux, flx = r_floor(px), py = (R)y + uy, fly = r_floor(py);
    const int dx = (int)flx, dy = (int)fly;
    const R fx = px - flx, fy = py - fly;
    const long self = (long)y * nx + x;
    if (dx < 0 || dx >= nx || dy < 0 || dy >= ny) return src[self];
    const long o = (long)dy * nx + dx;
    if (dx < nx - 1 && dy < ny - 1) {
        const R s00 = src[o], s10 = src[o + 1], s01 = src[o + nx], s11 = src[o + nx + 1];
        const R lo = s00 + fx * (s10 - s00), hi = s01 + fx * (s11 - s01);
        return lo + fy * (hi - lo);
    }
    const R one = (R)1;
    R val = src[o] * (one - fx) * (one - fy), weight = (one - fx) * (one - fy);
    if (dx < nx - 1) { val += src[o + 1] * fx * (one - fy); weight += fx * (one - fy); }
    if (dy < ny - 1) { val += src[o + nx] * (one - fx) * fy; weight += (one - fx) * fy; }
    return weight != 0 ? val / weight : src[self];
}
// straight-line hot path (taps from the staged window, weights sum to 1: no renormalisation); the rare rest goes to warp_slow.
// (xf, yf) = (float)(x, y); (ox, oy) = window origin; the tap box is passed as four scalars.
template <class R>
__device__ __forceinline__ R warp_windowed(const R *sI, const R *__restrict__ src, int nx, int ny, int ox, int oy, int lx0, unsigned lxn, int ly0, unsigned lyn,
                                            int x, int y, R xf, R yf, vec2_t<R> u) {
    const R px = xf + u.x, flx = r_floor(px), py = yf + u.y, fly = r_floor(py);
    const int lx = (int)flx - ox, ly = (int)fly - oy;
    const R fx = px - flx, fy = py - fly;
    const bool ok = (unsigned)(lx - lx0) < lxn && (unsigned)(ly - ly0) < lyn;
    const R *p = sI + (ok ? ly * FW + lx : 0);
    const R s00 = p[0], s10 = p[1], s01 = p[FW], s11 = p[FW + 1];
    const R lo = s00 + fx * (s10 - s00), hi = s01 + fx * (s11 - s01);
    R val = lo + fy * (hi - lo);
    if (!ok) val = warp_slow<R>(src, nx, ny, x, y, u.x, u.y);
    return val;
}

// Motion::accumulate for pixel (x, y) (Motion.cpp:137-169): v + u o (id + v); outside the image the pixel keeps u
template <class R>
__device__ __noinline__ vec2_t<R> compose_slow(const vec2_t<R> *__restrict__ u, int nx, int ny, int x, int y, R vx_, R vy_) {
    using V = vec2_t<R>;
    const V v = mk2<R>(vx_, vy_);
    const R px = (R)x + v.x, flx = r_floor(px), py = (R)y + v.y, fly = r_floor(py);
    const int dx = (int)flx, dy = (int)fly;
    const R fx = px - flx, fy = py - fly;
    if (dx < 0 || dx >= nx || dy < 0 || dy >= ny) return u[(long)y * nx + x];
    const long o = (long)dy * nx + dx;
    if (dx < nx - 1 && dy < ny - 1) {
        const V s00 = u[o], s10 = u[o + 1], s01 = u[o + nx], s11 = u[o + nx + 1];
        const R lx_ = s00.x + fx * (s10.x - s00.x), hx_ = s01.x + fx * (s11.x - s01.x);
        const R ly_ = s00.y + fx * (s10.y - s00.y), hy_ = s01.y + fx * (s11.y - s01.y);
        return mk2<R>(v.x + (lx_ + fy * (hx_ - lx_)), v.y + (ly_ + fy * (hy_ - ly_)));
    }
    const R one = (R)1;
    V s = u[o];
    R vx = s.x * (one - fx) * (one - fy), vy = s.y * (one - fx) * (one - fy), weight = (one - fx) * (one - fy);
    if (dx < nx - 1) { s = u[o + 1]; vx += s.x * fx * (one - fy); vy += s.y * fx * (one - fy); weight += fx * (one - fy); }
    if (dy < ny - 1) { s = u[o + nx]; vx += s.x * (one - fx) * fy; vy += s.y * (one - fx) * fy; weight += (one - fx) * fy; }
    if (weight != 0) return mk2<R>(v.x + vx / weight, v.y + vy / weight);
    return v;
}
template <class R>
__device__ __forceinline__ vec2_t<R> compose_windowed(const vec2_t<R> *sU, const vec2_t<R> *__restrict__ u, int nx, int ny, int ox, int oy, int lx0, unsigned lxn, int ly0, unsigned lyn,
                                                      int x, int y, R xf, R yf, vec2_t<R> v) {
    using V = vec2_t<R>;
    const R px = xf + v.x, flx = r_floor(px), py = yf + v.y, fly = r_floor(py);
    const int lx = (int)flx - ox, ly = (int)fly - oy;
    const R fx = px - flx, fy = py - fly;
    const bool ok = (unsigned)(lx - lx0) < lxn && (unsigned)(ly - ly0) < lyn;
    const V *p = sU + (ok ? ly * FW + lx : 0);
    const V s00 = p[0], s10 = p[1], s01 = p[FW], s11 = p[FW + 1];
    const R lx_ = s00.x + fx * (s10.x - s00.x), hx_ = s01.x + fx * (s11.x - s01.x);
    const R ly_ = s00.y + fx * (s10.y - s00.y), hy_ = s01.y + fx * (s11.y - s01.y);
    V o = mk2<R>(v.x + (lx_ + fy * (hx_ - lx_)), v.y + (ly_ + fy * (hy_ - ly_)));
    if (!ok) o = compose_slow<R>(u, nx, ny, x, y, v.x, v.y);
    return o;
}

// Gaussian smoothing of the field held in shared memory (tile + CX halo, flat addressing) for the thread's 4 vertically
// adjacent pixels (column threadIdx.x, rows 4 threadIdx.y ..): separable taps where every tap of the window is inside the
// array, the reference's dense loop with its renormalisation (Field.tpp:236-262) on the first / last CX rows of the field
template <class R, int KW, int CP, class Emit>
__device__ __forceinline__ void fused_conv_p(const vec2_t<R> *sC, const ConvW<R> &W, int i0, int j0, int nx, int ny, long n, bool fast, Emit emit) {   // CP: row pitch of sC
    using V = vec2_t<R>;
    using G = FusedGeom<KW>;
    constexpr int CX = G::CX, NRW = 4 + KW - 1;
    const int x = threadIdx.x, jl0 = 4 * threadIdx.y;
    // a tap's flat index leaves [0, n) for rows < CX and >= ny - CX, and on the first / last CX pixels of rows CX and ny - 1 - CX
    const bool dense = !fast && (j0 + jl0 <= CX || j0 + jl0 + 3 >= ny - 1 - CX);
    if (!dense) {
        if constexpr (sizeof(R) == 4) {
            unsigned long long hrow[NRW];
#pragma unroll
            for (int r = 0; r < NRW; r++) {
                const float2 e0 = sC[(jl0 + r) * CP + x];
                hrow[r] = fma_f32x2(pack_f32x2(e0.x, e0.y), pack_f32x2(W.sx[0], W.sx[0]), pack_f32x2(0.0f, 0.0f));
#pragma unroll
                for (int ii = 1; ii < KW; ii++) {
                    const float2 e = sC[(jl0 + r) * CP + x + ii];
                    hrow[r] = fma_f32x2(pack_f32x2(e.x, e.y), pack_f32x2(W.sx[ii], W.sx[ii]), hrow[r]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                unsigned long long a2 = fma_f32x2(hrow[q], pack_f32x2(W.sy[0], W.sy[0]), pack_f32x2(0.0f, 0.0f));
#pragma unroll
                for (int jj = 1; jj < KW; jj++) a2 = fma_f32x2(hrow[q + jj], pack_f32x2(W.sy[jj], W.sy[jj]), a2);
                const float2 o = unpack_f32x2(a2);
                emit(q, o);
            }
        } else {
            V hrow[NRW];
#pragma unroll
            for (int r = 0; r < NRW; r++) {
                R hx = (R)0, hy = (R)0;
#pragma unroll
                for (int ii = 0; ii < KW; ii++) {
                    const V e = sC[(jl0 + r) * CP + x + ii];
                    hx += e.x * W.sx[ii]; hy += e.y * W.sx[ii];
                }
                hrow[r] = mk2<R>(hx, hy);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                R ox = (R)0, oy = (R)0;
#pragma unroll
                for (int jj = 0; jj < KW; jj++) { ox += hrow[q + jj].x * W.sy[jj]; oy += hrow[q + jj].y * W.sy[jj]; }
                emit(q, mk2<R>(ox, oy));
            }
        }
        return;
    }
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        const long idx = (long)(i0 + x) + (long)(j0 + jl0 + q) * nx;
        R ax = (R)0, ay = (R)0;
        double weight = 0.0;
        for (int ii = -CX; ii <= CX; ii++) {
            for (int jj = -CX; jj <= CX; jj++) {
                const long lin = idx + ii + (long)jj * nx;
                if (lin < 0 || lin >= n) continue;
                const int ik = (ii + CX) + (jj + CX) * KW;
                weight += W.taps_d[ik];
                const V f = sC[(jl0 + q + jj + CX) * CP + (x + ii + CX)];
                ax += f.x * W.w[ik];
                ay += f.y * W.w[ik];
            }
        }
        V o;
        if (weight != 0) { const R wg = (R)weight; o = mk2<R>(ax / wg, ay / wg); }
        else o = sC[(jl0 + q + CX) * CP + x + CX];
        emit(q, o);
    }
}

template <class R, int KW, class Emit>
__device__ __forceinline__ void fused_conv(const vec2_t<R> *sC, const ConvW<R> &W, int i0, int j0, int nx, int ny, long n, bool fast, Emit emit) {
    fused_conv_p<R, KW, FusedGeom<KW>::CP>(sC, W, i0, j0, nx, ny, n, fast, emit);
}

template <class R, int KW> constexpr size_t fused_smem_force() {
    using G = FusedGeom<KW>;
    return sizeof(R) * FW * FW + sizeof(vec2_t<R>) * G::CW * G::CP + sizeof(R) * G::WW * G::WP;
}
template <class R, int KW> constexpr size_t fused_smem_compose() {
    using G = FusedGeom<KW>;
    return sizeof(vec2_t<R>) * (FW * FW + G::CW * G::CP);
}

// Window elements -> threads.  A window of side TILE + 2 H is the 32 columns of the tile (H .. H + 31: thread column
// threadIdx.x, rows threadIdx.y + 8 k, so a warp reads one aligned 32-element row segment and consecutive k differ by the
// constant 8 nx) plus two side strips of H columns each (2 H (TILE + 2 H) elements, one per thread in a single extra pass).
template <int H> struct Strip {
    static constexpr int W = TILE + 2 * H;
    static constexpr int NMAIN = (W + TY - 1) / TY;          // passes over the tile's columns
    static constexpr int NSIDE = 2 * H * W;                   // elements of the two side strips (<= TX * TY for H <= 4)
    static_assert(NSIDE <= TX * TY, "side strips take one pass");
    int sr, sc;      // the thread's side-strip element (row, column of the window); sr < 0: none
    __device__ __forceinline__ explicit Strip(int tid) {
        sr = -1; sc = 0;
        if (tid < NSIDE) { sr = tid / (2 * H); const int q = tid - sr * (2 * H); sc = q < H ? q : q + TILE; }
    }
};

// ---------------------------------------------------------------------------------------------
// kernel 1: warp -> derivatives -> demons force -> smoothing with K_fluid
//   EPI 0: none (Thirion)   EPI 2: maxabs of the smoothed correspondence -> number of squarings (Diffeomorphic, Motion.cpp:253-260)
// ---------------------------------------------------------------------------------------------
template <class R, int KW, bool FAST>
struct ForceConvTile {
    using V = vec2_t<R>;
    using G = FusedGeom<KW>;
    static constexpr int CX = G::CX, HW = G::HW, WW = G::WW, WP = G::WP, CW = G::CW, CP = G::CP;

    // element (r, cc) of the warped-image window: value of u there (0 outside the field)
    static __device__ __forceinline__ V load_u(const V *__restrict__ u, int nx, int n, int base, int r, int cc) {
        const int lin = base + r * nx + cc;
        if (FAST) return u[lin];
        return (lin >= 0 && lin < n) ? u[lin] : mk2<R>((R)0, (R)0);
    }
    static __device__ __forceinline__ R load_s(const R *__restrict__ f, int nx, int n, int base, int r, int cc) {
        const int lin = base + r * nx + cc;
        if (FAST) return f[lin];
        return (lin >= 0 && lin < n) ? f[lin] : (R)0;
    }
    static __device__ __forceinline__ void warp_elem(R *sW, const R *sI, const R *__restrict__ Imov, int nx, int ny, int n, int i0, int j0, const TapBox &tb, int r, int cc, V uu) {
        int x = i0 - HW + cc, y = j0 - HW + r;
        bool valid = true;
        if (!FAST) { long lin; valid = flat_pixel(i0 - HW + cc, j0 - HW + r, nx, n, x, y, lin); }
        sW[r * WP + cc] = valid ? warp_windowed<R>(sI, Imov, nx, ny, i0 - FO, j0 - FO, tb.lx0, tb.lxn, tb.ly0, tb.lyn, x, y, (R)x, (R)y, uu) : (R)0;
    }
    static __device__ __forceinline__ void force_elem(V *sC, const R *sW, int nx, int ny, int n, int i0, int j0, int r, int cc, R iref, R sratio, bool &divzero) {
        const R *w = sW + (r + 1) * WP + (cc + 1);
        const R ce = w[0];
        R gx, gy;
        bool valid = true;
        if (FAST) {
            gx = (w[1] - w[-1]) * (R)0.5f;
            gy = (w[WP] - w[-WP]) * (R)0.5f;
        } else {
            int x, y;
            long lin;
            valid = flat_pixel(i0 - CX + cc, j0 - CX + r, nx, n, x, y, lin);
            // gradients.h:9-32: one-sided on the first / last column and row
            gx = x == 0 ? w[1] - ce : x == nx - 1 ? ce - w[-1] : (w[1] - w[-1]) * (R)0.5f;
            gy = y == 0 ? w[WP] - ce : y == ny - 1 ? ce - w[-WP] : (w[WP] - w[-WP]) * (R)0.5f;
        }
        const R It = ce - iref;
        const R den = gx * gx + gy * gy + (It * It) * sratio;
        V cv = mk2<R>((R)0, (R)0);
        if (valid) {
            if (den == 0) divzero = true;
            else {
                R s;
                if constexpr (sizeof(R) == 4) s = __fdividef(-It, den);
                else s = -It / den;
                cv = mk2<R>(gx * s, gy * s);
            }
        }
        sC[r * CP + cc] = cv;
    }

    template <int EPI>
    static __device__ __forceinline__ void run(R *sI, V *sC, R *sW, uint64_t *bar, unsigned &uses, const Strip<HW> &SW_, const Strip<CX> &SC_,
                                               const V *__restrict__ u, const R *__restrict__ Iref, const R *__restrict__ Imov, V *__restrict__ out,
                                               int nx, int ny, int n, int i0, int j0, R sratio, const ConvW<R> &W, bool &divzero, R &mx, const bool prestaged = false) {
        const int tid = threadIdx.x + threadIdx.y * TX;
        const int tx = threadIdx.x, ty = threadIdx.y;
        // prestaged: the caller has filled sI with the window (zeros outside the field: no tap box reaches them) and waited for it
        const bool tma = prestaged ? false : stage_window<R>(sI, Imov, nx, (long)n, i0 - FO, j0 - FO, bar, tid);
        // loads that do not depend on the window: u on the warped-image window, Iref on the correspondence window
        constexpr int NW = Strip<HW>::NMAIN, NC = Strip<CX>::NMAIN;
        const int baseW = (j0 - HW) * nx + (i0 - HW), baseC = (j0 - CX) * nx + (i0 - CX);
        V uu[NW], uside = mk2<R>((R)0, (R)0);
        R ir[NC], irside = (R)0;
#pragma unroll
        for (int k = 0; k < NW; k++) { const int r = min(ty + TY * k, WW - 1); uu[k] = load_u(u, nx, n, baseW, r, HW + tx); }
        if (SW_.sr >= 0) uside = load_u(u, nx, n, baseW, SW_.sr, SW_.sc);
#pragma unroll
        for (int k = 0; k < NC; k++) { const int r = min(ty + TY * k, CW - 1); ir[k] = load_s(Iref, nx, n, baseC, r, CX + tx); }
        if (SC_.sr >= 0) irside = load_s(Iref, nx, n, baseC, SC_.sr, SC_.sc);
        if (tma) { mbar_wait(bar, uses & 1u); uses++; }
        else if (!prestaged) __syncthreads();
        // ---- warped image on the tile + HW halo
        const TapBox tb(i0 - FO, j0 - FO, nx, ny);
#pragma unroll
        for (int k = 0; k < NW; k++) { const int r = ty + TY * k; if (r < WW) warp_elem(sW, sI, Imov, nx, ny, n, i0, j0, tb, r, HW + tx, uu[k]); }
        if (SW_.sr >= 0) warp_elem(sW, sI, Imov, nx, ny, n, i0, j0, tb, SW_.sr, SW_.sc, uside);
        __syncthreads();
        // ---- derivatives of the warped image + demons force on the tile + CX halo
#pragma unroll
        for (int k = 0; k < NC; k++) { const int r = ty + TY * k; if (r < CW) force_elem(sC, sW, nx, ny, n, i0, j0, r, CX + tx, ir[k], sratio, divzero); }
        if (SC_.sr >= 0) force_elem(sC, sW, nx, ny, n, i0, j0, SC_.sr, SC_.sc, irside, sratio, divzero);
        __syncthreads();
        // ---- smoothing
        const int i = i0 + tx, jb = j0 + 4 * ty;
        V *op = out + (i + jb * nx);
        fused_conv<R, KW>(sC, W, i0, j0, nx, ny, (long)n, FAST, [&](int q, V o) {
            if (FAST || (i < nx && jb + q < ny)) {
                op[q * nx] = o;
                if (EPI == 2) { const R s = maxabs_term<R>(o); mx = mx < s ? s : mx; }
            }
        });
    }
};

// Diffeomorphic: maxabs of the smoothed correspondence -> number of squarings of Motion::exp (Motion.cpp:253-260), decided by the last CTA
template <class R>
__device__ __forceinline__ void demons_nsquares_epilogue(const EngK<R> &K, PairCtl *c, int pair, R mx, int nsq_cap) {
    const int tid = threadIdx.x + threadIdx.y * TX;
    mx = block_extreme<R, true>(mx);
    const double vals[1] = {(double)mx};
    double *part = K.partials + (size_t)pair * K.pstride;
    if (publish_partials<1>(vals, part, &c->ticket[1], gridDim.x, blockIdx.x)) {
        double o1[1];
        reduce_partials<1>(part, gridDim.x, o1, 1u, 0u);
        if (tid == 0) {
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

template <class R, int EPI, int KW>
__global__ void __launch_bounds__(TX *TY, sizeof(R) == 4 ? OF2D_FUSED_MINB_K1 : 2)
k_rx_force_conv(EngK<R> K, const R *__restrict__ Iref_all, const R *__restrict__ Imov_all, R sratio, const __grid_constant__ ConvW<R> W, int dst_buf, int nsq_cap, int nofast) {
    pdl_enter();
    using V = vec2_t<R>;
    using G = FusedGeom<KW>;
    // fp32: static shared arrays (every address a compile-time constant); fp64 needs more than the 48 KiB static limit
    constexpr bool STATIC = sizeof(R) == 4;
    __shared__ __align__(128) R sI_static[STATIC ? FW * FW : 1];
    __shared__ __align__(16) V sC_static[STATIC ? G::CW * G::CP : 1];
    __shared__ __align__(16) R sW_static[STATIC ? G::WW * G::WP : 1];
    extern __shared__ __align__(128) unsigned char smem_dynamic[];
    R *sI, *sW;                                  // [FW][FW] source window (Imov); [WW][WP] warped image
    V *sC;                                       // [CW][CP] unsmoothed correspondence
    if constexpr (STATIC) { sI = sI_static; sC = sC_static; sW = sW_static; }
    else { sI = reinterpret_cast<R *>(smem_dynamic); sC = reinterpret_cast<V *>(sI + FW * FW); sW = reinterpret_cast<R *>(sC + G::CW * G::CP); }
    __shared__ uint64_t bar;
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny, n = (int)K.n;
    const R *__restrict__ Iref = Iref_all + (size_t)pair * K.n;
    const R *__restrict__ Imov = Imov_all + (size_t)pair * K.n;
    const V *__restrict__ u = pick(K, B_EST_CUR, h, pair);
    V *__restrict__ out = pick(K, dst_buf, h, pair);
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    unsigned uses = 0u;
    const Strip<G::HW> SW_(tid);
    const Strip<G::CX> SC_(tid);
    const TileWalk T(nx, ny);
    bool divzero = false;
    R mx = (R)0;
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        __syncthreads();   // the previous tile's reads of the shared arrays are over (first tile: the barrier is initialised)
        // the warped-image window inside the image: central differences everywhere, every flat index valid, no wrap
        const bool fast = !nofast && i0 >= G::HW && j0 >= G::HW && i0 + TILE + G::HW <= nx && j0 + TILE + G::HW <= ny;
        if (fast) ForceConvTile<R, KW, true>::template run<EPI>(sI, sC, sW, &bar, uses, SW_, SC_, u, Iref, Imov, out, nx, ny, n, i0, j0, sratio, W, divzero, mx);
        else ForceConvTile<R, KW, false>::template run<EPI>(sI, sC, sW, &bar, uses, SW_, SC_, u, Iref, Imov, out, nx, ny, n, i0, j0, sratio, W, divzero, mx);
    }
    if (divzero) atomicOr(&c->flags, OF2D_FLAG_DIVZERO);
    if (EPI == 2) demons_nsquares_epilogue<R>(K, c, pair, mx, nsq_cap);
}

// ---------------------------------------------------------------------------------------------
// kernel 2: u' = v + u o (id + v) (add_only: u + v) -> smoothing with K_diffusion -> next estimate + Logger norms
// ---------------------------------------------------------------------------------------------
template <class R, int KW, bool FAST>
struct ComposeConvTile {
    using V = vec2_t<R>;
    using G = FusedGeom<KW>;
    static constexpr int CX = G::CX, CW = G::CW, CP = G::CP;

    static __device__ __forceinline__ V load_v(const V *__restrict__ v, int nx, int n, int base, int r, int cc) {
        const int lin = base + r * nx + cc;
        if (FAST) return v[lin];
        return (lin >= 0 && lin < n) ? v[lin] : mk2<R>((R)0, (R)0);
    }
    static __device__ __forceinline__ void compose_elem(V *sS, const V *sU, const V *__restrict__ u, int nx, int ny, int n, int i0, int j0, const TapBox &tb, int r, int cc, V vv, int add_only) {
        int x = i0 - CX + cc, y = j0 - CX + r;
        long lin = 0;
        bool valid = true;
        if (!FAST) valid = flat_pixel(i0 - CX + cc, j0 - CX + r, nx, n, x, y, lin);
        V s = mk2<R>((R)0, (R)0);
        if (valid) {
            if (add_only) {
                const V uc = FAST ? sU[(r - CX + FO) * FW + (cc - CX + FO)] : u[lin];
                s = mk2<R>(uc.x + vv.x, uc.y + vv.y);
            } else s = compose_windowed<R>(sU, u, nx, ny, i0 - FO, j0 - FO, tb.lx0, tb.lxn, tb.ly0, tb.lyn, x, y, (R)x, (R)y, vv);
        }
        sS[r * CP + cc] = s;
    }
    static __device__ __forceinline__ void run(V *sU, V *sS, uint64_t *bar, unsigned &uses, const Strip<CX> &SC_, const V *__restrict__ u, const V *__restrict__ v,
                                               V *__restrict__ out, int nx, int ny, int n, int i0, int j0, int add_only, const ConvW<R> &W, NormAcc<R> &acc,
                                               const bool prestaged = false) {
        const int tid = threadIdx.x + threadIdx.y * TX;
        const int tx = threadIdx.x, ty = threadIdx.y;
        const bool tma = prestaged ? false : stage_window<V>(sU, u, nx, (long)n, i0 - FO, j0 - FO, bar, tid);
        constexpr int NC = Strip<CX>::NMAIN;
        const int baseC = (j0 - CX) * nx + (i0 - CX);
        V vv[NC], vside = mk2<R>((R)0, (R)0);
#pragma unroll
        for (int k = 0; k < NC; k++) { const int r = min(ty + TY * k, CW - 1); vv[k] = load_v(v, nx, n, baseC, r, CX + tx); }
        if (SC_.sr >= 0) vside = load_v(v, nx, n, baseC, SC_.sr, SC_.sc);
        if (tma) { mbar_wait(bar, uses & 1u); uses++; }
        else if (!prestaged) __syncthreads();
        const TapBox tb(i0 - FO, j0 - FO, nx, ny);
#pragma unroll
        for (int k = 0; k < NC; k++) { const int r = ty + TY * k; if (r < CW) compose_elem(sS, sU, u, nx, ny, n, i0, j0, tb, r, CX + tx, vv[k], add_only); }
        if (SC_.sr >= 0) compose_elem(sS, sU, u, nx, ny, n, i0, j0, tb, SC_.sr, SC_.sc, vside, add_only);
        __syncthreads();
        const int i = i0 + tx, jb = j0 + 4 * ty;
        V *op = out + (i + jb * nx);
        const V *prev = sU + (4 * ty + FO) * FW + tx + FO;   // Logger's prev: the current estimate at the pixel (in the window)
        fused_conv<R, KW>(sS, W, i0, j0, nx, ny, (long)n, FAST, [&](int q, V o) {
            if (FAST || (i < nx && jb + q < ny)) {
                op[q * nx] = o;
                acc.add(o, prev[q * FW]);
            }
        });
        acc.flush();
    }
};

template <class R, int KW>
__global__ void __launch_bounds__(TX *TY, sizeof(R) == 4 ? OF2D_FUSED_MINB_K2 : 2)
k_rx_compose_conv(EngK<R> K, int v_buf, int add_only, const __grid_constant__ ConvW<R> W, int nofast) {
    pdl_enter();
    using V = vec2_t<R>;
    using G = FusedGeom<KW>;
    // fp32: static shared arrays (addresses are compile-time constants); fp64 needs more than the 48 KiB static limit
    constexpr bool STATIC = sizeof(R) == 4;
    __shared__ __align__(128) V sU_static[STATIC ? FW * FW : 1];
    __shared__ __align__(16) V sS_static[STATIC ? G::CW * G::CP : 1];
    extern __shared__ __align__(128) unsigned char smem_dynamic[];
    V *sU, *sS;                                  // [FW][FW] window of the current estimate; [CW][CP] composed, unsmoothed field
    if constexpr (STATIC) { sU = sU_static; sS = sS_static; }
    else { sU = reinterpret_cast<V *>(smem_dynamic); sS = sU + FW * FW; }
    __shared__ uint64_t bar;
    const int pair = blockIdx.y;
    PairCtl *c = K.ctl + pair;
    const CtlHot h = load_ctl(c);
    if (!h.active) return;
    const int nx = K.nx, ny = K.ny, n = (int)K.n;
    const V *__restrict__ u = pick(K, B_EST_CUR, h, pair);
    const V *__restrict__ v = pick(K, v_buf, h, pair);
    V *__restrict__ out = pick(K, B_EST_NEXT, h, pair);
    const int tid = threadIdx.x + threadIdx.y * TX;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    unsigned uses = 0u;
    const Strip<G::CX> SC_(tid);
    const TileWalk T(nx, ny);
    NormAcc<R> acc;
    for (int tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const int i0 = T.tx(tile) * TILE, j0 = T.ty(tile) * TILE;
        __syncthreads();
        const bool fast = !nofast && i0 >= G::CX && j0 >= G::CX && i0 + TILE + G::CX <= nx && j0 + TILE + G::CX <= ny;
        if (fast) ComposeConvTile<R, KW, true>::run(sU, sS, &bar, uses, SC_, u, v, out, nx, ny, n, i0, j0, add_only, W, acc);
        else ComposeConvTile<R, KW, false>::run(sU, sS, &bar, uses, SC_, u, v, out, nx, ny, n, i0, j0, add_only, W, acc);
    }
    logger_epilogue<R>(K, c, pair, acc.dsd, acc.dsp);
}

}  // namespace

#endif  // OF2D_RELAXED
