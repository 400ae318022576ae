// device_math.cuh -- per-pixel device functions shared by the primitive and the fused solver
// kernels.  Each one evaluates the reference's expression for a single pixel in the reference's
// operation order (the library is built with -fmad=false), so fused kernels and stand-alone
// primitives give identical bits.
#pragma once

#include "common.cuh"

// OF2D_RELAXED (engine_relaxed.cu only): arithmetic level 2.  The translation unit is compiled with FMA contraction and
// approximate division, and the functions below may take shortcuts that are algebraically equal to the reference's
// expression (marked `#if OF2D_RELAXED`); every other translation unit sees OF2D_RELAXED == 0 and the literal expressions.
#ifndef OF2D_RELAXED
#define OF2D_RELAXED 0
#endif

// ---- gradients.h:9-32 on a scalar image ----------------------------------------------------------
template <class R>
__device__ __forceinline__ R partial_x(const R *__restrict__ f, int idx, int i, int nx) {
    if (i == 0) return f[idx + 1] - f[idx];
    if (i == nx - 1) return f[idx] - f[idx - 1];
    return (f[idx + 1] - f[idx - 1]) / (R)2.0f;
}
template <class R>
__device__ __forceinline__ R partial_y(const R *__restrict__ f, int idx, int j, int nx, int ny) {
    if (j == 0) return f[idx + nx] - f[idx];
    if (j == ny - 1) return f[idx] - f[idx - nx];
    return (f[idx + nx] - f[idx - nx]) / (R)2.0f;
}
// ---- the same templates on vector2d ----------------------------------------------------------------
template <class R>
__device__ __forceinline__ vec2_t<R> partial_x_v(const vec2_t<R> *__restrict__ f, int idx, int i, int nx) {
    if (i == 0) { const vec2_t<R> a = f[idx + 1], b = f[idx]; return mk2<R>(a.x - b.x, a.y - b.y); }
    if (i == nx - 1) { const vec2_t<R> a = f[idx], b = f[idx - 1]; return mk2<R>(a.x - b.x, a.y - b.y); }
    const vec2_t<R> a = f[idx + 1], b = f[idx - 1];
    return mk2<R>((a.x - b.x) / (R)2.0f, (a.y - b.y) / (R)2.0f);
}
template <class R>
__device__ __forceinline__ vec2_t<R> partial_y_v(const vec2_t<R> *__restrict__ f, int idx, int j, int nx, int ny) {
    if (j == 0) { const vec2_t<R> a = f[idx + nx], b = f[idx]; return mk2<R>(a.x - b.x, a.y - b.y); }
    if (j == ny - 1) { const vec2_t<R> a = f[idx], b = f[idx - nx]; return mk2<R>(a.x - b.x, a.y - b.y); }
    const vec2_t<R> a = f[idx + nx], b = f[idx - nx];
    return mk2<R>((a.x - b.x) / (R)2.0f, (a.y - b.y) / (R)2.0f);
}

// ---- Image::jacobian, Image.cpp:205-214 ------------------------------------------------------------
template <class R>
__device__ __forceinline__ R jacobian_pixel(const vec2_t<R> *__restrict__ u, int idx, int i, int j, int nx, int ny) {
    const vec2_t<R> dudx = partial_x_v<R>(u, idx, i, nx);
    const vec2_t<R> dudy = partial_y_v<R>(u, idx, j, nx, ny);
    return ((R)1.0f + dudx.x) * ((R)1.0f + dudy.y) - dudx.y * dudy.x;
}

// ---- bilinear sampling geometry shared by warp2d / accumulate (Image.cpp:144-151, Motion.cpp:137-144)
template <class R>
struct Bilin {
    int idxO;
    R fx, fy;
    bool inside, hx, hy;
};
template <class R>
__device__ __forceinline__ Bilin<R> bilin_setup(int i, int j, R ux, R uy, int nx, int ny) {
    Bilin<R> b;
#if OF2D_RELAXED
    // (R)dx == floor(px) exactly for every coordinate that can be inside the image: the same bits without the int -> real conversion
    const R px = (R)i + ux, flx = r_floor(px); const int dx = (int)flx; b.fx = px - flx;
    const R py = (R)j + uy, fly = r_floor(py); const int dy = (int)fly; b.fy = py - fly;
#else
    const R px = (R)i + ux; const int dx = (int)r_floor(px); b.fx = px - (R)dx;
    const R py = (R)j + uy; const int dy = (int)r_floor(py); b.fy = py - (R)dy;
#endif
    b.inside = !(dx < 0 || dx >= nx || dy < 0 || dy >= ny);
    b.idxO = dx + dy * nx;
    b.hx = dx < nx - 1;
    b.hy = dy < ny - 1;
    return b;
}

// Image::warp2d for one pixel; `keep` is the value the pixel has before the warp (Image.cpp:148-173)
template <class R>
__device__ __forceinline__ R warp_pixel(const R *__restrict__ src, int nx, int ny, int i, int j, vec2_t<R> u, R keep) {
    const Bilin<R> b = bilin_setup<R>(i, j, u.x, u.y, nx, ny);
    if (!b.inside) return keep;
#if OF2D_RELAXED
    if (b.hx && b.hy) {   // all four taps inside: the weights sum to 1 (the reference's float sum is 1 +- 1 ulp), so no renormalisation
        const R s00 = src[b.idxO], s10 = src[b.idxO + 1], s01 = src[b.idxO + nx], s11 = src[b.idxO + nx + 1];
        const R lo = s00 + b.fx * (s10 - s00), hi = s01 + b.fx * (s11 - s01);
        return lo + b.fy * (hi - lo);
    }
#endif
    const R one = (R)1;
    R val = src[b.idxO] * (one - b.fx) * (one - b.fy);
    R weight = (one - b.fx) * (one - b.fy);
    if (b.hx) { val += src[b.idxO + 1] * b.fx * (one - b.fy); weight += b.fx * (one - b.fy); }
    if (b.hy) { val += src[b.idxO + nx] * (one - b.fx) * b.fy; weight += (one - b.fx) * b.fy; }
    if (b.hx && b.hy) { val += src[b.idxO + 1 + nx] * b.fx * b.fy; weight += b.fx * b.fy; }
    return weight != 0 ? val / weight : keep;
}

// the same, with the value the pixel keeps (src[idx_self]) read only when the reference keeps it
template <class R>
__device__ __forceinline__ R warp_pixel_lazy(const R *__restrict__ src, int nx, int ny, int i, int j, vec2_t<R> u, int idx_self) {
    const Bilin<R> b = bilin_setup<R>(i, j, u.x, u.y, nx, ny);
    if (!b.inside) return src[idx_self];
#if OF2D_RELAXED
    if (b.hx && b.hy) {
        const R s00 = src[b.idxO], s10 = src[b.idxO + 1], s01 = src[b.idxO + nx], s11 = src[b.idxO + nx + 1];
        const R lo = s00 + b.fx * (s10 - s00), hi = s01 + b.fx * (s11 - s01);
        return lo + b.fy * (hi - lo);
    }
#endif
    const R one = (R)1;
    R val = src[b.idxO] * (one - b.fx) * (one - b.fy);
    R weight = (one - b.fx) * (one - b.fy);
    if (b.hx) { val += src[b.idxO + 1] * b.fx * (one - b.fy); weight += b.fx * (one - b.fy); }
    if (b.hy) { val += src[b.idxO + nx] * (one - b.fx) * b.fy; weight += (one - b.fx) * b.fy; }
    if (b.hx && b.hy) { val += src[b.idxO + 1 + nx] * b.fx * b.fy; weight += b.fx * b.fy; }
    return weight != 0 ? val / weight : src[idx_self];
}

// Motion::accumulate for one pixel: v + u o (id + v); `keep` = u at this pixel (Motion.cpp:141-169)
template <class R>
__device__ __forceinline__ vec2_t<R> compose_pixel(const vec2_t<R> *__restrict__ u, int nx, int ny, int i, int j, vec2_t<R> v, vec2_t<R> keep) {
    const Bilin<R> b = bilin_setup<R>(i, j, v.x, v.y, nx, ny);
    if (!b.inside) return keep;
#if OF2D_RELAXED
    if (b.hx && b.hy) {
        const vec2_t<R> s00 = u[b.idxO], s10 = u[b.idxO + 1], s01 = u[b.idxO + nx], s11 = u[b.idxO + nx + 1];
        const R lx = s00.x + b.fx * (s10.x - s00.x), hx = s01.x + b.fx * (s11.x - s01.x);
        const R ly = s00.y + b.fx * (s10.y - s00.y), hy = s01.y + b.fx * (s11.y - s01.y);
        return mk2<R>(v.x + (lx + b.fy * (hx - lx)), v.y + (ly + b.fy * (hy - ly)));
    }
#endif
    const R one = (R)1;
    vec2_t<R> s = u[b.idxO];
    R vx = s.x * (one - b.fx) * (one - b.fy), vy = s.y * (one - b.fx) * (one - b.fy);
    R weight = (one - b.fx) * (one - b.fy);
    if (b.hx) { s = u[b.idxO + 1]; vx += s.x * b.fx * (one - b.fy); vy += s.y * b.fx * (one - b.fy); weight += b.fx * (one - b.fy); }
    if (b.hy) { s = u[b.idxO + nx]; vx += s.x * (one - b.fx) * b.fy; vy += s.y * (one - b.fx) * b.fy; weight += (one - b.fx) * b.fy; }
    if (b.hx && b.hy) { s = u[b.idxO + 1 + nx]; vx += s.x * b.fx * b.fy; vy += s.y * b.fx * b.fy; weight += b.fx * b.fy; }
    if (weight != 0) return mk2<R>(v.x + vx / weight, v.y + vy / weight);
    return v;
}

// The same two operations with the four taps already in registers: the batched fast paths of the engine issue
// every gather of a thread's pixels before any arithmetic (taps the reference does not read are loaded from a
// clamped address and ignored), then evaluate the reference's expressions unchanged.
template <class R>
struct BilinTaps {
    int o, ox, oy;   // clamped gather addresses: o, o + ox, o + oy, o + ox + oy
};
template <class R>
__device__ __forceinline__ BilinTaps<R> bilin_taps(const Bilin<R> &b, int idx_self, int nx) {
    BilinTaps<R> t;
    t.o = b.inside ? b.idxO : idx_self;
    t.ox = (b.inside && b.hx) ? 1 : 0;
    t.oy = (b.inside && b.hy) ? nx : 0;
    return t;
}
template <class R>
__device__ __forceinline__ vec2_t<R> compose_taps(const Bilin<R> &b, vec2_t<R> s00, vec2_t<R> s10, vec2_t<R> s01, vec2_t<R> s11, vec2_t<R> v, vec2_t<R> keep) {
    if (!b.inside) return keep;
#if OF2D_RELAXED
    if (b.hx && b.hy) {
        const R lx = s00.x + b.fx * (s10.x - s00.x), hx = s01.x + b.fx * (s11.x - s01.x);
        const R ly = s00.y + b.fx * (s10.y - s00.y), hy = s01.y + b.fx * (s11.y - s01.y);
        return mk2<R>(v.x + (lx + b.fy * (hx - lx)), v.y + (ly + b.fy * (hy - ly)));
    }
#endif
    const R one = (R)1;
    R vx = s00.x * (one - b.fx) * (one - b.fy), vy = s00.y * (one - b.fx) * (one - b.fy);
    R weight = (one - b.fx) * (one - b.fy);
    if (b.hx) { vx += s10.x * b.fx * (one - b.fy); vy += s10.y * b.fx * (one - b.fy); weight += b.fx * (one - b.fy); }
    if (b.hy) { vx += s01.x * (one - b.fx) * b.fy; vy += s01.y * (one - b.fx) * b.fy; weight += (one - b.fx) * b.fy; }
    if (b.hx && b.hy) { vx += s11.x * b.fx * b.fy; vy += s11.y * b.fx * b.fy; weight += b.fx * b.fy; }
    if (weight != 0) return mk2<R>(v.x + vx / weight, v.y + vy / weight);
    return v;
}
template <class R>
__device__ __forceinline__ R warp_taps(const Bilin<R> &b, R s00, R s10, R s01, R s11, R keep) {
    if (!b.inside) return keep;
#if OF2D_RELAXED
    if (b.hx && b.hy) {
        const R lo = s00 + b.fx * (s10 - s00), hi = s01 + b.fx * (s11 - s01);
        return lo + b.fy * (hi - lo);
    }
#endif
    const R one = (R)1;
    R val = s00 * (one - b.fx) * (one - b.fy);
    R weight = (one - b.fx) * (one - b.fy);
    if (b.hx) { val += s10 * b.fx * (one - b.fy); weight += b.fx * (one - b.fy); }
    if (b.hy) { val += s01 * (one - b.fx) * b.fy; weight += (one - b.fx) * b.fy; }
    if (b.hx && b.hy) { val += s11 * b.fx * b.fy; weight += b.fx * b.fy; }
    return weight != 0 ? val / weight : keep;
}

// ---- reductions' per-pixel terms --------------------------------------------------------------------
// Motion::norm addend, Motion.cpp:45: sqrt(pow(x,2)+pow(y,2)) evaluated in double
template <class R>
__device__ __forceinline__ double vec_norm_d(vec2_t<R> v) {
    const double x = (double)v.x, y = (double)v.y;
    return sqrt(x * x + y * y);
}
// Motion::maxabs term, Motion.cpp:54: pow(y,2)+pow(y,2) (x is ignored), rounded to real
template <class R>
__device__ __forceinline__ R maxabs_term(vec2_t<R> v) {
    const double y = (double)v.y;
    return (R)(y * y + y * y);
}
// float data: y*y is exact in double (48 bits) and so is its doubling, so the one rounding to float is RN(2 y^2) = 2 RN(y^2):
// the same bits from two float instructions (no double pipe)
template <>
__device__ __forceinline__ float maxabs_term<float>(float2 v) {
    const float t = v.y * v.y;
    return t + t;
}

// ---- OpticalFlow::get_force, OpticalFlow.cpp:33 -------------------------------------------------------
template <class R>
__device__ __forceinline__ vec2_t<R> lssd_force(vec2_t<R> dI, R It, vec2_t<R> u) {
    const R s = It + u.x * dI.x + u.y * dI.y;
    return mk2<R>(dI.x * s, dI.y * s);
}
