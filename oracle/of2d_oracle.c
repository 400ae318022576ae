/*
 * of2d_oracle.c -- plain-C CPU restatement of the reference's registration
 * solve.  TEST INFRASTRUCTURE ONLY (see of2d_oracle.h for the parity pin).
 * Every function cites the reference file:line it follows; citations are
 * relative to /root/reference.  Compile with -ffp-contract=off: the reference
 * is built for baseline x86-64, where no FMA contraction can happen.
 */
#include "of2d_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "stubs/fftw3.h"

/* ------------------------------------------------------------------------------------------
 * precision-dependent libm selection: where the reference calls an overloaded std:: function
 * on a `float`, the fp64 variant calls the double overload.
 * ---------------------------------------------------------------------------------------- */
#if OF2D_REAL_IS_DOUBLE
#define R_FLOOR floor
#define R_SQRT sqrt
#define R_LOG2 log2
#define R_CEIL ceil
#define R_EXP exp
#else
#define R_FLOOR floorf
#define R_SQRT sqrtf
#define R_LOG2 log2f
#define R_CEIL ceilf
#define R_EXP expf
#endif

enum { REG_DIFFUSION = 0, REG_CURVATURE = 1, REG_ELASTIC = 2, REG_THIRION = 3, REG_DIFFEO = 4, REG_FLUID = 5 };

static char g_err[256];
static int g_status; /* sticky status raised inside loops (the reference throws) */

static void raise_status(int st, const char *msg) {
    if (g_status == OF2D_OK) {
        g_status = st;
        snprintf(g_err, sizeof(g_err), "%s", msg);
    }
}
#define DIVZERO() raise_status(OF2D_ERUNTIME, "Divide by zero exception")

int of2d_oracle_sizeof_real(void) { return (int)sizeof(real); }
const char *of2d_oracle_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------------------------
 * control-flow trace (what the reference prints through mexPrintf)
 * ---------------------------------------------------------------------------------------- */
typedef struct { double *a, *b; int n, cap; } trace_t;
static trace_t g_tr[3];
static int g_trace_on = 1;

static void trace_push(int which, double a, double b) {
    trace_t *t = &g_tr[which];
    if (!g_trace_on) return;
    if (t->n == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 256;
        t->a = (double *)realloc(t->a, sizeof(double) * (size_t)t->cap);
        t->b = (double *)realloc(t->b, sizeof(double) * (size_t)t->cap);
    }
    t->a[t->n] = a;
    t->b[t->n] = b;
    t->n++;
}
void of2d_trace_reset(void) { for (int w = 0; w < 3; w++) g_tr[w].n = 0; }
void of2d_trace_enable(int on) { g_trace_on = on != 0; }
int of2d_trace_count(int which) { return g_tr[which].n; }
void of2d_trace_get(int which, double *a, double *b, int cap) {
    int n = g_tr[which].n < cap ? g_tr[which].n : cap;
    memcpy(a, g_tr[which].a, sizeof(double) * (size_t)n);
    memcpy(b, g_tr[which].b, sizeof(double) * (size_t)n);
}

/* ------------------------------------------------------------------------------------------
 * small helpers
 * ---------------------------------------------------------------------------------------- */
static vec2 *vnew(size_t n) { return (vec2 *)calloc(n ? n : 1, sizeof(vec2)); }
static real *rnew(size_t n) { return (real *)calloc(n ? n : 1, sizeof(real)); }
static vec2 v2(real x, real y) { vec2 v; v.x = x; v.y = y; return v; }
static vec2 vadd(vec2 a, vec2 b) { return v2(a.x + b.x, a.y + b.y); }   /* coord2d.h:42-44 */
static vec2 vsub(vec2 a, vec2 b) { return v2(a.x - b.x, a.y - b.y); }   /* coord2d.h:63-65 */
static vec2 vmul(vec2 a, real s) { return v2(a.x * s, a.y * s); }       /* coord2d.h:84-86 */
static vec2 vdiv(vec2 a, real s) {                                      /* coord2d.h:95-100 */
    if (s == 0) { DIVZERO(); return v2(0, 0); }
    return v2(a.x / s, a.y / s);
}

/* gradients.h:9-19 / 22-32 on a scalar image */
static real partial_x_r(const real *f, unsigned idx, unsigned i, unsigned dimx) {
    if (i == 0) return f[idx + 1] - f[idx];
    if (i == dimx - 1) return f[idx] - f[idx - 1];
    return (f[idx + 1] - f[idx - 1]) / 2.0f;
}
static real partial_y_r(const real *f, unsigned idx, unsigned j, unsigned dimx, unsigned dimy) {
    if (j == 0) return f[idx + dimx] - f[idx];
    if (j == dimy - 1) return f[idx] - f[idx - dimx];
    return (f[idx + dimx] - f[idx - dimx]) / 2.0f;
}
/* the same templates instantiated on vector2d (used by Image::jacobian and the fluid increment) */
static vec2 partial_x_v(const vec2 *f, unsigned idx, unsigned i, unsigned dimx) {
    if (i == 0) return vsub(f[idx + 1], f[idx]);
    if (i == dimx - 1) return vsub(f[idx], f[idx - 1]);
    return vdiv(vsub(f[idx + 1], f[idx - 1]), 2.0f);
}
static vec2 partial_y_v(const vec2 *f, unsigned idx, unsigned j, unsigned dimx, unsigned dimy) {
    if (j == 0) return vsub(f[idx + dimx], f[idx]);
    if (j == dimy - 1) return vsub(f[idx], f[idx - dimx]);
    return vdiv(vsub(f[idx + dimx], f[idx - dimx]), 2.0f);
}

/* ------------------------------------------------------------------------------------------
 * Image / Motion primitives
 * ---------------------------------------------------------------------------------------- */

/* src/Image.cpp:15-29 */
static void set_image(real *dst, const double *src, size_t n) {
    for (size_t k = 0; k < n; k++) dst[k] = (real)src[k];
}

/* src/Motion.cpp:42-49: float accumulator, double addend (std::pow(float,int) promotes) */
static real motion_norm(const vec2 *u, unsigned n) {
    real norm = 0.0f;
    for (unsigned k = 0; k < n; k++) {
        double a = sqrt((double)u[k].x * (double)u[k].x + (double)u[k].y * (double)u[k].y);
        norm = (real)((double)norm + a);
    }
    return norm / n;
}

/* src/Motion.cpp:51-58: y is used twice, x is ignored */
static real motion_maxabs(const vec2 *u, unsigned n) {
    real maxabs = 0.0f;
    for (unsigned k = 0; k < n; k++) {
        real normsq = (real)((double)u[k].y * (double)u[k].y + (double)u[k].y * (double)u[k].y);
        if (maxabs < normsq) maxabs = normsq; /* std::max(a,b) = (a<b)?b:a */
    }
    return R_SQRT(maxabs);
}

/* src/Image.cpp:119-182 */
static void warp2d(real *img, const vec2 *mo, unsigned dimx, unsigned dimy) {
    const size_t n = (size_t)dimx * dimy;
    real *tmp = rnew(n);
    memcpy(tmp, img, n * sizeof(real));
    for (int i = 0; i < (int)dimx; i++) {
        for (int j = 0; j < (int)dimy; j++) {
            const int idx = i + j * (int)dimx;
            real px = i + mo[idx].x; int dx = (int)R_FLOOR(px); real fx = px - dx;
            real py = j + mo[idx].y; int dy = (int)R_FLOOR(py); real fy = py - dy;
            if (dx < 0 || dx >= (int)dimx || dy < 0 || dy >= (int)dimy) continue;
            const int idxO = dx + dy * (int)dimx;
            real val = tmp[idxO] * (1 - fx) * (1 - fy);
            real weight = (1 - fx) * (1 - fy);
            if (dx < (int)dimx - 1) { val += tmp[idxO + 1] * fx * (1 - fy); weight += fx * (1 - fy); }
            if (dy < (int)dimy - 1) { val += tmp[idxO + dimx] * (1 - fx) * fy; weight += (1 - fx) * fy; }
            if (dx < (int)dimx - 1 && dy < (int)dimy - 1) { val += tmp[idxO + 1 + dimx] * fx * fy; weight += fx * fy; }
            if (weight != 0) img[idx] = val / weight;
        }
    }
    free(tmp);
}

/* src/Motion.cpp:113-178: u <- v + u o (id + v); out-of-range source cell keeps the old u */
static void accumulate(vec2 *u, const vec2 *v, unsigned dimx, unsigned dimy) {
    const size_t n = (size_t)dimx * dimy;
    vec2 *old = vnew(n);
    memcpy(old, u, n * sizeof(vec2));
    for (unsigned i = 0; i < dimx; i++) {
        for (unsigned j = 0; j < dimy; j++) {
            const unsigned idx = i + j * dimx;
            real px = i + v[idx].x; int dx = (int)R_FLOOR(px); real fx = px - dx;
            real py = j + v[idx].y; int dy = (int)R_FLOOR(py); real fy = py - dy;
            if (dx < 0 || dx >= (int)dimx || dy < 0 || dy >= (int)dimy) continue;
            u[idx] = v[idx];
            const int idxO = dx + dy * (int)dimx;
            vec2 val = vmul(vmul(old[idxO], 1 - fx), 1 - fy);
            real weight = (1 - fx) * (1 - fy);
            if (dx < (int)dimx - 1) { val = vadd(val, vmul(vmul(old[idxO + 1], fx), 1 - fy)); weight += fx * (1 - fy); }
            if (dy < (int)dimy - 1) { val = vadd(val, vmul(vmul(old[idxO + dimx], 1 - fx), fy)); weight += (1 - fx) * fy; }
            if (dx < (int)dimx - 1 && dy < (int)dimy - 1) { val = vadd(val, vmul(vmul(old[idxO + 1 + dimx], fx), fy)); weight += fx * fy; }
            if (weight != 0) u[idx] = vadd(u[idx], vdiv(val, weight));
        }
    }
    free(old);
}

/* src/Kernel.cpp:45-73: float exp of a float argument, stored and normalised in double */
static void gaussian_kernel(double *k, unsigned w, real sigma) {
    const int cx = (int)((w - 1) / 2), cy = cx;
    double weight = 0;
    for (int i = 0; i < (int)w; i++) {
        for (int j = 0; j < (int)w; j++) {
            const unsigned idx = (unsigned)i + (unsigned)j * w;
            k[idx] = R_EXP(-((i - cx) * (i - cx) + (j - cy) * (j - cy)) / (2 * sigma * sigma));
            weight += k[idx];
        }
    }
    for (unsigned t = 0; t < w * w; t++) k[t] /= weight;
}

/* src/Field.tpp:210-269 instantiated on vector2d (Motion::convolute, src/Motion.cpp:279-282).
   The bounds test is on the linear index only, so x taps wrap into the neighbouring column. */
static void convolute(vec2 *u, const double *k, unsigned w, unsigned dimx, unsigned dimy) {
    const size_t n = (size_t)dimx * dimy;
    const int cx = (int)((w - 1) / 2), cy = cx;
    vec2 *tmp = vnew(n);
    memcpy(tmp, u, n * sizeof(vec2));
    for (int i = 0; i < (int)dimx; i++) {
        for (int j = 0; j < (int)dimy; j++) {
            const int idx = i + j * (int)dimx;
            vec2 val = v2(0, 0);
            double weight = 0.0f;
            for (int ii = -cx; ii <= cx; ii++) {
                for (int jj = -cy; jj <= cy; jj++) {
                    const long lin = (long)(i + ii) + (long)(j + jj) * (long)dimx;
                    if (lin < 0 || lin >= (long)n) continue;
                    const int ik = (ii + cx) + (jj + cy) * (int)w;
                    val = vadd(val, vmul(tmp[lin], (real)k[ik]));
                    weight += k[ik];
                }
            }
            if (weight != 0) u[idx] = vdiv(val, (real)weight);
        }
    }
    free(tmp);
}

/* src/Motion.cpp:253-277: scaling and squaring */
static int motion_exp(vec2 *u, unsigned dimx, unsigned dimy) {
    const size_t n = (size_t)dimx * dimy;
    const real ma = motion_maxabs(u, (unsigned)n);
    int nsquares;
    if (ma == 0) {
        nsquares = 0; /* log2(0) = -inf -> int cast is UB; x86 yields INT_MIN, clamped to 0 */
    } else {
        nsquares = (int)R_CEIL(1 + R_LOG2(ma));
        if (nsquares < 0) nsquares = 0;
    }
    if (nsquares == 0) return 0;
    const real scale = (real)pow(2, -nsquares);
    for (size_t t = 0; t < n; t++) u[t] = vmul(u[t], scale);
    vec2 *tmp = vnew(n);
    for (int s = 0; s < nsquares; s++) {
        memcpy(tmp, u, n * sizeof(vec2));
        accumulate(u, tmp, dimx, dimy);
    }
    free(tmp);
    return nsquares;
}

/* src/Image.cpp:189-218 and :96-104 */
static void jacobian(real *jac, const vec2 *u, unsigned dimx, unsigned dimy) {
    for (unsigned i = 0; i < dimx; i++) {
        for (unsigned j = 0; j < dimy; j++) {
            const unsigned idx = i + j * dimx;
            vec2 dudx = partial_x_v(u, idx, i, dimx);
            vec2 dudy = partial_y_v(u, idx, j, dimx, dimy);
            jac[idx] = (1.0f + dudx.x) * (1.0f + dudy.y) - dudx.y * dudy.x;
        }
    }
}
static real image_min(const real *f, size_t n) {
    real m = f[0];
    for (size_t k = 1; k < n; k++) if (f[k] < m) m = f[k];
    return m;
}

/* src/regularization/IterativeSolver.cpp:22-56 */
static void set_derivatives(vec2 *gradI, real *It, const real *Iref, const real *Imov, unsigned dimx, unsigned dimy) {
    for (unsigned i = 0; i < dimx; i++) {
        for (unsigned j = 0; j < dimy; j++) {
            const unsigned idx = i + j * dimx;
            gradI[idx] = v2(partial_x_r(Imov, idx, i, dimx), partial_y_r(Imov, idx, j, dimx, dimy));
        }
    }
    for (size_t k = 0; k < (size_t)dimx * dimy; k++) It[k] = Imov[k] - Iref[k];
}

/* src/Field.tpp:76-143 (box down-sampling) */
#define DEF_DOWNSAMPLE(NAME, T, ZERO, ADD, DIVP)                                                     \
    static int NAME(T *out, unsigned ox, unsigned oy, const T *in, unsigned ix, unsigned iy) {        \
        const unsigned sizein = ix * iy;                                                              \
        if (ox > ix || oy > iy) return OF2D_EINVAL;                                                   \
        if (ox == 0 || oy == 0) { DIVZERO(); return OF2D_ERUNTIME; }                                  \
        const unsigned fx = ix / ox, fy = iy / oy;                                                    \
        for (int i = 0; i < (int)ox; i++) {                                                           \
            for (int j = 0; j < (int)oy; j++) {                                                       \
                const unsigned idxout = (unsigned)i + (unsigned)j * ox;                               \
                const unsigned idxin = (unsigned)i * fx + (unsigned)j * fy * ix;                      \
                T val = ZERO;                                                                         \
                int p = 0;                                                                            \
                for (int ii = 0; ii < (int)fx; ii++) {                                                \
                    for (int jj = 0; jj < (int)fy; jj++) {                                            \
                        const unsigned q = idxin + (unsigned)ii + (unsigned)jj * ix;                  \
                        if (q >= sizein) continue;                                                    \
                        val = ADD(val, in[q]);                                                        \
                        p++;                                                                          \
                    }                                                                                 \
                }                                                                                     \
                if (p != 0) out[idxout] = DIVP(val, (real)p);                                         \
            }                                                                                         \
        }                                                                                             \
        return OF2D_OK;                                                                               \
    }
#define R_ADD(a, b) ((a) + (b))
#define R_DIV(a, b) ((a) / (b))
DEF_DOWNSAMPLE(downsample_r, real, 0.0f, R_ADD, R_DIV)
DEF_DOWNSAMPLE(downsample_v, vec2, v2(0.0f, 0.0f), vadd, vdiv)

/* src/Field.tpp:146-206 (bilinear up-sampling) */
#define DEF_UPSAMPLE(NAME, T, MUL, ADD, DIVW)                                                        \
    static int NAME(T *out, unsigned ox, unsigned oy, const T *in, unsigned ix, unsigned iy) {        \
        const unsigned sizein = ix * iy;                                                              \
        if (ox < ix || oy < iy) return OF2D_EINVAL;                                                   \
        for (unsigned i = 0; i < ox; i++) {                                                           \
            for (unsigned j = 0; j < oy; j++) {                                                       \
                const unsigned idx = i + j * ox;                                                      \
                real px = (real)i * ix / (real)ox; int dx = (int)R_FLOOR(px); real fx = px - dx;      \
                real py = (real)j * iy / (real)oy; int dy = (int)R_FLOOR(py); real fy = py - dy;      \
                const unsigned idxO = (unsigned)dx + (unsigned)dy * ix;                               \
                if (idxO >= sizein) continue;                                                         \
                T val = MUL(MUL(in[idxO], (1 - fx)), (1 - fy));                                       \
                real weight = (1 - fx) * (1 - fy);                                                    \
                if ((unsigned)dx < ix - 1) { val = ADD(val, MUL(MUL(in[idxO + 1], fx), (1 - fy))); weight += fx * (1 - fy); } \
                if ((unsigned)dy < iy - 1) { val = ADD(val, MUL(MUL(in[idxO + ix], (1 - fx)), fy)); weight += (1 - fx) * fy; } \
                if ((unsigned)dx < ix - 1 && (unsigned)dy < iy - 1) { val = ADD(val, MUL(MUL(in[idxO + 1 + ix], fx), fy)); weight += fx * fy; } \
                if (weight != 0) out[idx] = DIVW(val, weight);                                        \
            }                                                                                         \
        }                                                                                             \
        return OF2D_OK;                                                                               \
    }
#define R_MUL(a, b) ((a) * (b))
DEF_UPSAMPLE(upsample_r, real, R_MUL, R_ADD, R_DIV)
DEF_UPSAMPLE(upsample_v, vec2, vmul, vadd, vdiv)

/* src/Motion.cpp:61-111: resample, then rescale the displacement magnitudes by the dim ratio */
static int motion_resample(vec2 *out, unsigned ox, unsigned oy, const vec2 *in, unsigned ix, unsigned iy, int up) {
    int st = up ? upsample_v(out, ox, oy, in, ix, iy) : downsample_v(out, ox, oy, in, ix, iy);
    if (st != OF2D_OK) return st;
    const real rx = (real)ox / (real)ix, ry = (real)oy / (real)iy;
    for (size_t k = 0; k < (size_t)ox * oy; k++) { out[k].x *= rx; out[k].y *= ry; }
    return OF2D_OK;
}

/* ------------------------------------------------------------------------------------------
 * Logger (src/Logger.cpp:6-59)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    unsigned n;
    vec2 *prev, *diff;
    real *error;
    unsigned niter, iter;
    int verbose;
} logger_t;

static void logger_init(logger_t *L, unsigned n, unsigned niter, int verbose) {
    L->n = n; L->prev = vnew(n); L->diff = vnew(n);
    L->error = rnew(niter + 1); L->niter = niter; L->iter = 0; L->verbose = verbose;
}
static void logger_free(logger_t *L) { free(L->prev); free(L->diff); free(L->error); }
static void logger_update(logger_t *L, const vec2 *u) {
    for (unsigned k = 0; k < L->n; k++) L->diff[k] = vsub(u[k], L->prev[k]);
    const real prevnorm = motion_norm(L->prev, L->n);
    L->error[L->iter] = (prevnorm == 0 ? 0.0f : motion_norm(L->diff, L->n) / prevnorm);
    memcpy(L->prev, u, sizeof(vec2) * L->n);
    if (L->verbose) trace_push(0, (double)L->iter, (double)L->error[L->iter]);
    L->iter++;
}
static real logger_current(const logger_t *L) { return L->error[L->iter - 1]; }

/* ------------------------------------------------------------------------------------------
 * Solvers (src/regularization)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int reg;
    unsigned dimx, dimy, n;
    vec2 *gradI; real *It;             /* IterativeSolver.h:29-30 */
    vec2 *force;                       /* OpticalFlow.h:23 */
    vec2 *qdiff; real alpha;           /* Diffusion */
    real tau; double *rhs_x, *rhs_y, *eig; fftw_plan pf, pb; /* Curvature */
    real mu, lambda, omega;            /* Elastic / Fluid */
    vec2 *velocity, *increment; real timestep; /* Fluid (velocity is never reset) */
    real *Iwar; vec2 *corr; real sigma_i, sigma_x; double *k_diff, *k_fluid; unsigned kw; int accumulation; /* Demons */
} solver_t;

/* src/regularization/OpticalFlow/OpticalFlow.cpp:15-39 */
static void get_force(const solver_t *S, vec2 *f, const vec2 *u) {
    for (unsigned idx = 0; idx < S->n; idx++) {
        const vec2 dI = S->gradI[idx];
        f[idx] = vmul(dI, S->It[idx] + u[idx].x * dI.x + u[idx].y * dI.y);
    }
}

/* src/regularization/OpticalFlow/OpticalFlowCurvature.cpp:6-30 (PI truncated as in :4) */
#define REF_PI 3.14159265
static void curvature_eigenvalues(solver_t *S) {
    for (unsigned p = 0; p < S->dimx; p++) {
        for (unsigned q = 0; q < S->dimy; q++) {
            const unsigned idx = p * S->dimy + q;
            const double lap = -4 + 2 * cos(p * REF_PI / S->dimx) + 2 * cos(q * REF_PI / S->dimy);
            S->eig[idx] = 1.0f / (1.0f + S->tau * S->alpha * pow(lap, 2));
        }
    }
}

static int solver_init(solver_t *S, int reg, unsigned dimx, unsigned dimy, const real *p, int np) {
    memset(S, 0, sizeof(*S));
    S->reg = reg; S->dimx = dimx; S->dimy = dimy; S->n = dimx * dimy;
    S->gradI = vnew(S->n); S->It = rnew(S->n);
    switch (reg) {
        case REG_DIFFUSION: /* ImageRegistrationOpticalFlow.cpp:22-31 */
            S->force = vnew(S->n); S->qdiff = vnew(S->n); S->alpha = p[0];
            break;
        case REG_CURVATURE: /* :32-48, OpticalFlowCurvature.cpp:33-56 */
            S->force = vnew(S->n); S->alpha = p[0]; S->tau = np >= 2 ? p[1] : 1.0f;
            S->rhs_x = (double *)calloc(S->n, sizeof(double));
            S->rhs_y = (double *)calloc(S->n, sizeof(double));
            S->eig = (double *)calloc(S->n, sizeof(double));
            curvature_eigenvalues(S);
            S->pf = fftw_plan_r2r_2d((int)dimx, (int)dimy, S->rhs_x, S->rhs_x, FFTW_REDFT10, FFTW_REDFT10, FFTW_MEASURE);
            S->pb = fftw_plan_r2r_2d((int)dimx, (int)dimy, S->rhs_x, S->rhs_x, FFTW_REDFT01, FFTW_REDFT01, FFTW_MEASURE);
            break;
        case REG_ELASTIC: /* :49-66 */
            S->force = vnew(S->n); S->mu = p[0]; S->lambda = p[1]; S->omega = np >= 3 ? p[2] : 0.66f;
            break;
        case REG_FLUID: /* ImageRegistrationFluid.cpp:17-34, OpticalFlowFluid.h:10 (default 0.66 is a double literal) */
            S->force = vnew(S->n); S->mu = p[0]; S->lambda = p[1]; S->omega = np >= 3 ? p[2] : (real)0.66;
            S->velocity = vnew(S->n); S->increment = vnew(S->n);
            break;
        case REG_THIRION: /* ImageRegistrationDemons.cpp:21-39, Demons.cpp:4-24 */
        case REG_DIFFEO:
            S->Iwar = rnew(S->n); S->corr = vnew(S->n);
            S->sigma_i = p[0]; S->sigma_x = p[1];
            S->kw = (unsigned)p[4];
            S->k_diff = (double *)calloc((size_t)S->kw * S->kw + 1, sizeof(double));
            S->k_fluid = (double *)calloc((size_t)S->kw * S->kw + 1, sizeof(double));
            gaussian_kernel(S->k_diff, S->kw, p[2]);
            gaussian_kernel(S->k_fluid, S->kw, p[3]);
            S->accumulation = (reg == REG_THIRION) ? (int)p[5] : 0;
            break;
        default:
            return OF2D_EINVAL;
    }
    return OF2D_OK;
}

static void solver_free(solver_t *S) {
    free(S->gradI); free(S->It); free(S->force); free(S->qdiff);
    free(S->rhs_x); free(S->rhs_y); free(S->eig);
    if (S->pf) fftw_destroy_plan(S->pf);
    if (S->pb) fftw_destroy_plan(S->pb);
    free(S->velocity); free(S->increment); free(S->Iwar); free(S->corr); free(S->k_diff); free(S->k_fluid);
}

/* OpticalFlowDiffusion.cpp:19-84 (Horn-Schunck, Jacobi) */
static void update_diffusion(solver_t *S, vec2 *u) {
    const unsigned dimx = S->dimx, dimy = S->dimy;
    for (unsigned i = 0; i < dimx; i++) {
        for (unsigned j = 0; j < dimy; j++) {
            const unsigned idx = i + j * dimx;
            if (i == 0 || i == dimx - 1 || j == 0 || j == dimy - 1) {
                S->qdiff[idx] = v2(0.0f, 0.0f);
            } else { /* gradients.h:72-80 */
                S->qdiff[idx] = vdiv(vadd(vadd(vadd(u[idx - 1], u[idx + 1]), u[idx - dimx]), u[idx + dimx]), 4.0f);
            }
        }
    }
    get_force(S, S->force, S->qdiff);
    const real alphasq = S->alpha * S->alpha;
    for (unsigned idx = 0; idx < S->n; idx++) {
        const vec2 dI = S->gradI[idx];
        u[idx] = vsub(S->qdiff[idx], vdiv(S->force[idx], alphasq + dI.x * dI.x + dI.y * dI.y));
    }
}

/* OpticalFlowCurvature.cpp:70-167 */
static void update_curvature(solver_t *S, vec2 *u) {
    const unsigned dimx = S->dimx, dimy = S->dimy;
    get_force(S, S->force, u);
    for (unsigned i = 0; i < dimx; i++) {
        for (unsigned j = 0; j < dimy; j++) {
            const unsigned rm = i * dimy + j, cm = i + j * dimx;
            S->rhs_x[rm] = u[cm].x - S->tau * S->force[cm].x;
            S->rhs_y[rm] = u[cm].y - S->tau * S->force[cm].y;
        }
    }
    fftw_execute_r2r(S->pf, S->rhs_x, S->rhs_x);
    fftw_execute_r2r(S->pf, S->rhs_y, S->rhs_y);
    for (unsigned k = 0; k < S->n; k++) { S->rhs_x[k] *= S->eig[k]; S->rhs_y[k] *= S->eig[k]; }
    fftw_execute_r2r(S->pb, S->rhs_x, S->rhs_x);
    fftw_execute_r2r(S->pb, S->rhs_y, S->rhs_y);
    for (unsigned i = 0; i < dimx; i++) {
        for (unsigned j = 0; j < dimy; j++) {
            const unsigned rm = i * dimy + j, cm = i + j * dimx;
            u[cm] = vdiv(v2((real)S->rhs_x[rm], (real)S->rhs_y[rm]), 4.0f * S->n);
        }
    }
}

/* OpticalFlowElastic.cpp:21-55 == OpticalFlowFluid.cpp:7-41: one in-place lexicographic SOR sweep,
   x (i) outer and y (j) inner; the y equation reuses the x[i+-1].y neighbours in the (mu+lambda)
   term exactly as the reference writes it. */
static void sor_sweep(const solver_t *S, vec2 *x) {
    const unsigned dimx = S->dimx, dimy = S->dimy;
    const real omega = S->omega, mu = S->mu, lambda = S->lambda;
    const vec2 *b = S->force;
    const int sx = 1, sy = (int)dimx;
    for (unsigned i = 1; i + 1 < dimx; i++) {
        for (unsigned j = 1; j + 1 < dimy; j++) {
            const int idx = (int)(i + j * dimx);
            x[idx].x = (1.0f - omega) * x[idx].x + omega / (-6 * mu - 2 * lambda) * (b[idx].x -
                mu * (x[idx + sx].x + x[idx - sx].x + x[idx + sy].x + x[idx - sy].x) -
                (mu + lambda) * (x[idx + sx].x + x[idx - sx].x + 0.25f * (x[idx + sx + sy].y - x[idx - sx + sy].y - x[idx + sx - sy].y + x[idx - sx - sy].y)));
            x[idx].y = (1.0f - omega) * x[idx].y + omega / (-6 * mu - 2 * lambda) * (b[idx].y -
                mu * (x[idx + sx].y + x[idx - sx].y + x[idx + sy].y + x[idx - sy].y) -
                (mu + lambda) * (x[idx + sx].y + x[idx - sx].y + 0.25f * (x[idx + sx + sy].x - x[idx - sx + sy].x - x[idx + sx - sy].x + x[idx - sx - sy].x)));
        }
    }
}

/* OpticalFlowElastic.cpp:13-19 */
static void update_elastic(solver_t *S, vec2 *u) {
    get_force(S, S->force, u);
    sor_sweep(S, u);
}

/* OpticalFlowFluid.cpp:60-140 */
static void update_fluid(solver_t *S, vec2 *u) {
    const unsigned dimx = S->dimx, dimy = S->dimy;
    get_force(S, S->force, u);
    sor_sweep(S, S->velocity);
    for (unsigned i = 0; i < dimx; i++) {
        for (unsigned j = 0; j < dimy; j++) {
            const unsigned idx = i + j * dimx;
            const vec2 v = S->velocity[idx];
            const vec2 dudx = partial_x_v(u, idx, i, dimx);
            const vec2 dudy = partial_y_v(u, idx, j, dimx, dimy);
            S->increment[idx] = vsub(vsub(v, vmul(dudx, v.x)), vmul(dudy, v.y));
        }
    }
    const real dumax = 0.65f;
    const real ma = motion_maxabs(S->increment, S->n);
    S->timestep = dumax / ma;
    trace_push(2, (double)ma, (double)S->timestep);
    if (S->timestep >= 65.0f) return;
    for (unsigned idx = 0; idx < S->n; idx++) u[idx] = vadd(u[idx], vmul(S->increment[idx], S->timestep));
}

/* Demons.cpp:34-63 */
static void demons_force(solver_t *S) {
    const real sigma_xsq = S->sigma_x * S->sigma_x;
    const real sigma_isq = S->sigma_i * S->sigma_i;
    for (unsigned idx = 0; idx < S->n; idx++) {
        const vec2 dI = S->gradI[idx];
        const real It = S->It[idx];
        S->corr[idx] = vmul(vdiv(vmul(dI, It), dI.x * dI.x + dI.y * dI.y + It * It * sigma_isq / sigma_xsq), -1);
    }
}

/* DemonsThirions.cpp:18-42 and DemonsDiffeomorphic.cpp:15-35 */
static void update_demons(solver_t *S, vec2 *u, const real *Iref, const real *Imov) {
    memcpy(S->Iwar, Imov, sizeof(real) * S->n);
    warp2d(S->Iwar, u, S->dimx, S->dimy);
    set_derivatives(S->gradI, S->It, Iref, S->Iwar, S->dimx, S->dimy);
    demons_force(S);
    convolute(S->corr, S->k_fluid, S->kw, S->dimx, S->dimy);
    if (S->reg == REG_DIFFEO) {
        motion_exp(S->corr, S->dimx, S->dimy);
        accumulate(u, S->corr, S->dimx, S->dimy);
    } else if (S->accumulation == 0) {
        accumulate(u, S->corr, S->dimx, S->dimy);
    } else if (S->accumulation == 1) {
        for (unsigned k = 0; k < S->n; k++) u[k] = vadd(u[k], S->corr[k]);
    }
    convolute(u, S->k_diff, S->kw, S->dimx, S->dimy);
}

static void solver_update(solver_t *S, vec2 *u, const real *Iref, const real *Imov) {
    switch (S->reg) {
        case REG_DIFFUSION: update_diffusion(S, u); break;
        case REG_CURVATURE: update_curvature(S, u); break;
        case REG_ELASTIC: update_elastic(S, u); break;
        case REG_FLUID: update_fluid(S, u); break;
        default: update_demons(S, u, Iref, Imov); break;
    }
}

/* ------------------------------------------------------------------------------------------
 * Drivers (src/ImageRegistration*.cpp)
 * ---------------------------------------------------------------------------------------- */
static int valid_params(int reg, int np) {
    switch (reg) {
        case REG_DIFFUSION: return np == 1;                 /* ImageRegistrationOpticalFlow.cpp:8-12 */
        case REG_CURVATURE: return np >= 1 && np <= 2;
        case REG_ELASTIC: return np >= 2 && np <= 3;
        case REG_THIRION: return np == 6;                   /* ImageRegistrationDemons.cpp:7-10 */
        case REG_DIFFEO: return np == 5;
        case REG_FLUID: return np >= 2 && np <= 3;          /* ImageRegistrationFluid.cpp:5-7 */
    }
    return 0;
}

/* ImageRegistrationOpticalFlow.cpp:97-151, ImageRegistrationDemons.cpp:86-137, ImageRegistrationFluid.cpp:67-142 */
static void estimate_level(solver_t *S, vec2 *motion, const real *Iref, const real *Imov, int niter, int nrefine, int verbose) {
    const unsigned n = S->n, dimx = S->dimx, dimy = S->dimy;
    real *Iaux = rnew(n);
    real *jac = rnew(n);
    vec2 *est = vnew(n);
    const int demons = (S->reg == REG_THIRION || S->reg == REG_DIFFEO);
    for (int refine = 0; refine < nrefine; refine++) {
        memcpy(Iaux, Imov, sizeof(real) * n);
        warp2d(Iaux, motion, dimx, dimy);
        logger_t log;
        logger_init(&log, n, (unsigned)niter, verbose);
        if (!demons) set_derivatives(S->gradI, S->It, Iref, Iaux, dimx, dimy);
        for (int iter = 0; iter < niter; iter++) {
            solver_update(S, est, Iref, Iaux);
            if (g_status != OF2D_OK) break; /* the reference would have thrown out of get_update */
            logger_update(&log, est);
            if (logger_current(&log) < 0.001f && iter > 1) break;
            if (S->reg == REG_FLUID) {
                jacobian(jac, est, dimx, dimy);
                const real mj = image_min(jac, n);
                if (mj < 0.5) {
                    trace_push(1, (double)iter, (double)mj);
                    accumulate(motion, est, dimx, dimy);
                    memset(est, 0, sizeof(vec2) * n);
                    memcpy(Iaux, Imov, sizeof(real) * n);
                    warp2d(Iaux, motion, dimx, dimy);
                    set_derivatives(S->gradI, S->It, Iref, Iaux, dimx, dimy);
                }
            }
        }
        logger_free(&log);
        if (g_status != OF2D_OK) break;
        accumulate(motion, est, dimx, dimy);
        memset(est, 0, sizeof(vec2) * n);
    }
    free(Iaux); free(jac); free(est);
}

/* WrapperOpticalFlow2d.cpp:18-155 + ImageRegistration.cpp:49-156, replaying test_opticalflow2d.m:42-59 */
int of2d_oracle_mex_register(int dimx, int dimy, int nscales, const double *niter_d, int nrefine, int reg,
                             const double *regparams_d, int nparams, int verbose, const double *Iref_d,
                             const double *Imov_d, double *motion_out, double *warped_out) {
    g_status = OF2D_OK; g_err[0] = 0;
    if (reg < 0 || reg > 5) { raise_status(OF2D_ERUNTIME, "mexErrMsgTxt: Error: invalid regularisation given\n"); return g_status; }
    if (!valid_params(reg, nparams)) {
        raise_status(OF2D_EINVAL, "Invalid number of regularisation parameters for given regularisation method.\n");
        return g_status;
    }
    const int L = nscales + 1;
    int *niter = (int *)calloc((size_t)L, sizeof(int));
    for (int s = 0; s < L; s++) niter[s] = (int)niter_d[s];
    real params[8] = {0};
    for (int p = 0; p < nparams && p < 8; p++) params[p] = (real)regparams_d[p];

    unsigned *dx = (unsigned *)calloc((size_t)L, sizeof(unsigned));
    unsigned *dy = (unsigned *)calloc((size_t)L, sizeof(unsigned));
    real **Iref = (real **)calloc((size_t)L, sizeof(real *));
    real **Imov = (real **)calloc((size_t)L, sizeof(real *));
    vec2 **motion = (vec2 **)calloc((size_t)L, sizeof(vec2 *));
    solver_t *solver = (solver_t *)calloc((size_t)L, sizeof(solver_t));
    for (int s = nscales; s >= 0; s--) { /* ImageRegistration.cpp:56-61 */
        const real scale = (real)pow(2, s);
        dx[s] = (unsigned)((unsigned)dimx / scale);
        dy[s] = (unsigned)((unsigned)dimy / scale);
        const size_t n = (size_t)dx[s] * dy[s];
        Iref[s] = rnew(n); Imov[s] = rnew(n); motion[s] = vnew(n);
        solver_init(&solver[s], reg, dx[s], dy[s], params, nparams);
    }
    /* register call: set images (+ pyramid), ImageRegistration.cpp:103-121 */
    set_image(Iref[0], Iref_d, (size_t)dimx * dimy);
    set_image(Imov[0], Imov_d, (size_t)dimx * dimy);
    for (int s = nscales; s >= 1; s--) {
        int st = downsample_r(Iref[s], dx[s], dy[s], Iref[0], dx[0], dy[0]);
        if (st == OF2D_OK) st = downsample_r(Imov[s], dx[s], dy[s], Imov[0], dx[0], dy[0]);
        if (st != OF2D_OK) raise_status(OF2D_ERUNTIME, "mexErrMsgTxt: Error in Image::downSample");
    }
    /* estimate_motion, ImageRegistration.cpp:133-156 */
    for (int s = nscales; s >= 0 && g_status == OF2D_OK; s--) {
        if (s > 0 && s < nscales) motion_resample(motion[s], dx[s], dy[s], motion[0], dx[0], dy[0], 0);
        estimate_level(&solver[s], motion[s], Iref[s], Imov[s], niter[s], nrefine, verbose);
        if (s > 0 && g_status == OF2D_OK) motion_resample(motion[0], dx[0], dy[0], motion[s], dx[s], dy[s], 1);
    }
    if (g_status == OF2D_OK) {
        const size_t n = (size_t)dimx * dimy;
        if (motion_out) { /* Motion.cpp:23-39: planar double */
            for (size_t k = 0; k < n; k++) { motion_out[k] = (double)motion[0][k].x; motion_out[k + n] = (double)motion[0][k].y; }
        }
        if (warped_out) { /* WrapperOpticalFlow2d.cpp:120-137 */
            real *w = rnew(n);
            set_image(w, Imov_d, n);
            warp2d(w, motion[0], (unsigned)dimx, (unsigned)dimy);
            for (size_t k = 0; k < n; k++) warped_out[k] = (double)w[k];
            free(w);
        }
    }
    for (int s = 0; s < L; s++) { free(Iref[s]); free(Imov[s]); free(motion[s]); solver_free(&solver[s]); }
    free(Iref); free(Imov); free(motion); free(solver); free(dx); free(dy); free(niter);
    return g_status;
}

/* WrapperOpticalFlow2d.cpp:149-151: any call shape that is not one of the five (with no live singleton) */
int of2d_oracle_mex_badcall(int nlhs, int nrhs) {
    g_status = OF2D_OK; g_err[0] = 0;
    if (nlhs == 0 && nrhs == 8) {
        /* an init call with dummy zero args: reg = 0 (Diffusion) with nparams = 0 -> invalid_argument */
        raise_status(OF2D_EINVAL, "Invalid number of regularisation parameters for given regularisation method.\n");
    } else {
        raise_status(OF2D_ERUNTIME, "mexErrMsgTxt: Error: invalid number of input and output variables gives.\n");
    }
    return g_status;
}

/* ------------------------------------------------------------------------------------------
 * primitive entry points (same shapes as oracle/ref_shim.cpp)
 * ---------------------------------------------------------------------------------------- */
#define BEGIN() do { g_status = OF2D_OK; g_err[0] = 0; } while (0)

int of2d_oracle_set_image(int dimx, int dimy, const double *in, real *out) {
    BEGIN(); set_image(out, in, (size_t)dimx * dimy); return g_status;
}
int of2d_oracle_copy_motion_to_input(int dimx, int dimy, const real *u, double *out) {
    BEGIN();
    const size_t n = (size_t)dimx * dimy;
    const vec2 *m = (const vec2 *)u;
    for (size_t k = 0; k < n; k++) { out[k] = (double)m[k].x; out[k + n] = (double)m[k].y; }
    return g_status;
}
int of2d_oracle_warp2d(int dimx, int dimy, real *img, const real *u) {
    BEGIN(); warp2d(img, (const vec2 *)u, (unsigned)dimx, (unsigned)dimy); return g_status;
}
int of2d_oracle_accumulate(int dimx, int dimy, real *u, const real *v) {
    BEGIN(); accumulate((vec2 *)u, (const vec2 *)v, (unsigned)dimx, (unsigned)dimy); return g_status;
}
int of2d_oracle_gaussian_kernel(int w, real sigma, double *out) {
    BEGIN(); gaussian_kernel(out, (unsigned)w, sigma); return g_status;
}
int of2d_oracle_convolute_motion(int dimx, int dimy, real *u, int w, real sigma) {
    BEGIN();
    double *k = (double *)calloc((size_t)w * w + 1, sizeof(double));
    gaussian_kernel(k, (unsigned)w, sigma);
    convolute((vec2 *)u, k, (unsigned)w, (unsigned)dimx, (unsigned)dimy);
    free(k);
    return g_status;
}
/* ---- the rest of the public Image / Motion / Kernel surface ---- */
/* src/Image.cpp:78-104: sequential accumulation in `real`; max starts from 0, min from the first element */
int of2d_oracle_image_stats(int dimx, int dimy, const real *img, real *sum, real *mx, real *mn) {
    BEGIN();
    const size_t n = (size_t)dimx * dimy;
    real s = 0.0f, hi = 0.0f;
    for (size_t k = 0; k < n; k++) s += img[k];
    for (size_t k = 0; k < n; k++) if (img[k] > hi) hi = img[k];
    *sum = s; *mx = hi; *mn = image_min(img, n);
    return g_status;
}
/* src/Image.cpp:107-116 */
int of2d_oracle_image_normalize(int dimx, int dimy, real *img) {
    BEGIN();
    const size_t n = (size_t)dimx * dimy;
    real hi = 0.0f;
    for (size_t k = 0; k < n; k++) if (img[k] > hi) hi = img[k];
    const real lo = image_min(img, n);
    for (size_t k = 0; k < n; k++) img[k] = (img[k] - lo) / (hi - lo);
    return g_status;
}
/* src/Motion.cpp:181-251; step = (1, dimx).  kind 0: Neumann, 1: Dirichlet.  The assignments are sequential; the
   (dimx-1, 0) corner of the Neumann variant reads u[(dimy-2) + dimx] (src/Motion.cpp:213 uses the y extent for an x index). */
int of2d_oracle_boundary_conditions(int dimx, int dimy, int kind, real *uu) {
    BEGIN();
    vec2 *u = (vec2 *)uu;
    const unsigned nx = (unsigned)dimx, ny = (unsigned)dimy;
    const vec2 zero = v2(0, 0);
    for (unsigned i = 1; i + 1 < nx; i++) {
        unsigned idx = i;
        u[idx] = kind ? zero : u[idx + nx];
        idx += (ny - 1) * nx;
        u[idx] = kind ? zero : u[idx - nx];
    }
    for (unsigned j = 1; j + 1 < ny; j++) {
        unsigned idx = j * nx;
        u[idx] = kind ? zero : u[idx + 1];
        idx += nx - 1;
        u[idx] = kind ? zero : u[idx - 1];
    }
    u[0] = kind ? zero : u[1 + nx];
    u[(ny - 1) * nx] = kind ? zero : u[1 + (ny - 2) * nx];
    u[nx - 1] = kind ? zero : u[(ny - 2) + nx];
    u[(nx - 1) + (ny - 1) * nx] = kind ? zero : u[(nx - 2) + (ny - 2) * nx];
    return g_status;
}
/* src/Kernel.cpp:75-82: 1.0f / (float)size, stored as double */
int of2d_oracle_average_kernel(int w, double *out) {
    BEGIN();
    const unsigned n = (unsigned)w * (unsigned)w;
    for (unsigned t = 0; t < n; t++) out[t] = 1.0f / (real)n;
    return g_status;
}
/* src/Field.tpp:210-269 instantiated on float (Image::convolute, src/Image.cpp:184-187).  The reference leaves the accumulator
   `val` uninitialised there (src/Field.tpp:240, undefined behaviour): this restatement takes the evident intent, val = 0. */
int of2d_oracle_convolute_image(int dimx, int dimy, real *img, int w, real sigma) {
    BEGIN();
    const size_t n = (size_t)dimx * dimy;
    double *k = (double *)calloc((size_t)w * w + 1, sizeof(double));
    if (sigma > 0) gaussian_kernel(k, (unsigned)w, sigma); else of2d_oracle_average_kernel(w, k);
    const int cx = (w - 1) / 2, cy = cx;
    real *tmp = (real *)malloc(n * sizeof(real));
    memcpy(tmp, img, n * sizeof(real));
    for (int i = 0; i < dimx; i++) {
        for (int j = 0; j < dimy; j++) {
            const int idx = i + j * dimx;
            real val = 0;
            double weight = 0.0f;
            for (int ii = -cx; ii <= cx; ii++) {
                for (int jj = -cy; jj <= cy; jj++) {
                    const long lin = (long)(i + ii) + (long)(j + jj) * (long)dimx;
                    if (lin < 0 || lin >= (long)n) continue;
                    const int ik = (ii + cx) + (jj + cy) * w;
                    val += tmp[lin] * k[ik];     /* float * double: promoted, as in the reference's `fieldtmp[...] * k[...]` */
                    weight += k[ik];
                }
            }
            if (weight != 0) img[idx] = val / weight;
        }
    }
    free(tmp); free(k);
    return g_status;
}
int of2d_oracle_exp(int dimx, int dimy, real *u) {
    BEGIN(); motion_exp((vec2 *)u, (unsigned)dimx, (unsigned)dimy); return g_status;
}
int of2d_oracle_norm_maxabs(int dimx, int dimy, const real *u, real *norm, real *maxabs) {
    BEGIN();
    *norm = motion_norm((const vec2 *)u, (unsigned)(dimx * dimy));
    *maxabs = motion_maxabs((const vec2 *)u, (unsigned)(dimx * dimy));
    return g_status;
}
int of2d_oracle_jacobian(int dimx, int dimy, const real *u, real *jac, real *minjac) {
    BEGIN();
    jacobian(jac, (const vec2 *)u, (unsigned)dimx, (unsigned)dimy);
    *minjac = image_min(jac, (size_t)dimx * dimy);
    return g_status;
}
int of2d_oracle_derivatives(int dimx, int dimy, const real *Iref, const real *Imov, real *grad, real *It) {
    BEGIN(); set_derivatives((vec2 *)grad, It, Iref, Imov, (unsigned)dimx, (unsigned)dimy); return g_status;
}
int of2d_oracle_image_resample(int inx, int iny, const real *in, int outx, int outy, real *out, int up) {
    BEGIN();
    int st = up ? upsample_r(out, (unsigned)outx, (unsigned)outy, in, (unsigned)inx, (unsigned)iny)
                : downsample_r(out, (unsigned)outx, (unsigned)outy, in, (unsigned)inx, (unsigned)iny);
    /* Image::upSample / downSample turn invalid_argument into mexErrMsgTxt (Image.cpp:53-75) */
    if (st == OF2D_EINVAL) raise_status(OF2D_ERUNTIME, "mexErrMsgTxt: Error in Image::resample: input has to have same dimensions as target");
    return g_status;
}
int of2d_oracle_motion_resample(int inx, int iny, const real *in, int outx, int outy, real *out, int up) {
    BEGIN();
    int st = motion_resample((vec2 *)out, (unsigned)outx, (unsigned)outy, (const vec2 *)in, (unsigned)inx, (unsigned)iny, up);
    if (st == OF2D_EINVAL) raise_status(OF2D_ERUNTIME, "mexErrMsgTxt: Error in Motion::resample: input has to have same dimensions as target");
    return g_status;
}
int of2d_oracle_logger(int dimx, int dimy, const real *useq, int nseq, real *err) {
    BEGIN();
    const unsigned n = (unsigned)(dimx * dimy);
    logger_t L;
    logger_init(&L, n, (unsigned)nseq, 0);
    for (int s = 0; s < nseq; s++) {
        logger_update(&L, (const vec2 *)useq + (size_t)s * n);
        err[s] = logger_current(&L);
    }
    logger_free(&L);
    return g_status;
}
int of2d_oracle_solver_steps(int reg, const real *params, int nparams, int dimx, int dimy, const real *Iref,
                             const real *Imov, real *u, int nsteps) {
    BEGIN();
    solver_t S;
    if (solver_init(&S, reg, (unsigned)dimx, (unsigned)dimy, params, nparams) != OF2D_OK) {
        raise_status(OF2D_EINVAL, "unknown regularisation");
        return g_status;
    }
    const int demons = (reg == REG_THIRION || reg == REG_DIFFEO);
    if (!demons) set_derivatives(S.gradI, S.It, Iref, Imov, (unsigned)dimx, (unsigned)dimy);
    for (int k = 0; k < nsteps && g_status == OF2D_OK; k++) solver_update(&S, (vec2 *)u, Iref, Imov);
    solver_free(&S);
    return g_status;
}
