/*
 * dct_standin.c -- CPU stand-in for the three fftw3 entry points the reference
 * calls.  TEST INFRASTRUCTURE ONLY (see oracle/README.md): it is linked into
 * the checker libraries (the .so files under oracle/_ref/ and oracle/) and is
 * never loaded by the product path.
 *
 * Third-party dependency being restated: fftw3 (un-vendored, version not
 * pinned by the reference; `-lfftw3` in compile_mex_function.m:30-32).
 * Call sites in the reference:
 *   src/regularization/OpticalFlow/OpticalFlowCurvature.cpp:52-55  plan_r2r_2d(nx, ny, buf, buf, K, K, FFTW_MEASURE)
 *   src/regularization/OpticalFlow/OpticalFlowCurvature.cpp:152-160 execute_r2r
 *   src/regularization/OpticalFlow/OpticalFlowCurvature.cpp:64-67   destroy_plan
 * Published definitions (FFTW manual, "1d Real-even DFTs (DCTs)"), unnormalised:
 *   REDFT10 (DCT-II):  Y_k = 2 sum_{j=0}^{n-1} X_j cos(pi (j+1/2) k / n)
 *   REDFT01 (DCT-III): Y_k = X_0 + 2 sum_{j=1}^{n-1} X_j cos(pi j (k+1/2) / n)
 * A 2-D r2r plan applies the 1-D transform along each dimension of the
 * row-major n0 x n1 array.  tests/test_oracle_dct.py pins this file against
 * scipy.fft.dct(type=2|3, norm=None), which implements the same definitions.
 *
 * Algorithm: power-of-two lengths use Makhoul's reordering with one complex
 * FFT shared by two real lines (O(n log n)); any other length falls back to
 * the direct O(n^2) sum with a cosine table.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "stubs/fftw3.h"

typedef struct {
    int n;
    int pow2;
    int log2n;
    double *tw_re, *tw_im;   /* e^{-2 pi i k / n}, k < n/2 */
    double *q_re, *q_im;     /* e^{-i pi k / (2n)}, k < n */
    int *bitrev;
    double *costab;          /* cos(pi m / (2n)), m < 4n, only when !pow2 */
} dct1d;

struct of2d_dct_plan {
    int n0, n1;
    fftw_r2r_kind kind0, kind1;
    dct1d d0, d1;
    double *za_re, *za_im;   /* scratch of length max(n0,n1) */
    double *col_a, *col_b;   /* column gather buffers, length n0 */
    double *tmp;             /* length max(n0,n1) */
};

static const double OF2D_PI = 3.14159265358979323846264338327950288;

static void dct1d_init(dct1d *d, int n) {
    memset(d, 0, sizeof(*d));
    d->n = n;
    d->pow2 = (n >= 2) && ((n & (n - 1)) == 0);
    if (d->pow2) {
        int l = 0;
        while ((1 << l) < n) l++;
        d->log2n = l;
        d->tw_re = (double *)malloc(sizeof(double) * (size_t)(n / 2 + 1));
        d->tw_im = (double *)malloc(sizeof(double) * (size_t)(n / 2 + 1));
        for (int k = 0; k < n / 2; k++) {
            d->tw_re[k] = cos(-2.0 * OF2D_PI * k / n);
            d->tw_im[k] = sin(-2.0 * OF2D_PI * k / n);
        }
        d->q_re = (double *)malloc(sizeof(double) * (size_t)n);
        d->q_im = (double *)malloc(sizeof(double) * (size_t)n);
        for (int k = 0; k < n; k++) {
            d->q_re[k] = cos(-OF2D_PI * k / (2.0 * n));
            d->q_im[k] = sin(-OF2D_PI * k / (2.0 * n));
        }
        d->bitrev = (int *)malloc(sizeof(int) * (size_t)n);
        for (int i = 0; i < n; i++) {
            int r = 0;
            for (int b = 0; b < l; b++)
                if (i & (1 << b)) r |= 1 << (l - 1 - b);
            d->bitrev[i] = r;
        }
    } else {
        d->costab = (double *)malloc(sizeof(double) * (size_t)(4 * n));
        for (int m = 0; m < 4 * n; m++) d->costab[m] = cos(OF2D_PI * m / (2.0 * n));
    }
}

static void dct1d_free(dct1d *d) {
    free(d->tw_re); free(d->tw_im); free(d->q_re); free(d->q_im);
    free(d->bitrev); free(d->costab);
}

/* in-place radix-2 DIT complex FFT; sign = -1 forward, +1 inverse (unscaled) */
static void cfft(const dct1d *d, double *re, double *im, int sign) {
    const int n = d->n;
    for (int i = 0; i < n; i++) {
        int r = d->bitrev[i];
        if (r > i) {
            double t = re[i]; re[i] = re[r]; re[r] = t;
            t = im[i]; im[i] = im[r]; im[r] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1;
        const int stride = n / len;
        for (int base = 0; base < n; base += len) {
            for (int k = 0; k < half; k++) {
                const double wr = d->tw_re[k * stride];
                const double wi = (sign < 0) ? d->tw_im[k * stride] : -d->tw_im[k * stride];
                const int a = base + k, b = a + half;
                const double xr = re[b] * wr - im[b] * wi;
                const double xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr;        im[a] += xi;
            }
        }
    }
}

/* DCT-II of two real lines a, b (stride s) in place; b may be NULL */
static void dct2_pair(const dct1d *d, double *a, double *b, int s, double *zr, double *zi, double *tmp) {
    const int n = d->n;
    if (!d->pow2) {
        for (int line = 0; line < 2; line++) {
            double *x = line ? b : a;
            if (!x) continue;
            for (int k = 0; k < n; k++) {
                double acc = 0.0;
                for (int j = 0; j < n; j++) acc += x[(size_t)j * s] * d->costab[((2 * j + 1) * (long)k) % (4 * n)];
                tmp[k] = 2.0 * acc;
            }
            for (int k = 0; k < n; k++) x[(size_t)k * s] = tmp[k];
        }
        return;
    }
    for (int m = 0; m < n / 2; m++) {
        zr[m] = a[(size_t)(2 * m) * s];
        zr[n - 1 - m] = a[(size_t)(2 * m + 1) * s];
        zi[m] = b ? b[(size_t)(2 * m) * s] : 0.0;
        zi[n - 1 - m] = b ? b[(size_t)(2 * m + 1) * s] : 0.0;
    }
    cfft(d, zr, zi, -1);
    for (int k = 0; k < n; k++) {
        const int nk = (n - k) & (n - 1);
        /* V_a = (Z_k + conj Z_{n-k})/2,  V_b = (Z_k - conj Z_{n-k})/(2i) */
        const double var = 0.5 * (zr[k] + zr[nk]), vai = 0.5 * (zi[k] - zi[nk]);
        const double vbr = 0.5 * (zi[k] + zi[nk]), vbi = -0.5 * (zr[k] - zr[nk]);
        a[(size_t)k * s] = 2.0 * (var * d->q_re[k] - vai * d->q_im[k]);
        if (b) b[(size_t)k * s] = 2.0 * (vbr * d->q_re[k] - vbi * d->q_im[k]);
    }
}

/* DCT-III of two real lines a, b (stride s) in place; b may be NULL */
static void dct3_pair(const dct1d *d, double *a, double *b, int s, double *zr, double *zi, double *tmp) {
    const int n = d->n;
    if (!d->pow2) {
        for (int line = 0; line < 2; line++) {
            double *x = line ? b : a;
            if (!x) continue;
            for (int k = 0; k < n; k++) {
                double acc = 0.0;
                for (int j = 1; j < n; j++) acc += x[(size_t)j * s] * d->costab[((long)j * (2 * k + 1)) % (4 * n)];
                tmp[k] = x[0] + 2.0 * acc;
            }
            for (int k = 0; k < n; k++) x[(size_t)k * s] = tmp[k];
        }
        return;
    }
    /* h_j = (X_j - i X_{n-j}) e^{+i pi j/(2n)}, h_0 = X_0; t = n*IFFT(h) is real.
       Two lines share one transform: z = h_a + i h_b. */
    for (int j = 0; j < n; j++) {
        const double cr = d->q_re[j], ci = -d->q_im[j];   /* e^{+i pi j/(2n)} */
        double har, hai, hbr, hbi;
        if (j == 0) {
            har = a[0]; hai = 0.0;
            hbr = b ? b[0] : 0.0; hbi = 0.0;
        } else {
            const double xa = a[(size_t)j * s], ya = a[(size_t)(n - j) * s];
            har = xa * cr + ya * ci;  hai = xa * ci - ya * cr;
            if (b) {
                const double xb = b[(size_t)j * s], yb = b[(size_t)(n - j) * s];
                hbr = xb * cr + yb * ci;  hbi = xb * ci - yb * cr;
            } else { hbr = 0.0; hbi = 0.0; }
        }
        zr[j] = har - hbi;
        zi[j] = hai + hbr;
    }
    cfft(d, zr, zi, +1);
    for (int m = 0; m < n / 2; m++) {
        tmp[2 * m] = zr[m];
        tmp[2 * m + 1] = zr[n - 1 - m];
    }
    for (int k = 0; k < n; k++) a[(size_t)k * s] = tmp[k];
    if (b) {
        for (int m = 0; m < n / 2; m++) {
            tmp[2 * m] = zi[m];
            tmp[2 * m + 1] = zi[n - 1 - m];
        }
        for (int k = 0; k < n; k++) b[(size_t)k * s] = tmp[k];
    }
}

static void run_lines(const dct1d *d, fftw_r2r_kind kind, double *a, double *b, int s,
                      double *zr, double *zi, double *tmp) {
    if (kind == FFTW_REDFT10) dct2_pair(d, a, b, s, zr, zi, tmp);
    else dct3_pair(d, a, b, s, zr, zi, tmp);
}

fftw_plan fftw_plan_r2r_2d(int n0, int n1, double *in, double *out,
                           fftw_r2r_kind kind0, fftw_r2r_kind kind1, unsigned flags) {
    (void)in; (void)out; (void)flags;
    struct of2d_dct_plan *p = (struct of2d_dct_plan *)calloc(1, sizeof(*p));
    p->n0 = n0; p->n1 = n1; p->kind0 = kind0; p->kind1 = kind1;
    dct1d_init(&p->d0, n0);
    dct1d_init(&p->d1, n1);
    const int nmax = n0 > n1 ? n0 : n1;
    p->za_re = (double *)malloc(sizeof(double) * (size_t)nmax);
    p->za_im = (double *)malloc(sizeof(double) * (size_t)nmax);
    p->tmp   = (double *)malloc(sizeof(double) * (size_t)nmax);
    p->col_a = (double *)malloc(sizeof(double) * (size_t)n0);
    p->col_b = (double *)malloc(sizeof(double) * (size_t)n0);
    return p;
}

void fftw_execute_r2r(const fftw_plan p, double *in, double *out) {
    const int n0 = p->n0, n1 = p->n1;
    if (out != in) memcpy(out, in, sizeof(double) * (size_t)n0 * (size_t)n1);
    /* dimension 1 (contiguous) */
    for (int r = 0; r < n0; r += 2) {
        double *a = out + (size_t)r * n1;
        double *b = (r + 1 < n0) ? a + n1 : NULL;
        run_lines(&p->d1, p->kind1, a, b, 1, p->za_re, p->za_im, p->tmp);
    }
    /* dimension 0 (stride n1): gather two columns, transform, scatter */
    for (int c = 0; c < n1; c += 2) {
        const int two = (c + 1 < n1);
        for (int r = 0; r < n0; r++) {
            p->col_a[r] = out[(size_t)r * n1 + c];
            if (two) p->col_b[r] = out[(size_t)r * n1 + c + 1];
        }
        run_lines(&p->d0, p->kind0, p->col_a, two ? p->col_b : NULL, 1, p->za_re, p->za_im, p->tmp);
        for (int r = 0; r < n0; r++) {
            out[(size_t)r * n1 + c] = p->col_a[r];
            if (two) out[(size_t)r * n1 + c + 1] = p->col_b[r];
        }
    }
}

void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    dct1d_free(&p->d0);
    dct1d_free(&p->d1);
    free(p->za_re); free(p->za_im); free(p->tmp); free(p->col_a); free(p->col_b);
    free(p);
}

/* direct entry points for the tests (1-D, in place) */
void of2d_standin_dct1d(double *x, int n, int type /*2 or 3*/) {
    dct1d d;
    dct1d_init(&d, n);
    double *zr = (double *)malloc(sizeof(double) * (size_t)n);
    double *zi = (double *)malloc(sizeof(double) * (size_t)n);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    run_lines(&d, type == 2 ? FFTW_REDFT10 : FFTW_REDFT01, x, NULL, 1, zr, zi, tmp);
    free(zr); free(zi); free(tmp);
    dct1d_free(&d);
}
