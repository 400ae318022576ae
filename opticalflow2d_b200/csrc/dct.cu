// dct.cu -- shared-memory DCT kernels that replace fftw in the curvature solver
// (OpticalFlowCurvature.cpp:6-167; fftw_plan_r2r_2d REDFT10 / REDFT01, :52-55).
//
// Transform definitions (FFTW manual, unnormalised):
//   REDFT10 (DCT-II):  Y_k = 2 sum_j X_j cos(pi (j+1/2) k / n)
//   REDFT01 (DCT-III): Y_k = X_0 + 2 sum_{j>=1} X_j cos(pi j (k+1/2) / n)
//
// The x and y components of the motion field always travel together, so every line transform is
// ONE complex FFT of z = a + i b (a = x component, b = y component):
//   DCT-II  (Makhoul): v[m] = a[2m], v[n-1-m] = a[2m+1]; Z = FFT(v_a + i v_b);
//                      V_a = (Z_k + conj Z_{n-k})/2, V_b = (Z_k - conj Z_{n-k})/(2i); A_k = 2 Re(V_a e^{-i pi k/2n})
//   DCT-III          : h_j = (X_j - i X_{n-j}) e^{+i pi j/2n}, h_0 = X_0; t = n IFFT(h_a + i h_b);
//                      a[2m] = Re t[m], a[2m+1] = Re t[n-1-m] (b from Im)
// The forward FFT is an in-place radix-2 DIT fed in bit-reversed order by the (scattering) load; the
// inverse is an in-place DIF whose bit-reversed output is undone by the (gathering) store, so no
// separate permutation pass exists.
//
// One curvature iteration = three kernels (algorithmic traffic in DESIGN.md):
//   P1 rows    : rhs = u - tau f (f = L-SSD force, fused)  -> DCT-II along x  -> spectrum
//   P2 columns : DCT-II along y -> x 1/(1 + tau alpha lap^2) -> DCT-III along y (spectrum stays in smem)
//   P3 rows    : DCT-III along x -> u' = rhs / (4 N)
// Non-power-of-two lengths (fftw takes any n; the reference's demo pads to 278 = 2 x 139) use Bluestein's algorithm: the
// n-point DFT inside Makhoul's DCT is a chirp-modulated circular convolution of length M = 2^k >= 2n - 1, evaluated with the
// same radix-2 kernels (DIF forward -> pointwise product with the transformed chirp, stored bit-reversed -> DIT inverse), so
// a line costs two M-point FFTs: O(n log n) for every n.  Lines too long for shared memory (n > 4096) fall back to the direct
// O(n^2) sum.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <complex>
#include <vector>

#include "device_math.cuh"
#include "engine_internal.cuh"

namespace {

// engine hook: gating by the pair control block, ping-pong selection and the Logger epilogue (enabled = 0: plain step)
struct CurvHook {
    PairCtl *ctl;
    int *n_active;
    double *partials;
    size_t pstride;
    TraceDev tr;
    int enabled;
};

template <class S> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };
template <class S> using cplx_t = typename Cplx<S>::type;

template <class S> __device__ __forceinline__ cplx_t<S> cmul(cplx_t<S> a, cplx_t<S> b) {
    cplx_t<S> r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}

constexpr int FFT_THREADS = 256;

struct LineTables {
    const void *tw;    // e^{-2 pi i k / n}, k < n/2       (cplx<S>)
    const void *q;     // e^{-i pi k / (2n)}, k < n        (cplx<S>)
    const double *costab;  // cos(pi m / (2n)), m < 4n (direct path only)
    const void *tw16a;     // [4][n/16]:  e^{-2 pi i k 2^i / n}       (register path, dct_reg.cuh; 512 <= n <= 4096)
    const void *tw16b;     // [4][n/256]: e^{-2 pi i k 2^i / (n/16)}
    // Bluestein (non-power-of-two n): M = 2^log2M >= 2n - 1
    const void *twM;       // e^{-2 pi i k / M}, k < M/2
    const void *chirp;     // w_j = e^{-i pi j^2 / n}, j < n
    const void *bhat;      // FFT_M(b) / M in BIT-REVERSED order, b_j = conj(w_|j|) wrapped to length M
    int n, log2n, pow2;
    int M, log2M, bluestein;
};

// in-place radix-2 DIT, input in bit-reversed order, output natural. sign<0: forward.
template <class S>
__device__ void fft_dit(cplx_t<S> *x, int n, int log2n, const cplx_t<S> *__restrict__ tw, int sign) {
    for (int s = 0; s < log2n; s++) {
        const int half = 1 << s;
        for (int b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
            const int k = b & (half - 1);
            const int i0 = ((b >> s) << (s + 1)) + k;
            cplx_t<S> w = tw[k << (log2n - 1 - s)];
            if (sign > 0) w.y = -w.y;
            const cplx_t<S> a = x[i0], c = cmul<S>(w, x[i0 + half]);
            cplx_t<S> p, m;
            p.x = a.x + c.x; p.y = a.y + c.y; m.x = a.x - c.x; m.y = a.y - c.y;
            x[i0] = p; x[i0 + half] = m;
        }
        __syncthreads();
    }
}
// in-place radix-2 DIF, input natural, output in bit-reversed order.
template <class S>
__device__ void fft_dif(cplx_t<S> *x, int n, int log2n, const cplx_t<S> *__restrict__ tw, int sign) {
    for (int s = log2n - 1; s >= 0; s--) {
        const int half = 1 << s;
        for (int b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
            const int k = b & (half - 1);
            const int i0 = ((b >> s) << (s + 1)) + k;
            cplx_t<S> w = tw[k << (log2n - 1 - s)];
            if (sign > 0) w.y = -w.y;
            const cplx_t<S> a = x[i0], c = x[i0 + half];
            cplx_t<S> p, m;
            p.x = a.x + c.x; p.y = a.y + c.y; m.x = a.x - c.x; m.y = a.y - c.y;
            x[i0] = p; x[i0 + half] = cmul<S>(w, m);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int bitrev(int i, int log2n) { return (int)(__brev((unsigned)i) >> (32 - log2n)); }
// Makhoul reordering: natural position m -> FFT input position
__device__ __forceinline__ int makhoul_pos(int m, int n) { return (m & 1) ? n - 1 - (m >> 1) : (m >> 1); }

// dct_fast.cuh (included below): radix-8 / radix-4 FFTs on swizzled shared memory, M >= 64
__device__ __forceinline__ int swz(int i);
__device__ void fft_dit_fast(double2 *x, int n, int L, int nlines, const double2 *__restrict__ tw);
__device__ void fft_dif_fast(double2 *x, int n, int L, int nlines, const double2 *__restrict__ tw);

// n-point DFT by Bluestein's algorithm in the M-point buffer y.  Element j of the input sequence a_j = v_j w_j (j < n; the
// caller zero-fills the rest) goes to y[bs_in(j)]; after bluestein_core(), DFT_n(v)_k = w_k y[bs_out(k)], k < n.
//   M >= 64 (fast): forward DIT (bit-reversed in -> natural out), product with FFT_M(b) / M in natural order, inverse DIF
//                   (natural in -> bit-reversed out), all through the bank swizzle of dct_fast.cuh;
//   M <  64       : radix-2 DIF forward (natural -> bit-reversed), product with the bit-reversed table, DIT inverse.
__device__ __forceinline__ int bs_in(int j, const LineTables &T) { return T.log2M >= 6 ? swz(bitrev(j, T.log2M)) : j; }
__device__ __forceinline__ int bs_out(int k, const LineTables &T) { return T.log2M >= 6 ? swz(bitrev(k, T.log2M)) : k; }
template <class S>
__device__ void bluestein_core(cplx_t<S> *y, const LineTables &T) {
    const cplx_t<S> *bh = (const cplx_t<S> *)T.bhat;
    if constexpr (sizeof(S) == 8) {
        if (T.log2M >= 6) {
            fft_dit_fast(y, T.M, T.log2M, 1, (const double2 *)T.twM);
            for (int k = threadIdx.x; k < T.M; k += blockDim.x) { const int s = swz(k); y[s] = cmul<S>(y[s], bh[k]); }
            __syncthreads();
            fft_dif_fast(y, T.M, T.log2M, 1, (const double2 *)T.twM);
            return;
        }
    }
    fft_dif<S>(y, T.M, T.log2M, (const cplx_t<S> *)T.twM, -1);          // natural -> bit-reversed
    for (int k = threadIdx.x; k < T.M; k += blockDim.x) y[k] = cmul<S>(y[k], bh[k]);
    __syncthreads();
    fft_dit<S>(y, T.M, T.log2M, (const cplx_t<S> *)T.twM, +1);          // bit-reversed -> natural (1 / M is folded into bhat)
}

// smem slot where sample m of a line must be placed before dct2_line()
template <class S>
__device__ __forceinline__ int dct2_load_slot(int m, const LineTables &T) {
    return T.pow2 ? bitrev(makhoul_pos(m, T.n), T.log2n) : m;
}
// smem slot where output sample m of dct3_line() is found
template <class S>
__device__ __forceinline__ int dct3_store_slot(int m, const LineTables &T) {
    return T.pow2 ? bitrev(makhoul_pos(m, T.n), T.log2n) : m;
}

// DCT-II of the two real sequences packed in x (slots filled through dct2_load_slot); on return
// x[k] = (A_k, B_k) in natural order. `tmp` is only used by the direct path.
template <class S>
__device__ void dct2_line(cplx_t<S> *x, cplx_t<S> *tmp, const LineTables &T) {
    const int n = T.n;
    if (T.pow2) {
        fft_dit<S>(x, n, T.log2n, (const cplx_t<S> *)T.tw, -1);
        const cplx_t<S> *q = (const cplx_t<S> *)T.q;
        for (int k = threadIdx.x; k <= (n >> 1); k += blockDim.x) {
            const int nk = (n - k) & (n - 1);
            const cplx_t<S> zk = x[k], zn = x[nk];
            // V_a = (Z_k + conj Z_nk)/2, V_b = (Z_k - conj Z_nk)/(2i)
            {
                const S var = (S)0.5 * (zk.x + zn.x), vai = (S)0.5 * (zk.y - zn.y);
                const S vbr = (S)0.5 * (zk.y + zn.y), vbi = (S)-0.5 * (zk.x - zn.x);
                const cplx_t<S> w = q[k];
                cplx_t<S> o; o.x = (S)2 * (var * w.x - vai * w.y); o.y = (S)2 * (vbr * w.x - vbi * w.y);
                x[k] = o;
            }
            if (nk != k) {
                const S var = (S)0.5 * (zn.x + zk.x), vai = (S)0.5 * (zn.y - zk.y);
                const S vbr = (S)0.5 * (zn.y + zk.y), vbi = (S)-0.5 * (zn.x - zk.x);
                const cplx_t<S> w = q[nk];
                cplx_t<S> o; o.x = (S)2 * (var * w.x - vai * w.y); o.y = (S)2 * (vbr * w.x - vbi * w.y);
                x[nk] = o;
            }
        }
        __syncthreads();
    } else if (T.bluestein) {
        // x[m]: samples in natural order; tmp[M]: Makhoul-permuted, chirp-modulated, zero-padded
        const cplx_t<S> *w = (const cplx_t<S> *)T.chirp, *q = (const cplx_t<S> *)T.q;
        for (int j = n + threadIdx.x; j < T.M; j += blockDim.x) { cplx_t<S> z; z.x = (S)0; z.y = (S)0; tmp[bs_in(j, T)] = z; }
        for (int m = threadIdx.x; m < n; m += blockDim.x) { const int pos = makhoul_pos(m, n); tmp[bs_in(pos, T)] = cmul<S>(x[m], w[pos]); }
        __syncthreads();
        bluestein_core<S>(tmp, T);
        for (int k = threadIdx.x; k < n; k += blockDim.x) x[k] = cmul<S>(tmp[bs_out(k, T)], w[k]);   // Z_k = FFT(v_a + i v_b)_k
        __syncthreads();
        for (int k = threadIdx.x; k <= (n >> 1); k += blockDim.x) {
            const int nk = k == 0 ? 0 : n - k;
            const cplx_t<S> zk = x[k], zn = x[nk];
            cplx_t<S> ok, on;
            {
                const S var = (S)0.5 * (zk.x + zn.x), vai = (S)0.5 * (zk.y - zn.y);
                const S vbr = (S)0.5 * (zk.y + zn.y), vbi = (S)-0.5 * (zk.x - zn.x);
                const cplx_t<S> ww = q[k];
                ok.x = (S)2 * (var * ww.x - vai * ww.y); ok.y = (S)2 * (vbr * ww.x - vbi * ww.y);
            }
            {
                const S var = (S)0.5 * (zn.x + zk.x), vai = (S)0.5 * (zn.y - zk.y);
                const S vbr = (S)0.5 * (zn.y + zk.y), vbi = (S)-0.5 * (zn.x - zk.x);
                const cplx_t<S> ww = q[nk];
                on.x = (S)2 * (var * ww.x - vai * ww.y); on.y = (S)2 * (vbr * ww.x - vbi * ww.y);
            }
            x[k] = ok;
            if (nk != k) x[nk] = on;
        }
        __syncthreads();
    } else {
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            double sa = 0.0, sb = 0.0;
            for (int j = 0; j < n; j++) {
                const double c = T.costab[((long)(2 * j + 1) * k) % (4 * n)];
                sa += (double)x[j].x * c; sb += (double)x[j].y * c;
            }
            cplx_t<S> o; o.x = (S)(2.0 * sa); o.y = (S)(2.0 * sb);
            tmp[k] = o;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) x[k] = tmp[k];
        __syncthreads();
    }
}

// DCT-III of x[k] = (A_k, B_k) (natural order); outputs are read through dct3_store_slot().
template <class S>
__device__ void dct3_line(cplx_t<S> *x, cplx_t<S> *tmp, const LineTables &T) {
    const int n = T.n;
    if (T.pow2) {
        const cplx_t<S> *q = (const cplx_t<S> *)T.q;
        for (int j = threadIdx.x; j <= (n >> 1); j += blockDim.x) {
            if (j == 0) {
                cplx_t<S> o; o.x = x[0].x; o.y = x[0].y;   // h_0 = X_0 for both sequences: z_0 = A_0 + i B_0
                x[0] = o;
                continue;
            }
            const int nj = n - j;
            const cplx_t<S> Xj = x[j], Xn = x[nj];
            {   // slot j: e^{+i pi j/2n} = conj(q[j])
                const S cr = q[j].x, ci = -q[j].y;
                const S har = Xj.x * cr + Xn.x * ci, hai = Xj.x * ci - Xn.x * cr;
                const S hbr = Xj.y * cr + Xn.y * ci, hbi = Xj.y * ci - Xn.y * cr;
                cplx_t<S> o; o.x = har - hbi; o.y = hai + hbr;
                x[j] = o;
            }
            if (nj != j) {
                const S cr = q[nj].x, ci = -q[nj].y;
                const S har = Xn.x * cr + Xj.x * ci, hai = Xn.x * ci - Xj.x * cr;
                const S hbr = Xn.y * cr + Xj.y * ci, hbi = Xn.y * ci - Xj.y * cr;
                cplx_t<S> o; o.x = har - hbi; o.y = hai + hbr;
                x[nj] = o;
            }
        }
        __syncthreads();
        fft_dif<S>(x, n, T.log2n, (const cplx_t<S> *)T.tw, +1);
    } else if (T.bluestein) {
        const cplx_t<S> *w = (const cplx_t<S> *)T.chirp, *q = (const cplx_t<S> *)T.q;
        // h_j = (X_j - i X_{n-j}) e^{+i pi j / 2n} for both sequences, packed z_j = h_a + i h_b; the unnormalised inverse DFT
        // t = n IFFT(z) is conj(DFT(conj z)): tmp_j = conj(z_j) w_j
        for (int j = n + threadIdx.x; j < T.M; j += blockDim.x) { cplx_t<S> z; z.x = (S)0; z.y = (S)0; tmp[bs_in(j, T)] = z; }
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            cplx_t<S> z;
            if (j == 0) z = x[0];
            else {
                const cplx_t<S> Xj = x[j], Xn = x[n - j];
                const S cr = q[j].x, ci = -q[j].y;
                const S har = Xj.x * cr + Xn.x * ci, hai = Xj.x * ci - Xn.x * cr;
                const S hbr = Xj.y * cr + Xn.y * ci, hbi = Xj.y * ci - Xn.y * cr;
                z.x = har - hbi; z.y = hai + hbr;
            }
            z.y = -z.y;
            tmp[bs_in(j, T)] = cmul<S>(z, w[j]);
        }
        __syncthreads();
        bluestein_core<S>(tmp, T);
        // sample m of the output is t[makhoul_pos(m)] (a[2m'] = Re t[m'], a[2m'+1] = Re t[n-1-m'])
        for (int m = threadIdx.x; m < n; m += blockDim.x) {
            const int pos = makhoul_pos(m, n);
            cplx_t<S> t = cmul<S>(tmp[bs_out(pos, T)], w[pos]);
            t.y = -t.y;
            x[m] = t;
        }
        __syncthreads();
    } else {
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            double sa = 0.0, sb = 0.0;
            for (int j = 1; j < n; j++) {
                const double c = T.costab[((long)j * (2 * k + 1)) % (4 * n)];
                sa += (double)x[j].x * c; sb += (double)x[j].y * c;
            }
            cplx_t<S> o; o.x = (S)((double)x[0].x + 2.0 * sa); o.y = (S)((double)x[0].y + 2.0 * sb);
            tmp[k] = o;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) x[k] = tmp[k];
        __syncthreads();
    }
}

// ---- P1: rows.  One CTA per row j: rhs = u - tau f, DCT-II along x, spectrum row out -----------------
template <class R, class S>
__global__ void __launch_bounds__(FFT_THREADS) k_curv_rows_fwd(int nx, int ny, const vec2_t<R> *est0, const vec2_t<R> *est1, const vec2_t<R> *__restrict__ gradI,
                                                               const R *__restrict__ It, R tau, cplx_t<S> *__restrict__ spec, LineTables T, CurvHook H) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx_t<S> *x = reinterpret_cast<cplx_t<S> *>(smem_raw);
    cplx_t<S> *tmp = x + nx;
    const int j = blockIdx.x;
    const size_t pair_off = (size_t)blockIdx.y * nx * ny;
    const vec2_t<R> *__restrict__ u = est0;
    if (H.enabled) {
        const PairCtl *c = H.ctl + blockIdx.y;
        if (!__ldcg(&c->active)) return;
        u = __ldcg(&c->sel) ? est1 : est0;
    }
    u += pair_off; gradI += pair_off; It += pair_off; spec += pair_off;
    const size_t row = (size_t)j * nx;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const vec2_t<R> uu = u[row + i];
        const vec2_t<R> f = lssd_force<R>(gradI[row + i], It[row + i], uu);   // OpticalFlow.cpp:33
        cplx_t<S> v;                                                          // OpticalFlowCurvature.cpp:90-91
        v.x = (S)(uu.x - tau * f.x);
        v.y = (S)(uu.y - tau * f.y);
        x[dct2_load_slot<S>(i, T)] = v;
    }
    __syncthreads();
    dct2_line<S>(x, tmp, T);
    for (int p = threadIdx.x; p < nx; p += blockDim.x) spec[row + p] = x[p];
}

// ---- P2: columns.  One CTA per group of C adjacent columns p: DCT-II along y, eigenvalue, DCT-III ----
template <class S>
__global__ void __launch_bounds__(FFT_THREADS) k_curv_cols(int nx, int ny, int C, cplx_t<S> *__restrict__ spec, const double *__restrict__ cosx,
                                                           const double *__restrict__ cosy, double tau_alpha, LineTables T, CurvHook H) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (H.enabled && !__ldcg(&H.ctl[blockIdx.y].active)) return;
    spec += (size_t)blockIdx.y * nx * ny;
    cplx_t<S> *lines = reinterpret_cast<cplx_t<S> *>(smem_raw);   // [C][ny]
    cplx_t<S> *tmp = lines + (size_t)C * ny;                      // [ny] (direct path only)
    const int p0 = blockIdx.x * C;
    const int nc = min(C, nx - p0);
    for (int e = threadIdx.x; e < ny * nc; e += blockDim.x) {
        const int jj = e / nc, c = e % nc;
        lines[(size_t)c * ny + dct2_load_slot<S>(jj, T)] = spec[(size_t)jj * nx + p0 + c];
    }
    __syncthreads();
    for (int c = 0; c < nc; c++) {
        cplx_t<S> *x = lines + (size_t)c * ny;
        dct2_line<S>(x, tmp, T);
        const double cxp = cosx[p0 + c];
        for (int qq = threadIdx.x; qq < ny; qq += blockDim.x) {   // OpticalFlowCurvature.cpp:24, :135-136
            const double lap = -4 + cxp + cosy[qq];
            const double eig = 1.0f / (1.0f + tau_alpha * (lap * lap));
            cplx_t<S> v = x[qq];
            v.x = (S)((double)v.x * eig); v.y = (S)((double)v.y * eig);
            x[qq] = v;
        }
        __syncthreads();
        dct3_line<S>(x, tmp, T);
    }
    for (int e = threadIdx.x; e < ny * nc; e += blockDim.x) {
        const int jj = e / nc, c = e % nc;
        spec[(size_t)jj * nx + p0 + c] = lines[(size_t)c * ny + dct3_store_slot<S>(jj, T)];
    }
}

// ---- P3: rows.  DCT-III along x, u' = rhs / (4 N) (OpticalFlowCurvature.cpp:116-117) ----------------
template <class R, class S>
__global__ void __launch_bounds__(FFT_THREADS) k_curv_rows_inv(int nx, int ny, const cplx_t<S> *__restrict__ spec, vec2_t<R> *est0, vec2_t<R> *est1, R fourN,
                                                               LineTables T, CurvHook H) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx_t<S> *x = reinterpret_cast<cplx_t<S> *>(smem_raw);
    cplx_t<S> *tmp = x + nx;
    const int j = blockIdx.x, pair = blockIdx.y;
    const size_t pair_off = (size_t)pair * nx * ny;
    vec2_t<R> *__restrict__ unew = est1;
    const vec2_t<R> *__restrict__ uold = est0;
    PairCtl *c = nullptr;
    if (H.enabled) {
        c = H.ctl + pair;
        if (!__ldcg(&c->active)) return;
        if (__ldcg(&c->sel)) { unew = est0; uold = est1; }
    }
    unew += pair_off; uold += pair_off; spec += pair_off;
    const size_t row = (size_t)j * nx;
    for (int p = threadIdx.x; p < nx; p += blockDim.x) x[p] = spec[row + p];
    __syncthreads();
    dct3_line<S>(x, tmp, T);
    double sd = 0.0, sp = 0.0;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const cplx_t<S> v = x[dct3_store_slot<S>(i, T)];
        const vec2_t<R> o = mk2<R>((R)v.x / fourN, (R)v.y / fourN);
        unew[row + i] = o;
        if (H.enabled) {   // Logger.cpp:32-51: prev is the estimate this iteration started from
            const vec2_t<R> old = uold[row + i];
            sd += vec_norm_d<R>(mk2<R>(o.x - old.x, o.y - old.y));
            sp += vec_norm_d<R>(old);
        }
    }
    if (!H.enabled) return;
    block_sum2(sd, sp);
    const double vals[2] = {sd, sp};
    double *part = H.partials + (size_t)pair * H.pstride;
    if (publish_partials<2>(vals, part, &c->ticket[0], ny, j)) {
        double out[2];
        reduce_partials<2>(part, ny, out, 0u, 0u);
        if (threadIdx.x == 0) {
            c->sel ^= 1;
            finalize_logger<R>(c, H.tr, pair, out[0], out[1], (unsigned)(nx * ny), H.n_active);
        }
    }
}

// ---- stand-alone 2-D transform of a row-major n0 x n1 real array (one component; imaginary lane idle)
template <class S>
__global__ void __launch_bounds__(FFT_THREADS) k_dct_lines(int nlines, int n, size_t line_stride, size_t elem_stride, int kind, double *__restrict__ data,
                                                           LineTables T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx_t<S> *x = reinterpret_cast<cplx_t<S> *>(smem_raw);
    cplx_t<S> *tmp = x + n;
    double *line = data + (size_t)blockIdx.x * line_stride;
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        cplx_t<S> v; v.x = (S)line[(size_t)m * elem_stride]; v.y = (S)0;
        x[kind == 2 ? dct2_load_slot<S>(m, T) : m] = v;
    }
    __syncthreads();
    if (kind == 2) dct2_line<S>(x, tmp, T); else dct3_line<S>(x, tmp, T);
    for (int m = threadIdx.x; m < n; m += blockDim.x)
        line[(size_t)m * elem_stride] = (double)x[kind == 2 ? m : dct3_store_slot<S>(m, T)].x;
}

}  // namespace

#include "dct_fast.cuh"
#include "dct_reg.cuh"     // namespace rg: double precision (the reference's)
namespace {
// bank swizzle for 8-byte elements (float2): a shared-memory wavefront is 16 of them, so the low FOUR index bits are XORed with the
// folded higher bits (swz of dct_fast.cuh, made for 16-byte elements, spreads power-of-two strides over 8 slots only: 2-way conflicts)
__device__ __forceinline__ int swz8(int i) {
    const int x = i >> 4;
    return i ^ ((x ^ (x >> 4) ^ (x >> 8)) & 15);
}
}  // namespace
#define RG_NS rgf
#define RG_SWZ swz8
#define RG_C2 float2
#define RG_S float
#define RG_MK2 make_float2
#define RG_PACKED 1
#ifndef OF2D_RGF_MINB
#define OF2D_RGF_MINB 3
#endif
#define RG_MINB OF2D_RGF_MINB   // resident CTAs of 256 threads the register allocation aims at (16 float2 points per thread: 3 -> 85 registers)
#include "dct_reg.cuh"     // namespace rgf: single precision (relaxed engine, fp32 fields)

namespace {

// ---- host-side tables ---------------------------------------------------------------------------
const double kPi = 3.14159265358979323846264338327950288;

template <class S>
int build_tables(int n, LineTables *T, void **d_blob) {
    memset(T, 0, sizeof(*T));
    T->n = n;
    T->pow2 = n >= 2 && (n & (n - 1)) == 0;
    *d_blob = nullptr;
    if (T->pow2) {
        int l = 0;
        while ((1 << l) < n) l++;
        T->log2n = l;
        const bool reg = l >= 9 && l <= 12;
        const size_t ntw = n / 2, nq = n, s1 = reg ? n / 16 : 0, s2 = reg ? n / 256 : 0, tot = ntw + nq + 4 * s1 + 4 * s2;
        cplx_t<S> *h = (cplx_t<S> *)malloc(sizeof(cplx_t<S>) * tot);
        for (size_t k = 0; k < ntw; k++) { h[k].x = (S)cos(-2.0 * kPi * k / n); h[k].y = (S)sin(-2.0 * kPi * k / n); }
        for (size_t k = 0; k < nq; k++) { h[ntw + k].x = (S)cos(-kPi * k / (2.0 * n)); h[ntw + k].y = (S)sin(-kPi * k / (2.0 * n)); }
        for (int i = 0; i < 4; i++) {
            for (size_t k = 0; k < s1; k++) { const double a = -2.0 * kPi * (double)(k << i) / n; h[ntw + nq + i * s1 + k].x = (S)cos(a); h[ntw + nq + i * s1 + k].y = (S)sin(a); }
            for (size_t k = 0; k < s2; k++) { const double a = -2.0 * kPi * (double)(k << i) / (n / 16); h[ntw + nq + 4 * s1 + i * s2 + k].x = (S)cos(a); h[ntw + nq + 4 * s1 + i * s2 + k].y = (S)sin(a); }
        }
        cudaError_t e = cudaMalloc(d_blob, sizeof(cplx_t<S>) * tot);
        if (e == cudaSuccess) e = cudaMemcpy(*d_blob, h, sizeof(cplx_t<S>) * tot, cudaMemcpyHostToDevice);
        free(h);
        if (e != cudaSuccess) { of2d_set_error("dct tables: %s", cudaGetErrorString(e)); return OF2D_ERR_CUDA; }
        T->tw = *d_blob;
        T->q = (const cplx_t<S> *)*d_blob + ntw;
        if (reg) { T->tw16a = (const cplx_t<S> *)*d_blob + ntw + nq; T->tw16b = (const cplx_t<S> *)*d_blob + ntw + nq + 4 * s1; }
    } else if (n >= 2 && n <= 4096) {
        // Bluestein tables: [twM (M/2)] [q (n)] [chirp (n)] [bhat (M, bit-reversed, / M)]
        int l = 0;
        while ((1 << l) < 2 * n - 1) l++;
        const int M = 1 << l;
        T->M = M; T->log2M = l; T->bluestein = 1;
        const size_t tot = (size_t)M / 2 + 2 * (size_t)n + (size_t)M;
        cplx_t<S> *h = (cplx_t<S> *)malloc(sizeof(cplx_t<S>) * tot);
        cplx_t<S> *twM = h, *q = h + M / 2, *chirp = q + n, *bhat = chirp + n;
        for (int k = 0; k < M / 2; k++) { twM[k].x = (S)cos(-2.0 * kPi * k / M); twM[k].y = (S)sin(-2.0 * kPi * k / M); }
        for (int k = 0; k < n; k++) { q[k].x = (S)cos(-kPi * k / (2.0 * n)); q[k].y = (S)sin(-kPi * k / (2.0 * n)); }
        std::vector<std::complex<double>> b((size_t)M, std::complex<double>(0.0, 0.0));
        for (int j = 0; j < n; j++) {
            const long r = ((long)j * j) % (2L * n);                 // j^2 mod 2n keeps the angle exact
            const double a = kPi * (double)r / n;
            chirp[j].x = (S)cos(a); chirp[j].y = (S)-sin(a);         // w_j = e^{-i pi j^2 / n}
            const std::complex<double> bj(cos(a), sin(a));
            b[(size_t)j] = bj;
            if (j) b[(size_t)(M - j)] = bj;
        }
        // in-place radix-2 DIF on the host: natural in, bit-reversed out -- the order the device multiplies in
        for (int s = l - 1; s >= 0; s--) {
            const int half = 1 << s;
            for (int blk = 0; blk < M; blk += 2 * half)
                for (int k = 0; k < half; k++) {
                    const double a = -2.0 * kPi * (double)((long)k << (l - 1 - s)) / M;
                    const std::complex<double> wv(cos(a), sin(a)), p = b[(size_t)(blk + k)], m2 = b[(size_t)(blk + k + half)];
                    b[(size_t)(blk + k)] = p + m2;
                    b[(size_t)(blk + k + half)] = (p - m2) * wv;
                }
        }
        // b[] now holds FFT_M(b) in bit-reversed order: kept like that for the radix-2 device path (M < 64), natural order for the fast one
        for (int k = 0; k < M; k++) {
            unsigned r = 0;
            for (int t = 0; t < l; t++) r |= ((unsigned)k >> t & 1u) << (l - 1 - t);
            const std::complex<double> v = l >= 6 ? b[(size_t)r] : b[(size_t)k];
            bhat[k].x = (S)(v.real() / M); bhat[k].y = (S)(v.imag() / M);
        }
        cudaError_t e = cudaMalloc(d_blob, sizeof(cplx_t<S>) * tot);
        if (e == cudaSuccess) e = cudaMemcpy(*d_blob, h, sizeof(cplx_t<S>) * tot, cudaMemcpyHostToDevice);
        free(h);
        if (e != cudaSuccess) { of2d_set_error("dct tables: %s", cudaGetErrorString(e)); return OF2D_ERR_CUDA; }
        T->twM = *d_blob;
        T->q = (const cplx_t<S> *)*d_blob + M / 2;
        T->chirp = (const cplx_t<S> *)*d_blob + M / 2 + n;
        T->bhat = (const cplx_t<S> *)*d_blob + M / 2 + 2 * n;
    } else {
        double *h = (double *)malloc(sizeof(double) * 4 * (size_t)n);
        for (int m = 0; m < 4 * n; m++) h[m] = cos(kPi * m / (2.0 * n));
        cudaError_t e = cudaMalloc(d_blob, sizeof(double) * 4 * (size_t)n);
        if (e == cudaSuccess) e = cudaMemcpy(*d_blob, h, sizeof(double) * 4 * (size_t)n, cudaMemcpyHostToDevice);
        free(h);
        if (e != cudaSuccess) { of2d_set_error("dct tables: %s", cudaGetErrorString(e)); return OF2D_ERR_CUDA; }
        T->costab = (const double *)*d_blob;
    }
    return OF2D_SUCCESS;
}

constexpr size_t kMaxSmem = 220 * 1024;   // dynamic budget: leaves room for the static reduction scratch

}  // namespace

struct of2d_curvature_plan {
    of2d_ctx *ctx;
    int nx, ny, real_is_double, spec_is_double;
    double tau, alpha, tau_alpha;
    LineTables Tx, Ty;
    void *blob_x, *blob_y;
    LineTables Txf, Tyf;        // single-precision tables of the register path (relaxed engine, fp32 fields)
    void *blob_xf, *blob_yf;
    int spec_f32;
    double *d_cosx, *d_cosy;
    float *d_cosxf, *d_cosyf;   // the same rounded to single precision (column kernel of the relaxed register path)
    void *d_spec;    // spectrum between the row and the column pass (transposed on the fast paths)
    void *d_spec2;   // register path: column pass output in the natural layout
    int cols_per_cta;
    size_t smem_rows, smem_cols;
    int batch;
};

namespace {

// dct_reg.cuh: register-blocked radix-16 path (both line lengths in 512 .. 4096).  A = rg::Api (double transform) or rgf::Api
// (float transform, tables P->Txf / P->Tyf: relaxed engine on fp32 fields).
// fuse_next: the inverse row pass also runs the forward row pass of the next iteration (k_rg_rows_inv<.., FUSE>)
template <class A> const LineTables &tables_x(const of2d_curvature_plan *P) { return sizeof(typename A::C2) == 8 ? P->Txf : P->Tx; }
template <class A> const LineTables &tables_y(const of2d_curvature_plan *P) { return sizeof(typename A::C2) == 8 ? P->Tyf : P->Ty; }
template <class A, class R, int LX>
int launch_reg_rows(of2d_curvature_plan *P, bool fwd, const R *u, R *unew, const R *gradI, const R *It, const CurvHook &H, int batch, bool fuse_next = false) {
    using C2 = typename A::C2;
    of2d_ctx *ctx = P->ctx;
    constexpr int LPC = 2, NT = LPC * A::template G<LX>::TPL;
    const size_t smem = sizeof(C2) * (size_t)P->nx * LPC;
    const LineTables &TX = tables_x<A>(P);
    const typename A::Tw T{(const C2 *)TX.tw16a, (const C2 *)TX.tw16b};
    const R fourN = (R)4.0f * (R)(unsigned)(P->nx * P->ny);
    if (fwd) {
        constexpr auto kern = A::template rows_fwd<R, LX, LPC>();
        { int st = of2d_ensure_dynamic_smem((const void *)kern, smem); if (st) return st; }
        ProfScope _ps(ctx, "curv_rows_fwd");
        pdl_launch(kern, dim3(P->ny / LPC, batch), NT, smem, ctx->stream, P->ny, (const vec2_t<R> *)u, (const vec2_t<R> *)unew, (const vec2_t<R> *)gradI, It,
                   (R)P->tau, (C2 *)P->d_spec, (const C2 *)TX.q, T, H);
    } else if (fuse_next) {
        constexpr auto kern = A::template rows_inv<R, LX, LPC, true>();
        { int st = of2d_ensure_dynamic_smem((const void *)kern, smem); if (st) return st; }
        ProfScope _ps(ctx, "curv_rows_inv_fwd");
        pdl_launch(kern, dim3(P->ny / LPC, batch), NT, smem, ctx->stream, P->ny, (const C2 *)P->d_spec2, (vec2_t<R> *)u, (vec2_t<R> *)unew, fourN,
                   (const C2 *)TX.q, T, H, (const vec2_t<R> *)gradI, It, (R)P->tau, (C2 *)P->d_spec);
    } else {
        constexpr auto kern = A::template rows_inv<R, LX, LPC, false>();
        { int st = of2d_ensure_dynamic_smem((const void *)kern, smem); if (st) return st; }
        ProfScope _ps(ctx, "curv_rows_inv");
        pdl_launch(kern, dim3(P->ny / LPC, batch), NT, smem, ctx->stream, P->ny, (const C2 *)P->d_spec2, (vec2_t<R> *)u, (vec2_t<R> *)unew, fourN,
                   (const C2 *)TX.q, T, H, (const vec2_t<R> *)nullptr, (const R *)nullptr, (R)0, (C2 *)nullptr);
    }
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}
template <class A, int LY>
int launch_reg_cols(of2d_curvature_plan *P, const CurvHook &H, int batch) {
    using C2 = typename A::C2;
    of2d_ctx *ctx = P->ctx;
    constexpr int NT = 2 * A::template G<LY>::TPL;
    const size_t smem = sizeof(C2) * (size_t)P->ny * 2;
    const LineTables &TY = tables_y<A>(P);
    const typename A::Tw T{(const C2 *)TY.tw16a, (const C2 *)TY.tw16b};
    constexpr auto kern = A::template cols<LY>();
    { int st = of2d_ensure_dynamic_smem((const void *)kern, smem); if (st) return st; }
    ProfScope _ps(ctx, "curv_cols");
    pdl_launch(kern, dim3(P->nx / 2, batch), NT, smem, ctx->stream, P->nx, (const C2 *)P->d_spec, (C2 *)P->d_spec2,
               (const typename A::S *)(sizeof(typename A::S) == 4 ? (const void *)P->d_cosxf : (const void *)P->d_cosx),
               (const typename A::S *)(sizeof(typename A::S) == 4 ? (const void *)P->d_cosyf : (const void *)P->d_cosy), P->tau_alpha, (const C2 *)TY.q, T, H);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}
// flags (engine loop only): OF2D_CURV_SKIP_FWD = the forward row pass of this iteration was already run by the previous
// iteration's fused inverse pass; OF2D_CURV_FUSE_NEXT = this iteration's inverse pass runs the next iteration's forward pass
template <class A, class R>
int curvature_step_reg_a(of2d_curvature_plan *P, const R *u, R *unew, const R *gradI, const R *It, const CurvHook &H, int flags) {
    const int batch = H.enabled ? P->batch : 1;
    int st = OF2D_SUCCESS;
    const bool fuse_next = H.enabled && (flags & OF2D_CURV_FUSE_NEXT);
    for (int phase = (H.enabled && (flags & OF2D_CURV_SKIP_FWD)) ? 1 : 0; phase < 3 && st == OF2D_SUCCESS; phase++) {
        if (phase == 1) {
            switch (P->Ty.log2n) {
                case 9: st = launch_reg_cols<A, 9>(P, H, batch); break;
                case 10: st = launch_reg_cols<A, 10>(P, H, batch); break;
                case 11: st = launch_reg_cols<A, 11>(P, H, batch); break;
                default: st = launch_reg_cols<A, 12>(P, H, batch); break;
            }
        } else {
            switch (P->Tx.log2n) {
                case 9: st = launch_reg_rows<A, R, 9>(P, phase == 0, u, unew, gradI, It, H, batch, phase == 2 && fuse_next); break;
                case 10: st = launch_reg_rows<A, R, 10>(P, phase == 0, u, unew, gradI, It, H, batch, phase == 2 && fuse_next); break;
                case 11: st = launch_reg_rows<A, R, 11>(P, phase == 0, u, unew, gradI, It, H, batch, phase == 2 && fuse_next); break;
                default: st = launch_reg_rows<A, R, 12>(P, phase == 0, u, unew, gradI, It, H, batch, phase == 2 && fuse_next); break;
            }
        }
    }
    return st;
}
template <class R>
int curvature_step_reg(of2d_curvature_plan *P, const R *u, R *unew, const R *gradI, const R *It, const CurvHook &H, int flags) {
    if constexpr (sizeof(R) == 4) {
        if (P->spec_f32) return curvature_step_reg_a<rgf::Api, R>(P, u, unew, gradI, It, H, flags);
    }
    return curvature_step_reg_a<rg::Api, R>(P, u, unew, gradI, It, H, flags);
}

template <class R, class S>
int curvature_step_impl(of2d_curvature_plan *P, const R *u, R *unew, const R *gradI, const R *It, const CurvHook &H, int flags = 0) {
    of2d_ctx *ctx = P->ctx;
    const int batch = H.enabled ? P->batch : 1;
    const int nx = P->nx, ny = P->ny;
    cplx_t<S> *spec = (cplx_t<S> *)P->d_spec;
    const R fourN = (R)4.0f * (R)(unsigned)(nx * ny);
    if (P->Tx.tw16a && P->Ty.tw16a && !getenv("OF2D_NO_REG_FFT")) return curvature_step_reg<R>(P, u, unew, gradI, It, H, flags);
    constexpr int LPC = 2;
    const bool fast = P->Tx.pow2 && P->Ty.pow2 && P->Tx.log2n >= 6 && P->Ty.log2n >= 6 && sizeof(double2) * (size_t)nx * LPC <= kMaxSmem &&
                      sizeof(double2) * (size_t)ny <= kMaxSmem;
    if (fast) {   // dct_fast.cuh: transposed spectrum, radix-8/4 FFT
        const size_t smem_r = sizeof(double2) * (size_t)nx * LPC, smem_c = sizeof(double2) * (size_t)ny;
        { int st = of2d_ensure_dynamic_smem((const void *)k_cf_rows_fwd<R, LPC>, smem_r); if (st) return st; }
        { int st = of2d_ensure_dynamic_smem((const void *)k_cf_rows_inv<R, LPC>, smem_r); if (st) return st; }
        { int st = of2d_ensure_dynamic_smem((const void *)k_cf_cols, smem_c); if (st) return st; }
        double2 *specT = (double2 *)P->d_spec;
        { ProfScope _ps(ctx, "curv_rows_fwd");
        k_cf_rows_fwd<R, LPC><<<dim3(ny / LPC, batch), FFT_THREADS, smem_r, ctx->stream>>>(nx, ny, (const vec2_t<R> *)u, (const vec2_t<R> *)unew, (const vec2_t<R> *)gradI, It,
                                                                                          (R)P->tau, specT, P->Tx, H); }
        OF2D_LAUNCH_CHECK(ctx);
        { ProfScope _ps(ctx, "curv_cols");
        static int cols_threads = 0;
        if (!cols_threads) { const char *e = getenv("OF2D_FFT_COLS_THREADS"); cols_threads = e && atoi(e) >= 32 ? atoi(e) : FFT_THREADS; }
        k_cf_cols<<<dim3(nx, batch), cols_threads, smem_c, ctx->stream>>>(nx, ny, specT, P->d_cosx, P->d_cosy, P->tau_alpha, P->Ty, H); }
        OF2D_LAUNCH_CHECK(ctx);
        { ProfScope _ps(ctx, "curv_rows_inv");
        k_cf_rows_inv<R, LPC><<<dim3(ny / LPC, batch), FFT_THREADS, smem_r, ctx->stream>>>(nx, ny, specT, (vec2_t<R> *)u, (vec2_t<R> *)unew, fourN, P->Tx, H); }
        OF2D_LAUNCH_CHECK(ctx);
        return OF2D_SUCCESS;
    }
    { int st = of2d_ensure_dynamic_smem((const void *)k_curv_rows_fwd<R, S>, P->smem_rows); if (st) return st; }
    { int st = of2d_ensure_dynamic_smem((const void *)k_curv_rows_inv<R, S>, P->smem_rows); if (st) return st; }
    { int st = of2d_ensure_dynamic_smem((const void *)k_curv_cols<S>, P->smem_cols); if (st) return st; }
    k_curv_rows_fwd<R, S><<<dim3(ny, batch), FFT_THREADS, P->smem_rows, ctx->stream>>>(nx, ny, (const vec2_t<R> *)u, (const vec2_t<R> *)unew, (const vec2_t<R> *)gradI, It,
                                                                                      (R)P->tau, spec, P->Tx, H);
    OF2D_LAUNCH_CHECK(ctx);
    k_curv_cols<S><<<dim3(ceil_div(nx, P->cols_per_cta), batch), FFT_THREADS, P->smem_cols, ctx->stream>>>(nx, ny, P->cols_per_cta, spec, P->d_cosx, P->d_cosy,
                                                                                                       P->tau_alpha, P->Ty, H);
    OF2D_LAUNCH_CHECK(ctx);
    k_curv_rows_inv<R, S><<<dim3(ny, batch), FFT_THREADS, P->smem_rows, ctx->stream>>>(nx, ny, spec, (vec2_t<R> *)u, (vec2_t<R> *)unew, fourN, P->Tx, H);
    OF2D_LAUNCH_CHECK(ctx);
    return OF2D_SUCCESS;
}

}  // namespace

extern "C" {

int of2d_curvature_plan_create(of2d_ctx *ctx, int nx, int ny, double alpha, double tau, int real_is_double, of2d_curvature_plan **out) {
    *out = nullptr;
    OF2D_REQUIRE(nx > 0 && ny > 0, "bad dimensions");
    of2d_curvature_plan *P = new of2d_curvature_plan();
    memset(P, 0, sizeof(*P));
    P->ctx = ctx; P->nx = nx; P->ny = ny; P->real_is_double = real_is_double; P->spec_is_double = 1;
    P->alpha = alpha; P->tau = tau; P->batch = 1;
    // the reference multiplies tau*alpha in `float` (fp32 build) before promoting (OpticalFlowCurvature.cpp:24)
    P->tau_alpha = real_is_double ? tau * alpha : (double)((float)tau * (float)alpha);
    int st = build_tables<double>(nx, &P->Tx, &P->blob_x);
    if (st == OF2D_SUCCESS) st = build_tables<double>(ny, &P->Ty, &P->blob_y);
    if (st != OF2D_SUCCESS) { of2d_curvature_plan_destroy(P); return st; }
    // 2 cos(p PI / n) with the reference's truncated PI (OpticalFlowCurvature.cpp:4), evaluated by the host libm
    const double REF_PI = 3.14159265;
    double *hx = (double *)malloc(sizeof(double) * nx), *hy = (double *)malloc(sizeof(double) * ny);
    for (int p = 0; p < nx; p++) hx[p] = 2 * cos((unsigned)p * REF_PI / (unsigned)nx);
    for (int q = 0; q < ny; q++) hy[q] = 2 * cos((unsigned)q * REF_PI / (unsigned)ny);
    cudaError_t e = cudaMalloc(&P->d_cosx, sizeof(double) * nx);
    if (e == cudaSuccess) e = cudaMalloc(&P->d_cosy, sizeof(double) * ny);
    if (e == cudaSuccess) e = cudaMemcpy(P->d_cosx, hx, sizeof(double) * nx, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(P->d_cosy, hy, sizeof(double) * ny, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&P->d_spec, sizeof(double2) * (size_t)nx * ny);
    if (e == cudaSuccess && P->Tx.tw16a && P->Ty.tw16a) e = cudaMalloc(&P->d_spec2, sizeof(double2) * (size_t)nx * ny);
    free(hx); free(hy);
    if (e != cudaSuccess) {
        of2d_set_error("curvature plan: %s", cudaGetErrorString(e));
        of2d_curvature_plan_destroy(P);
        return OF2D_ERR_CUDA;
    }
    const size_t cs = sizeof(double2);
    // scratch next to the line: none (power of two), the M-point convolution buffer (Bluestein), a second line (direct sum)
    auto scratch = [&](const LineTables &T) -> size_t { return T.pow2 ? 0 : T.bluestein ? cs * (size_t)T.M : cs * (size_t)T.n; };
    P->smem_rows = cs * (size_t)nx + scratch(P->Tx);
    const size_t per_col = cs * (size_t)ny, extra = scratch(P->Ty);
    int C = 4;
    while (C > 1 && per_col * C + extra > kMaxSmem) C >>= 1;
    P->cols_per_cta = C;
    P->smem_cols = per_col * C + extra;
    if (P->smem_rows > kMaxSmem || P->smem_cols > kMaxSmem) {
        of2d_set_error("curvature plan: a %d x %d line does not fit in shared memory", nx, ny);
        of2d_curvature_plan_destroy(P);
        return OF2D_ERR_UNSUPPORTED;
    }
    *out = P;
    return OF2D_SUCCESS;
}

void of2d_curvature_plan_destroy(of2d_curvature_plan *P) {
    if (!P) return;
    cudaStreamSynchronize(P->ctx->stream);
    cudaFree(P->blob_x); cudaFree(P->blob_y); cudaFree(P->blob_xf); cudaFree(P->blob_yf); cudaFree(P->d_cosx); cudaFree(P->d_cosy); cudaFree(P->d_cosxf); cudaFree(P->d_cosyf); cudaFree(P->d_spec); cudaFree(P->d_spec2);
    delete P;
}

int of2d_curvature_step_f32(of2d_curvature_plan *P, const float *u, float *unew, const float *g, const float *It) {
    OF2D_REQUIRE(!P->real_is_double, "plan was created for double fields");
    OF2D_REQUIRE(u != unew, "curvature step is out of place");
    CurvHook H; memset(&H, 0, sizeof(H));
    return curvature_step_impl<float, double>(P, u, unew, g, It, H);
}
int of2d_curvature_step_f64(of2d_curvature_plan *P, const double *u, double *unew, const double *g, const double *It) {
    OF2D_REQUIRE(P->real_is_double, "plan was created for float fields");
    OF2D_REQUIRE(u != unew, "curvature step is out of place");
    CurvHook H; memset(&H, 0, sizeof(H));
    return curvature_step_impl<double, double>(P, u, unew, g, It, H);
}

}  // extern "C"

// engine entry points (engine_internal.cuh)
// relaxed engine, fp32 fields, register path: the DCTs and the spectrum between the passes in single precision
int of2d_curvature_plan_set_relaxed(of2d_curvature_plan *P, int on) {
    if (!on || P->real_is_double || !(P->Tx.tw16a && P->Ty.tw16a)) { P->spec_f32 = 0; return OF2D_SUCCESS; }
    { const char *e = getenv("OF2D_CURV_F32"); if (e && atoi(e) == 0) { P->spec_f32 = 0; return OF2D_SUCCESS; } }
    if (!P->blob_xf) {
        int st = build_tables<float>(P->nx, &P->Txf, &P->blob_xf);
        if (st == OF2D_SUCCESS) st = build_tables<float>(P->ny, &P->Tyf, &P->blob_yf);
        if (st != OF2D_SUCCESS) return st;
    }
    if (!P->d_cosxf) {   // 2 cos(p PI / n) as the column kernel used to round it on the fly: (float) of the double table's entries
        const double REF_PI = 3.14159265;
        const int nx = P->nx, ny = P->ny;
        float *h = (float *)malloc(sizeof(float) * (size_t)(nx + ny));
        for (int p = 0; p < nx; p++) h[p] = (float)(2 * cos((unsigned)p * REF_PI / (unsigned)nx));
        for (int q = 0; q < ny; q++) h[nx + q] = (float)(2 * cos((unsigned)q * REF_PI / (unsigned)ny));
        cudaError_t e = cudaMalloc(&P->d_cosxf, sizeof(float) * nx);
        if (e == cudaSuccess) e = cudaMalloc(&P->d_cosyf, sizeof(float) * ny);
        if (e == cudaSuccess) e = cudaMemcpy(P->d_cosxf, h, sizeof(float) * nx, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(P->d_cosyf, h + nx, sizeof(float) * ny, cudaMemcpyHostToDevice);
        free(h);
        if (e != cudaSuccess) { of2d_set_error("curvature plan: %s", cudaGetErrorString(e)); return OF2D_ERR_CUDA; }
    }
    P->spec_f32 = 1;
    return OF2D_SUCCESS;
}
int of2d_curvature_plan_set_batch(of2d_curvature_plan *P, int batch) {
    if (batch == P->batch) return OF2D_SUCCESS;
    OF2D_CUDA_TRY(cudaStreamSynchronize(P->ctx->stream));
    OF2D_CUDA_TRY(cudaFree(P->d_spec));
    P->d_spec = nullptr;
    OF2D_CUDA_TRY(cudaMalloc(&P->d_spec, sizeof(double2) * (size_t)P->nx * P->ny * batch));
    if (P->d_spec2) {
        OF2D_CUDA_TRY(cudaFree(P->d_spec2));
        P->d_spec2 = nullptr;
        OF2D_CUDA_TRY(cudaMalloc(&P->d_spec2, sizeof(double2) * (size_t)P->nx * P->ny * batch));
    }
    P->batch = batch;
    return OF2D_SUCCESS;
}

int of2d_curvature_engine_step(of2d_curvature_plan *P, PairCtl *ctl, int *n_active, double *partials, size_t pstride, TraceDev tr, void *est0, void *est1,
                               const void *gradI, const void *It, int flags) {
    CurvHook H;
    H.ctl = ctl; H.n_active = n_active; H.partials = partials; H.pstride = pstride; H.tr = tr; H.enabled = 1;
    if (P->real_is_double) return curvature_step_impl<double, double>(P, (const double *)est0, (double *)est1, (const double *)gradI, (const double *)It, H, flags);
    return curvature_step_impl<float, double>(P, (const float *)est0, (float *)est1, (const float *)gradI, (const float *)It, H, flags);
}
int of2d_curvature_plan_fuses_rows(const of2d_curvature_plan *P) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("OF2D_CURV_FUSE"); on = e ? atoi(e) : 1; }
    // fp32 fields by default: in fp64 the fused kernel holds the new estimate as 16 more double2 registers and spills at
    // 2048 / 4096 (OF2D_CURV_FUSE=2 fuses there too)
    return on && (on >= 2 || !P->real_is_double) && P->Tx.tw16a && P->Ty.tw16a && !getenv("OF2D_NO_REG_FFT");
}

extern "C" {

int of2d_dct2d_f64(of2d_ctx *ctx, int n0, int n1, int kind, double *d) {
    OF2D_REQUIRE(n0 > 0 && n1 > 0 && (kind == 2 || kind == 3), "bad arguments");
    LineTables T0, T1;
    void *b0 = nullptr, *b1 = nullptr;
    int st = build_tables<double>(n0, &T0, &b0);
    if (st == OF2D_SUCCESS) st = build_tables<double>(n1, &T1, &b1);
    if (st == OF2D_SUCCESS) {
        auto need = [&](const LineTables &T) -> size_t { return sizeof(double2) * ((size_t)T.n + (T.pow2 ? 0 : T.bluestein ? (size_t)T.M : (size_t)T.n)); };
        const size_t s1 = need(T1), s0 = need(T0);
        if (s0 > kMaxSmem || s1 > kMaxSmem) { of2d_set_error("of2d_dct2d_f64: line too long for shared memory"); st = OF2D_ERR_UNSUPPORTED; }
        else {
            of2d_ensure_dynamic_smem((const void *)k_dct_lines<double>, s1 > s0 ? s1 : s0);
            k_dct_lines<double><<<n0, FFT_THREADS, s1, ctx->stream>>>(n0, n1, (size_t)n1, 1, kind, d, T1);   // dimension 1 (contiguous)
            ctx->launches++;
            k_dct_lines<double><<<n1, FFT_THREADS, s0, ctx->stream>>>(n1, n0, 1, (size_t)n1, kind, d, T0);   // dimension 0 (stride n1)
            ctx->launches++;
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { of2d_set_error("of2d_dct2d_f64: %s", cudaGetErrorString(e)); st = OF2D_ERR_CUDA; }
        }
    }
    cudaFree(b0); cudaFree(b1);
    return st;
}

}  // extern "C"
