// dct_fast.cuh -- power-of-two fast path of the curvature solver's DCTs (included by dct.cu).
//
// Same mathematics as the generic kernels in dct.cu (Makhoul's N-point DCT through one N-point complex FFT
// that carries the x and y components of the motion as real and imaginary part), restructured for B200:
//   * the FFT runs radix-8 on aligned blocks in registers (stages 0-2) and fused radix-4 passes afterwards:
//     5 shared-memory passes and block barriers for N = 2048 instead of 11;
//   * shared-memory indices go through an XOR swizzle of the low three bits with the folded high bits, so the
//     8 lanes of a 128-bit access phase hit 8 distinct bank groups for every power-of-two stride, for the
//     bit-reversal scatter and for the contiguous blocks alike;
//   * the spectrum between the row and the column pass is stored TRANSPOSED (spec_T[p][j]); the row kernels
//     handle LPC lines per CTA so that every store / load is a full 32-byte sector, and the column kernel
//     streams one contiguous line per CTA (7 CTAs per SM instead of 1);
//   * DCT-II post-twiddle, eigenvalue multiply and DCT-III pre-twiddle of the column pass are one sweep over
//     the (k, n-k) pairs.
#pragma once

namespace {

__device__ __forceinline__ int swz(int i) {
    const int x = i >> 3;
    return i ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7);
}

template <int SIGN>
__device__ __forceinline__ double2 ld_tw(const double2 *__restrict__ tw, int idx) {
    double2 w = tw[idx];
    if (SIGN > 0) w.y = -w.y;
    return w;
}
__device__ __forceinline__ double2 cmuld(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 caddd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csubd(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// forward FFT, decimation in time: input in bit-reversed order, output natural; `nlines` lines of n = 2^L (L >= 6)
__device__ void fft_dit_fast(double2 *x, int n, int L, int nlines, const double2 *__restrict__ tw) {
    const int tid = threadIdx.x, nt = blockDim.x;
    {   // stages 0..2 on aligned blocks of 8, in registers
        const int bpl = n >> 3, nblk = nlines * bpl;
        for (int b = tid; b < nblk; b += nt) {
            double2 *base = x + (b / bpl) * n;
            const int i0 = (b % bpl) << 3;
            double2 a[8];
#pragma unroll
            for (int r = 0; r < 8; r++) a[r] = base[swz(i0 + r)];
#pragma unroll
            for (int r = 0; r < 8; r += 2) { const double2 p = caddd(a[r], a[r + 1]), m = csubd(a[r], a[r + 1]); a[r] = p; a[r + 1] = m; }
            {
                const double2 w1 = ld_tw<-1>(tw, 1 << (L - 2));
#pragma unroll
                for (int r = 0; r < 8; r += 4) {
                    const double2 c0 = a[r + 2], c1 = cmuld(w1, a[r + 3]);
                    const double2 p0 = caddd(a[r], c0), m0 = csubd(a[r], c0), p1 = caddd(a[r + 1], c1), m1 = csubd(a[r + 1], c1);
                    a[r] = p0; a[r + 2] = m0; a[r + 1] = p1; a[r + 3] = m1;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const double2 c = r == 0 ? a[4] : cmuld(ld_tw<-1>(tw, r << (L - 3)), a[r + 4]);
                const double2 p = caddd(a[r], c), m = csubd(a[r], c);
                a[r] = p; a[r + 4] = m;
            }
#pragma unroll
            for (int r = 0; r < 8; r++) base[swz(i0 + r)] = a[r];
        }
        __syncthreads();
    }
    int s = 3;
    if ((L - 3) & 1) {
        const int half = 1 << s, ppl = n >> 1, np = nlines * ppl;
        for (int b = tid; b < np; b += nt) {
            double2 *base = x + (b / ppl) * n;
            const int g = b % ppl, k = g & (half - 1), i0 = ((g >> s) << (s + 1)) + k;
            const double2 a = base[swz(i0)], c = cmuld(ld_tw<-1>(tw, k << (L - 1 - s)), base[swz(i0 + half)]);
            base[swz(i0)] = caddd(a, c);
            base[swz(i0 + half)] = csubd(a, c);
        }
        __syncthreads();
        s++;
    }
    for (; s + 1 < L; s += 2) {   // stages s and s+1 fused
        const int h = 1 << s, gpl = n >> 2, ng = nlines * gpl;
        for (int b = tid; b < ng; b += nt) {
            double2 *base = x + (b / gpl) * n;
            const int g = b % gpl, k0 = g & (h - 1), i = ((g >> s) << (s + 2)) + k0;
            const double2 w1 = ld_tw<-1>(tw, k0 << (L - 1 - s));
            const double2 w2a = ld_tw<-1>(tw, k0 << (L - 2 - s)), w2b = ld_tw<-1>(tw, (k0 + h) << (L - 2 - s));
            const double2 x0 = base[swz(i)], x1 = cmuld(w1, base[swz(i + h)]), x2 = base[swz(i + 2 * h)], x3 = cmuld(w1, base[swz(i + 3 * h)]);
            const double2 p0 = caddd(x0, x1), m0 = csubd(x0, x1), p1 = caddd(x2, x3), m1 = csubd(x2, x3);
            const double2 c0 = cmuld(w2a, p1), c1 = cmuld(w2b, m1);
            base[swz(i)] = caddd(p0, c0);
            base[swz(i + 2 * h)] = csubd(p0, c0);
            base[swz(i + h)] = caddd(m0, c1);
            base[swz(i + 3 * h)] = csubd(m0, c1);
        }
        __syncthreads();
    }
}

// inverse FFT (unnormalised), decimation in frequency: input natural, output in bit-reversed order
__device__ void fft_dif_fast(double2 *x, int n, int L, int nlines, const double2 *__restrict__ tw) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int s = L - 1;
    for (; s >= 4; s -= 2) {   // stages s and s-1 fused
        const int q = 1 << (s - 1), h = q << 1, gpl = n >> 2, ng = nlines * gpl;
        for (int b = tid; b < ng; b += nt) {
            double2 *base = x + (b / gpl) * n;
            const int g = b % gpl, k0 = g & (q - 1), i = ((g >> (s - 1)) << (s + 1)) + k0;
            const double2 wa = ld_tw<+1>(tw, k0 << (L - 1 - s)), wb = ld_tw<+1>(tw, (k0 + q) << (L - 1 - s)), wc = ld_tw<+1>(tw, k0 << (L - s));
            const double2 x0 = base[swz(i)], x1 = base[swz(i + q)], x2 = base[swz(i + h)], x3 = base[swz(i + h + q)];
            const double2 y0 = caddd(x0, x2), y2 = cmuld(wa, csubd(x0, x2)), y1 = caddd(x1, x3), y3 = cmuld(wb, csubd(x1, x3));
            base[swz(i)] = caddd(y0, y1);
            base[swz(i + q)] = cmuld(wc, csubd(y0, y1));
            base[swz(i + h)] = caddd(y2, y3);
            base[swz(i + h + q)] = cmuld(wc, csubd(y2, y3));
        }
        __syncthreads();
    }
    if (s == 3) {
        const int half = 1 << s, ppl = n >> 1, np = nlines * ppl;
        for (int b = tid; b < np; b += nt) {
            double2 *base = x + (b / ppl) * n;
            const int g = b % ppl, k = g & (half - 1), i0 = ((g >> s) << (s + 1)) + k;
            const double2 a = base[swz(i0)], c = base[swz(i0 + half)];
            base[swz(i0)] = caddd(a, c);
            base[swz(i0 + half)] = cmuld(ld_tw<+1>(tw, k << (L - 1 - s)), csubd(a, c));
        }
        __syncthreads();
        s--;
    }
    {   // stages 2, 1, 0 on aligned blocks of 8
        const int bpl = n >> 3, nblk = nlines * bpl;
        for (int b = tid; b < nblk; b += nt) {
            double2 *base = x + (b / bpl) * n;
            const int i0 = (b % bpl) << 3;
            double2 a[8];
#pragma unroll
            for (int r = 0; r < 8; r++) a[r] = base[swz(i0 + r)];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const double2 p = caddd(a[r], a[r + 4]), m = csubd(a[r], a[r + 4]);
                a[r] = p;
                a[r + 4] = r == 0 ? m : cmuld(ld_tw<+1>(tw, r << (L - 3)), m);
            }
            {
                const double2 w1 = ld_tw<+1>(tw, 1 << (L - 2));
#pragma unroll
                for (int r = 0; r < 8; r += 4) {
                    const double2 p0 = caddd(a[r], a[r + 2]), m0 = csubd(a[r], a[r + 2]), p1 = caddd(a[r + 1], a[r + 3]), m1 = cmuld(w1, csubd(a[r + 1], a[r + 3]));
                    a[r] = p0; a[r + 2] = m0; a[r + 1] = p1; a[r + 3] = m1;
                }
            }
#pragma unroll
            for (int r = 0; r < 8; r += 2) { const double2 p = caddd(a[r], a[r + 1]), m = csubd(a[r], a[r + 1]); a[r] = p; a[r + 1] = m; }
#pragma unroll
            for (int r = 0; r < 8; r++) base[swz(i0 + r)] = a[r];
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int slot_of(int m, int n, int L) { return swz(bitrev(makhoul_pos(m, n), L)); }

// DCT-II post-twiddle of the pair (k, n-k): Z -> (A, B) coefficients of both packed sequences
__device__ __forceinline__ void dct2_post_pair(double2 zk, double2 zn, double2 wk, double2 wn, bool self, double2 &ok, double2 &on) {
    {
        const double var = 0.5 * (zk.x + zn.x), vai = 0.5 * (zk.y - zn.y), vbr = 0.5 * (zk.y + zn.y), vbi = -0.5 * (zk.x - zn.x);
        ok = make_double2(2.0 * (var * wk.x - vai * wk.y), 2.0 * (vbr * wk.x - vbi * wk.y));
    }
    if (!self) {
        const double var = 0.5 * (zn.x + zk.x), vai = 0.5 * (zn.y - zk.y), vbr = 0.5 * (zn.y + zk.y), vbi = -0.5 * (zn.x - zk.x);
        on = make_double2(2.0 * (var * wn.x - vai * wn.y), 2.0 * (vbr * wn.x - vbi * wn.y));
    }
}
// DCT-III pre-twiddle of the pair (j, n-j), j >= 1
__device__ __forceinline__ void dct3_pre_pair(double2 Xj, double2 Xn, double2 qj, double2 qn, bool self, double2 &oj, double2 &on) {
    {
        const double cr = qj.x, ci = -qj.y;
        const double har = Xj.x * cr + Xn.x * ci, hai = Xj.x * ci - Xn.x * cr, hbr = Xj.y * cr + Xn.y * ci, hbi = Xj.y * ci - Xn.y * cr;
        oj = make_double2(har - hbi, hai + hbr);
    }
    if (!self) {
        const double cr = qn.x, ci = -qn.y;
        const double har = Xn.x * cr + Xj.x * ci, hai = Xn.x * ci - Xj.x * cr, hbr = Xn.y * cr + Xj.y * ci, hbi = Xn.y * ci - Xj.y * cr;
        on = make_double2(har - hbi, hai + hbr);
    }
}

// ---- P1: LPC rows per CTA.  rhs = u - tau f, DCT-II along x, spectrum written transposed ------------
template <class R, int LPC>
__global__ void __launch_bounds__(FFT_THREADS) k_cf_rows_fwd(int nx, int ny, const vec2_t<R> *est0, const vec2_t<R> *est1, const vec2_t<R> *__restrict__ gradI,
                                                             const R *__restrict__ It, R tau, double2 *__restrict__ specT, LineTables T, CurvHook H) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *x = reinterpret_cast<double2 *>(smem_raw);   // [LPC][nx]
    const size_t pair_off = (size_t)blockIdx.y * nx * ny;
    const vec2_t<R> *__restrict__ u = est0;
    if (H.enabled) {
        const PairCtl *c = H.ctl + blockIdx.y;
        if (!__ldcg(&c->active)) return;
        u = __ldcg(&c->sel) ? est1 : est0;
    }
    u += pair_off; gradI += pair_off; It += pair_off; specT += pair_off;
    const int j0 = blockIdx.x * LPC, L = T.log2n;
    for (int e = threadIdx.x; e < LPC * nx; e += blockDim.x) {
        const int l = e / nx, i = e - l * nx;
        const size_t g = (size_t)(j0 + l) * nx + i;
        const vec2_t<R> uu = u[g];
        const vec2_t<R> f = lssd_force<R>(gradI[g], It[g], uu);                  // OpticalFlow.cpp:33
        x[l * nx + slot_of(i, nx, L)] = make_double2((double)(uu.x - tau * f.x), (double)(uu.y - tau * f.y));   // OpticalFlowCurvature.cpp:90-91
    }
    __syncthreads();
    fft_dit_fast(x, nx, L, LPC, (const double2 *)T.tw);
    const double2 *__restrict__ q = (const double2 *)T.q;
    const int hp = (nx >> 1) + 1;
    for (int e = threadIdx.x; e < LPC * hp; e += blockDim.x) {
        const int l = e / hp, k = e - l * hp, nk = (nx - k) & (nx - 1);
        double2 *base = x + l * nx;
        double2 ok, on;
        dct2_post_pair(base[swz(k)], base[swz(nk)], q[k], q[nk], nk == k, ok, on);
        base[swz(k)] = ok;
        if (nk != k) base[swz(nk)] = on;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < LPC * nx; e += blockDim.x) {
        const int l = e % LPC, p = e / LPC;
        specT[(size_t)p * ny + j0 + l] = x[l * nx + swz(p)];
    }
}

// ---- P2: one spectrum column (contiguous in spec_T) per CTA: DCT-II along y, eigenvalues, DCT-III along y ----
#ifndef OF2D_COLS_MINB
#define OF2D_COLS_MINB 1
#endif
__global__ void __launch_bounds__(FFT_THREADS, OF2D_COLS_MINB) k_cf_cols(int nx, int ny, double2 *__restrict__ specT, const double *__restrict__ cosx, const double *__restrict__ cosy,
                                                         double tau_alpha, LineTables T, CurvHook H) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *x = reinterpret_cast<double2 *>(smem_raw);   // [ny]
    if (H.enabled && !__ldcg(&H.ctl[blockIdx.y].active)) return;
    const int p = blockIdx.x, L = T.log2n;
    double2 *__restrict__ line = specT + (size_t)blockIdx.y * nx * ny + (size_t)p * ny;
    for (int jj = threadIdx.x; jj < ny; jj += blockDim.x) x[slot_of(jj, ny, L)] = line[jj];
    __syncthreads();
    fft_dit_fast(x, ny, L, 1, (const double2 *)T.tw);
    const double2 *__restrict__ q = (const double2 *)T.q;
    const double cxp = cosx[p];
    for (int k = threadIdx.x; k <= (ny >> 1); k += blockDim.x) {
        const int nk = (ny - k) & (ny - 1);
        const bool self = nk == k;
        double2 ak, an;
        dct2_post_pair(x[swz(k)], x[swz(nk)], q[k], q[nk], self, ak, an);
        {   // OpticalFlowCurvature.cpp:24, :135-136
            const double lap = -4 + cxp + cosy[k];
            const double eig = 1.0f / (1.0f + tau_alpha * (lap * lap));
            ak.x *= eig; ak.y *= eig;
        }
        if (!self) {
            const double lap = -4 + cxp + cosy[nk];
            const double eig = 1.0f / (1.0f + tau_alpha * (lap * lap));
            an.x *= eig; an.y *= eig;
        }
        if (k == 0) {
            x[swz(0)] = ak;                       // h_0 = X_0
        } else if (self) {
            double2 oj, dummy;
            dct3_pre_pair(ak, ak, q[k], q[k], true, oj, dummy);
            x[swz(k)] = oj;
        } else {
            double2 oj, on;
            dct3_pre_pair(ak, an, q[k], q[nk], false, oj, on);
            x[swz(k)] = oj;
            x[swz(nk)] = on;
        }
    }
    __syncthreads();
    fft_dif_fast(x, ny, L, 1, (const double2 *)T.tw);
    for (int jj = threadIdx.x; jj < ny; jj += blockDim.x) line[jj] = x[slot_of(jj, ny, L)];
}

// ---- P3: LPC rows per CTA.  DCT-III along x, u' = rhs / (4 N), Logger epilogue ------------------------
template <class R, int LPC>
__global__ void __launch_bounds__(FFT_THREADS) k_cf_rows_inv(int nx, int ny, const double2 *__restrict__ specT, vec2_t<R> *est0, vec2_t<R> *est1, R fourN, LineTables T,
                                                             CurvHook H) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *x = reinterpret_cast<double2 *>(smem_raw);   // [LPC][nx]
    const int pair = blockIdx.y;
    const size_t pair_off = (size_t)pair * nx * ny;
    vec2_t<R> *__restrict__ unew = est1;
    const vec2_t<R> *__restrict__ uold = est0;
    PairCtl *c = nullptr;
    if (H.enabled) {
        c = H.ctl + pair;
        if (!__ldcg(&c->active)) return;
        if (__ldcg(&c->sel)) { unew = est0; uold = est1; }
    }
    unew += pair_off; uold += pair_off; specT += pair_off;
    const int j0 = blockIdx.x * LPC, L = T.log2n;
    for (int e = threadIdx.x; e < LPC * nx; e += blockDim.x) {
        const int l = e % LPC, p = e / LPC;
        x[l * nx + swz(p)] = specT[(size_t)p * ny + j0 + l];
    }
    __syncthreads();
    const double2 *__restrict__ q = (const double2 *)T.q;
    const int hp = nx >> 1;   // pairs j = 1 .. n/2
    for (int e = threadIdx.x; e < LPC * hp; e += blockDim.x) {
        const int l = e / hp, j = 1 + (e - l * hp), nj = nx - j;
        double2 *base = x + l * nx;
        double2 oj, on;
        dct3_pre_pair(base[swz(j)], base[swz(nj)], q[j], q[nj], nj == j, oj, on);
        base[swz(j)] = oj;
        if (nj != j) base[swz(nj)] = on;
    }
    __syncthreads();
    fft_dif_fast(x, nx, L, LPC, (const double2 *)T.tw);
    double sd = 0.0, sp = 0.0;
    for (int e = threadIdx.x; e < LPC * nx; e += blockDim.x) {
        const int l = e / nx, i = e - l * nx;
        const size_t g = (size_t)(j0 + l) * nx + i;
        const double2 v = x[l * nx + slot_of(i, nx, L)];
        const vec2_t<R> o = mk2<R>((R)v.x / fourN, (R)v.y / fourN);              // OpticalFlowCurvature.cpp:116-117
        unew[g] = o;
        if (H.enabled) {   // Logger.cpp:32-51: prev is the estimate this iteration started from
            const vec2_t<R> old = uold[g];
            const vec2_t<R> df = mk2<R>(o.x - old.x, o.y - old.y);
            sd += sizeof(R) == 4 ? (double)sqrtf((float)(df.x * df.x + df.y * df.y)) : sqrt((double)(df.x * df.x + df.y * df.y));
            sp += sizeof(R) == 4 ? (double)sqrtf((float)(old.x * old.x + old.y * old.y)) : sqrt((double)(old.x * old.x + old.y * old.y));
        }
    }
    if (!H.enabled) return;
    block_sum2(sd, sp);
    const double vals[2] = {sd, sp};
    double *part = H.partials + (size_t)pair * H.pstride;
    if (publish_partials<2>(vals, part, &c->ticket[0], gridDim.x, blockIdx.x)) {
        double out[2];
        reduce_partials<2>(part, gridDim.x, out, 0u, 0u);
        if (threadIdx.x == 0) {
            c->sel ^= 1;
            finalize_logger<R>(c, H.tr, pair, out[0], out[1], (unsigned)(nx * ny), H.n_active);
        }
    }
}

}  // namespace
