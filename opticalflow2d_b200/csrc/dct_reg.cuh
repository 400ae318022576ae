// dct_reg.cuh -- register-blocked DCT path of the curvature solver for line lengths 512 .. 4096 (included by dct.cu).
//
// Same transform definitions as dct.cu (Makhoul's N-point DCT-II / DCT-III through ONE N-point complex FFT whose real /
// imaginary parts carry the x / y components of the motion), restructured so that shared memory is touched as little as
// the dependency structure allows:
//   * every thread keeps 16 points in registers and a line is transformed in THREE super-passes of radix 16, 16 and
//     N/256 (2 shared-memory exchanges per FFT instead of 5 passes); threads per line = N/16;
//   * forward = decimation in frequency (natural in, digit-reversed out), inverse = decimation in time (digit-reversed
//     in, natural out), so no permutation pass exists: the pair stages (DCT-II post-twiddle, eigenvalues, DCT-III
//     pre-twiddle), which have to visit (k, N-k) anyway, address the digit-reversed slot of k directly;
//   * twiddles of a super-pass: four coalesced table loads (w^k, w^2k, w^4k, w^8k), the other eleven composed in registers
//     (the radix-4 passes of dct_fast.cuh spent more L1 wavefronts on scattered twiddle loads than on the data);
//   * in the first / last super-pass a warp owns the offsets k in [16w, 16w+16) and their mirror images S1-1-k, so that
//     v[m] and v[N-1-m] -- the two halves of a pixel pair (2m, 2m+1) in Makhoul's ordering -- live in lanes l and l^16:
//     one shuffle turns register contents into complete pixel pairs and global loads / stores are full 16-byte accesses;
//   * FMA is allowed inside the FFT (the reference's fftw has no defined operation order to reproduce); the
//     reference's own expressions (force, rhs, 1/(4N), eigenvalue) keep their order and stay unfused.
// Shared-memory indices go through an XOR swizzle (RG_SWZ: dct_fast.cuh's swz for 16-byte elements, swz8 below for 8-byte ones),
// conflict-free for every power-of-two stride; inside the super-passes the swizzled byte offset is (thread part) ^ (immediate), see RG_CS.
// Single-precision instance (RG_PACKED): a complex addition is one packed add.rn.f32x2 / sub.rn.f32x2 (SASS FADD2).
//
// The file is included TWICE by dct.cu: namespace rg (RG_S = double: the reference's precision, used by the strict path, the exact
// engine and fp64 fields) and namespace rgf (RG_S = float: the relaxed engine on fp32 fields -- transform, twiddles and the
// spectrum between the passes in single precision, 60 instead of 92 B/px per iteration and half the registers per point).
#ifndef RG_NS
#define RG_NS rg
#define RG_C2 double2
#define RG_S double
#define RG_MK2 make_double2
#define RG_MINB 2
#define RG_SWZ swz
#endif

namespace {
namespace RG_NS {

__device__ __forceinline__ RG_C2 cmulf(RG_C2 a, RG_C2 b) { return RG_MK2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ RG_C2 cmulcf(RG_C2 a, RG_C2 b) { return RG_MK2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y))); }   // a conj(b)
#ifdef RG_PACKED
// fp32: both components of a complex addition in one packed sm_100a instruction (SASS FADD2)
__device__ __forceinline__ unsigned long long pk2(float2 a) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ float2 upk2(unsigned long long v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ RG_C2 add2(RG_C2 a, RG_C2 b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ RG_C2 sub2(RG_C2 a, RG_C2 b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
// Logger addend of fp32 fields in the relaxed engine: MUFU.SQRT without the subnormal fix-up (as engine_kernels.cuh's sqrt_approx)
__device__ __forceinline__ float lg_sqrtf(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// eigenvalue 1 / (1 + tau alpha lap^2) of the relaxed engine: the argument is >= 1, MUFU.RCP (1 ulp) without the IEEE fix-up
__device__ __forceinline__ float rg_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
__device__ __forceinline__ RG_S rg_rcp(RG_S x) { return (RG_S)1 / x; }
__device__ __forceinline__ RG_C2 add2(RG_C2 a, RG_C2 b) { return RG_MK2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ RG_C2 sub2(RG_C2 a, RG_C2 b) { return RG_MK2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float lg_sqrtf(float x) { return sqrtf(x); }
#endif
template <int SIGN> __device__ __forceinline__ RG_C2 mul_i(RG_C2 a) { return SIGN > 0 ? RG_MK2(-a.y, a.x) : RG_MK2(a.y, -a.x); }   // a (SIGN i)

// a e^{SIGN 2 pi i n / 16}
template <int SIGN, int n> __device__ __forceinline__ RG_C2 mul_w16(RG_C2 a) {
    constexpr int m = n & 15;
    if constexpr (m == 0) return a;
    else if constexpr (m == 4) return mul_i<SIGN>(a);
    else if constexpr (m == 8) return RG_MK2(-a.x, -a.y);
    else if constexpr (m == 12) return mul_i<-SIGN>(a);
    else {
        constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, r = 0.70710678118654752440;
        constexpr double C[16] = {1, c1, r, s1, 0, -s1, -r, -c1, -1, -c1, -r, -s1, 0, s1, r, c1};
        constexpr double S[16] = {0, s1, r, c1, 1, c1, r, s1, 0, -s1, -r, -c1, -1, -c1, -r, -s1};
        constexpr RG_S c = (RG_S)C[m], s = (RG_S)(SIGN * S[m]);
        return RG_MK2(fma(a.x, c, -(a.y * s)), fma(a.x, s, a.y * c));
    }
}

template <int SIGN> __device__ __forceinline__ void radix4(RG_C2 &a0, RG_C2 &a1, RG_C2 &a2, RG_C2 &a3) {
    const RG_C2 p02 = add2(a0, a2), m02 = sub2(a0, a2), p13 = add2(a1, a3), d = mul_i<SIGN>(sub2(a1, a3));
    a0 = add2(p02, p13); a1 = add2(m02, d); a2 = sub2(p02, p13); a3 = sub2(m02, d);
}

// slot of a[] that holds output index q after dft<R>() (and its inverse map; both are involutions for R = 16)
__host__ __device__ constexpr int slot_q(int R, int s) { return R == 16 ? (s >> 2) + 4 * (s & 3) : R == 8 ? (s >> 1) + 4 * (s & 1) : s; }

// unnormalised R-point DFT with kernel e^{SIGN 2 pi i j q / R} of a[OFF .. OFF+R): output q is left in slot OFF + s, q = slot_q(R, s)
template <int R, int SIGN, int OFF> __device__ __forceinline__ void dft(RG_C2 (&a)[16]) {
    if constexpr (R == 2) {
        const RG_C2 p = add2(a[OFF], a[OFF + 1]), m = sub2(a[OFF], a[OFF + 1]);
        a[OFF] = p; a[OFF + 1] = m;
    } else if constexpr (R == 4) {
        radix4<SIGN>(a[OFF], a[OFF + 1], a[OFF + 2], a[OFF + 3]);
    } else if constexpr (R == 8) {   // j = j0 + 2 j1, q = q1 + 4 q0
        radix4<SIGN>(a[OFF + 0], a[OFF + 2], a[OFF + 4], a[OFF + 6]);
        radix4<SIGN>(a[OFF + 1], a[OFF + 3], a[OFF + 5], a[OFF + 7]);
        a[OFF + 3] = mul_w16<SIGN, 2>(a[OFF + 3]); a[OFF + 5] = mul_w16<SIGN, 4>(a[OFF + 5]); a[OFF + 7] = mul_w16<SIGN, 6>(a[OFF + 7]);
#pragma unroll
        for (int q1 = 0; q1 < 4; q1++) { const RG_C2 p = add2(a[OFF + 2 * q1], a[OFF + 2 * q1 + 1]), m = sub2(a[OFF + 2 * q1], a[OFF + 2 * q1 + 1]); a[OFF + 2 * q1] = p; a[OFF + 2 * q1 + 1] = m; }
    } else {                         // j = j0 + 4 j1, q = q1 + 4 q0
        radix4<SIGN>(a[0], a[4], a[8], a[12]);
        radix4<SIGN>(a[1], a[5], a[9], a[13]);
        radix4<SIGN>(a[2], a[6], a[10], a[14]);
        radix4<SIGN>(a[3], a[7], a[11], a[15]);
        a[5] = mul_w16<SIGN, 1>(a[5]); a[9] = mul_w16<SIGN, 2>(a[9]); a[13] = mul_w16<SIGN, 3>(a[13]);
        a[6] = mul_w16<SIGN, 2>(a[6]); a[10] = mul_w16<SIGN, 4>(a[10]); a[14] = mul_w16<SIGN, 6>(a[14]);
        a[7] = mul_w16<SIGN, 3>(a[7]); a[11] = mul_w16<SIGN, 6>(a[11]); a[15] = mul_w16<SIGN, 9>(a[15]);
        radix4<SIGN>(a[0], a[1], a[2], a[3]);
        radix4<SIGN>(a[4], a[5], a[6], a[7]);
        radix4<SIGN>(a[8], a[9], a[10], a[11]);
        radix4<SIGN>(a[12], a[13], a[14], a[15]);
    }
}

// a[idx(q)] *= w^q (CONJ: conj(w)^q), q = 1..15, from w, w^2, w^4, w^8; SLOT: a[] is in dft<16> output order
template <bool SLOT, bool CONJ> __device__ __forceinline__ void twiddle16(RG_C2 (&a)[16], RG_C2 w1, RG_C2 w2, RG_C2 w4, RG_C2 w8) {
#define RG_AT(q) a[SLOT ? slot_q(16, (q)) : (q)]
#define RG_MUL(q, w) RG_AT(q) = CONJ ? cmulcf(RG_AT(q), w) : cmulf(RG_AT(q), w)
    RG_MUL(1, w1); RG_MUL(2, w2); RG_MUL(4, w4); RG_MUL(8, w8);
    const RG_C2 w3 = cmulf(w1, w2), w5 = cmulf(w1, w4), w6 = cmulf(w2, w4), w12 = cmulf(w4, w8);
    RG_MUL(3, w3); RG_MUL(5, w5); RG_MUL(6, w6); RG_MUL(12, w12);
    const RG_C2 w7 = cmulf(w3, w4);
    RG_MUL(7, w7);
    { const RG_C2 w9 = cmulf(w1, w8); RG_MUL(9, w9); }
    { const RG_C2 w10 = cmulf(w2, w8); RG_MUL(10, w10); }
    { const RG_C2 w11 = cmulf(w3, w8); RG_MUL(11, w11); }
    { const RG_C2 w13 = cmulf(w5, w8); RG_MUL(13, w13); }
    { const RG_C2 w14 = cmulf(w6, w8); RG_MUL(14, w14); }
    { const RG_C2 w15 = cmulf(w7, w8); RG_MUL(15, w15); }
#undef RG_MUL
#undef RG_AT
}

template <int L> struct Geo {
    static constexpr int N = 1 << L, S1 = N >> 4, S2 = N >> 8, M3 = N >> 8, TPL = N >> 4;
    static_assert(L >= 9 && L <= 12, "register path: 512 <= N <= 4096");
    // digit-reversed slot of frequency k
    __device__ static __forceinline__ int pos(int k) { return (k & 15) * S1 + ((k >> 4) & 15) * S2 + (k >> 8); }
    // offset handled by thread u of a line in the first / last super-pass: lanes l and l^16 hold k and S1-1-k
    __device__ static __forceinline__ int k1(int u) {
        const int lane = u & 31, w = u >> 5;
        return lane < 16 ? 16 * w + lane : S1 - 1 - 16 * w - (lane - 16);
    }
};

// Shared-memory addressing of the super-passes.  RG_SWZ is linear over GF(2) (shifts, XORs, a mask), and in every access below the
// index is A + B with the thread's part A and the unrolled loop's constant B on DISJOINT bits, so RG_SWZ(A + B) = RG_SWZ(A) ^ RG_SWZ(B):
// the thread's byte offset is formed once per pass and every access costs one XOR with an immediate.
constexpr unsigned ES = sizeof(RG_C2);
#define RG_CS(i) ((unsigned)RG_SWZ(i) * ES)
__device__ __forceinline__ RG_C2 &sm_at(RG_C2 *x, unsigned byte_off) { return *reinterpret_cast<RG_C2 *>(reinterpret_cast<unsigned char *>(x) + byte_off); }

struct Tw16 { const RG_C2 *a, *b; };   // [4][S1] for super-pass 1 (w_N), [4][S2] for super-pass 2 (w_{N/16}); forward sign

// ---- forward (DIF): registers -> ... -> shared memory, digit-reversed --------------------------------------------
// in: a[j] = v[k + S1 j] (natural order); the threads of a line hold k1 = a permutation of 0 .. S1-1 and u = 0 .. S1-1.
// out: X_k at xl[RG_SWZ(pos(k))], xl = the line at byte offset lo of x, after the trailing barrier.
template <int L> __device__ __forceinline__ void fft_fwd(RG_C2 (&a)[16], RG_C2 *x, unsigned lo, int u, int k1, Tw16 T) {
    using G = Geo<L>;
    {
        const int k = k1;
        dft<16, -1, 0>(a);
        twiddle16<true, false>(a, T.a[k], T.a[G::S1 + k], T.a[2 * G::S1 + k], T.a[3 * G::S1 + k]);
        const unsigned t = lo + RG_CS(k);
#pragma unroll
        for (int s = 0; s < 16; s++) sm_at(x, t ^ RG_CS(G::S1 * slot_q(16, s))) = a[s];
    }
    __syncthreads();
    {
        const int b = u / G::S2, k = u % G::S2, base = b * G::S1 + k;
        const unsigned t = lo + RG_CS(base);
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = sm_at(x, t ^ RG_CS(G::S2 * j));
        dft<16, -1, 0>(a);
        twiddle16<true, false>(a, T.b[k], T.b[G::S2 + k], T.b[2 * G::S2 + k], T.b[3 * G::S2 + k]);
#pragma unroll
        for (int s = 0; s < 16; s++) sm_at(x, t ^ RG_CS(G::S2 * slot_q(16, s))) = a[s];
    }
    __syncthreads();
    {
        const unsigned t = lo + RG_CS(16 * u);
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = sm_at(x, t ^ RG_CS(j));
        if constexpr (G::M3 == 16) dft<16, -1, 0>(a);
        if constexpr (G::M3 == 8) { dft<8, -1, 0>(a); dft<8, -1, 8>(a); }
        if constexpr (G::M3 == 4) { dft<4, -1, 0>(a); dft<4, -1, 4>(a); dft<4, -1, 8>(a); dft<4, -1, 12>(a); }
        if constexpr (G::M3 == 2) { dft<2, -1, 0>(a); dft<2, -1, 2>(a); dft<2, -1, 4>(a); dft<2, -1, 6>(a); dft<2, -1, 8>(a); dft<2, -1, 10>(a); dft<2, -1, 12>(a); dft<2, -1, 14>(a); }
#pragma unroll
        for (int s = 0; s < 16; s++) sm_at(x, t ^ RG_CS((s & ~(G::M3 - 1)) + slot_q(G::M3, s & (G::M3 - 1)))) = a[s];
    }
    __syncthreads();
}

// ---- inverse (DIT, unnormalised): shared memory, digit-reversed -> ... -> registers -------------------------------
// in: h_k at xl[RG_SWZ(pos(k))] (caller has synchronised).  out: a[s] = t[k1 + S1 q], q = slot_q(16, s).
// `before_last_barrier` runs while a[] is dead (prefetches of the epilogue go there).
struct NoHook { __device__ __forceinline__ void operator()() const {} };
template <int L, class Hook = NoHook> __device__ __forceinline__ void fft_inv(RG_C2 (&a)[16], RG_C2 *x, unsigned lo, int u, int k1, Tw16 T, Hook before_last_barrier = Hook()) {
    using G = Geo<L>;
    {
        const unsigned t = lo + RG_CS(16 * u);
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = sm_at(x, t ^ RG_CS(j));
        if constexpr (G::M3 == 16) dft<16, +1, 0>(a);
        if constexpr (G::M3 == 8) { dft<8, +1, 0>(a); dft<8, +1, 8>(a); }
        if constexpr (G::M3 == 4) { dft<4, +1, 0>(a); dft<4, +1, 4>(a); dft<4, +1, 8>(a); dft<4, +1, 12>(a); }
        if constexpr (G::M3 == 2) { dft<2, +1, 0>(a); dft<2, +1, 2>(a); dft<2, +1, 4>(a); dft<2, +1, 6>(a); dft<2, +1, 8>(a); dft<2, +1, 10>(a); dft<2, +1, 12>(a); dft<2, +1, 14>(a); }
#pragma unroll
        for (int s = 0; s < 16; s++) sm_at(x, t ^ RG_CS((s & ~(G::M3 - 1)) + slot_q(G::M3, s & (G::M3 - 1)))) = a[s];
    }
    __syncthreads();
    {
        const int b = u / G::S2, k = u % G::S2, base = b * G::S1 + k;
        const unsigned t = lo + RG_CS(base);
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = sm_at(x, t ^ RG_CS(G::S2 * j));
        twiddle16<false, true>(a, T.b[k], T.b[G::S2 + k], T.b[2 * G::S2 + k], T.b[3 * G::S2 + k]);
        dft<16, +1, 0>(a);
#pragma unroll
        for (int s = 0; s < 16; s++) sm_at(x, t ^ RG_CS(G::S2 * slot_q(16, s))) = a[s];
    }
    before_last_barrier();
    __syncthreads();
    {
        const int k = k1;
        const unsigned t = lo + RG_CS(k);
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = sm_at(x, t ^ RG_CS(G::S1 * j));
        twiddle16<false, true>(a, T.a[k], T.a[G::S1 + k], T.a[2 * G::S1 + k], T.a[3 * G::S1 + k]);
        dft<16, +1, 0>(a);
    }
}

// DCT-II post-twiddle of the pair (k, N-k) of the packed FFT output (same algebra as dct2_post_pair with the factors
// 1/2 and 2 cancelled and q_{N-k} = -i conj(q_k)): ok = (A_k, B_k), on = (A_{N-k}, B_{N-k}).  zn = zk for k = 0, N/2.
__device__ __forceinline__ void post_pair(RG_C2 zk, RG_C2 zn, RG_C2 q, RG_C2 &ok, RG_C2 &on) {
    const RG_S sx = zk.x + zn.x, sy = zk.y - zn.y, dx = zk.x - zn.x, dy = zk.y + zn.y;
    ok = RG_MK2(fma(sx, q.x, -(sy * q.y)), fma(dy, q.x, dx * q.y));
    on = RG_MK2(-fma(sx, q.y, sy * q.x), fma(dx, q.x, -(dy * q.y)));
}
// DCT-III pre-twiddle of the pair (j, N-j), j >= 1: h_j = (X_j - i X_{N-j}) conj(q_j) for both packed sequences
__device__ __forceinline__ void pre_pair(RG_C2 Xj, RG_C2 Xn, RG_C2 q, RG_C2 &oj, RG_C2 &on) {
    const RG_S har = fma(Xj.x, q.x, -(Xn.x * q.y)), hai = -fma(Xj.x, q.y, Xn.x * q.x);
    const RG_S hbr = fma(Xj.y, q.x, -(Xn.y * q.y)), hbi = -fma(Xj.y, q.y, Xn.y * q.x);
    oj = RG_MK2(har - hbi, hai + hbr);
    on = RG_MK2(har + hbi, hbr - hai);
}

template <class R> __device__ __forceinline__ void load_px_pair(const vec2_t<R> *p, vec2_t<R> &a, vec2_t<R> &b) { a = p[0]; b = p[1]; }
template <> __device__ __forceinline__ void load_px_pair<float>(const float2 *p, float2 &a, float2 &b) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    a = make_float2(v.x, v.y); b = make_float2(v.z, v.w);
}
template <class R> __device__ __forceinline__ void store_px_pair(vec2_t<R> *p, vec2_t<R> a, vec2_t<R> b) { p[0] = a; p[1] = b; }
template <> __device__ __forceinline__ void store_px_pair<float>(float2 *p, float2 a, float2 b) { *reinterpret_cast<float4 *>(p) = make_float4(a.x, a.y, b.x, b.y); }
template <class R> __device__ __forceinline__ void load_s_pair(const R *p, R &a, R &b) { a = p[0]; b = p[1]; }
template <> __device__ __forceinline__ void load_s_pair<float>(const float *p, float &a, float &b) { const float2 v = *reinterpret_cast<const float2 *>(p); a = v.x; b = v.y; }

// ---- P1: LPC rows per CTA.  rhs = u - tau f, DCT-II along x, spectrum written transposed -----------------------------
template <class R, int L, int LPC>
__global__ void __launch_bounds__(LPC * Geo<L>::TPL, LPC * Geo<L>::TPL <= 256 ? RG_MINB : 1) k_rg_rows_fwd(int ny, const vec2_t<R> *est0, const vec2_t<R> *est1, const vec2_t<R> *__restrict__ gradI,
                                                                   const R *__restrict__ It, R tau, RG_C2 *__restrict__ specT, const RG_C2 *__restrict__ q, Tw16 T,
                                                                   CurvHook H) {
    pdl_enter();
    using G = Geo<L>;
    constexpr int N = G::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RG_C2 *x = reinterpret_cast<RG_C2 *>(smem_raw);   // [LPC][N]
    const size_t pair_off = (size_t)blockIdx.y * N * ny;
    const vec2_t<R> *__restrict__ uin = est0;
    if (H.enabled) {
        const PairCtl *c = H.ctl + blockIdx.y;
        if (!__ldcg(&c->active)) return;
        uin = __ldcg(&c->sel) ? est1 : est0;
    }
    uin += pair_off; gradI += pair_off; It += pair_off; specT += pair_off;
    const int tid = threadIdx.x, l = tid / G::TPL, u = tid % G::TPL, j0 = blockIdx.x * LPC;
    RG_C2 a[16];
    {   // pixel pairs (2m, 2m+1), m = k + S1 j (j < 8): the even pixel is v[m] (mine), the odd one v[N-1-m] (lane ^ 16, element 15-j)
        const int k = G::k1(u);
        const size_t row = (size_t)(j0 + l) * N;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const size_t g = row + 2 * (k + G::S1 * j);
            vec2_t<R> u0, u1, g0, g1; R t0, t1;
            load_px_pair<R>(uin + g, u0, u1); load_px_pair<R>(gradI + g, g0, g1); load_s_pair<R>(It + g, t0, t1);
            const vec2_t<R> f0 = lssd_force<R>(g0, t0, u0), f1 = lssd_force<R>(g1, t1, u1);                  // OpticalFlow.cpp:33
            const R ex = u0.x - tau * f0.x, ey = u0.y - tau * f0.y, ox = u1.x - tau * f1.x, oy = u1.y - tau * f1.y;   // OpticalFlowCurvature.cpp:90-91
            a[j] = RG_MK2((RG_S)ex, (RG_S)ey);
            a[15 - j] = RG_MK2((RG_S)__shfl_xor_sync(0xffffffffu, ox, 16), (RG_S)__shfl_xor_sync(0xffffffffu, oy, 16));
        }
    }
    fft_fwd<L>(a, x, (unsigned)(l * N) * ES, u, G::k1(u), T);
    constexpr int hp = (N >> 1) + 1;
    for (int e = tid; e < LPC * hp; e += LPC * G::TPL) {
        const int ll = e % LPC, k = e / LPC, nk = (N - k) & (N - 1);
        const RG_C2 *base = x + ll * N;
        RG_C2 ok, on;
        post_pair(base[RG_SWZ(G::pos(k))], base[RG_SWZ(G::pos(nk))], q[k], ok, on);
        specT[(size_t)k * ny + j0 + ll] = ok;
        if (nk != k) specT[(size_t)nk * ny + j0 + ll] = on;
    }
}

// ---- P2: two spectrum columns (contiguous lines of spec_T) per CTA: DCT-II along y, eigenvalues, DCT-III along y ------
// Lanes 0-15 of every warp work on the first line, lanes 16-31 on the second (quarter-warps never mix lines, so the
// shared-memory phases stay conflict-free), and the result goes out in the NATURAL layout spec_N[y][p]: lane pairs
// (l, l+16) fill whole 32-byte sectors, so P3 reads contiguous rows and all scattered traffic of the iteration is writes.
template <int L>
__global__ void __launch_bounds__(2 * Geo<L>::TPL, 2 * Geo<L>::TPL <= 256 ? RG_MINB : 1) k_rg_cols(int nx, const RG_C2 *__restrict__ specT, RG_C2 *__restrict__ specN,
                                                                                        const RG_S *__restrict__ cosx, const RG_S *__restrict__ cosy, double tau_alpha,
                                                                                        const RG_C2 *__restrict__ q, Tw16 T, CurvHook H) {
    pdl_enter();
    using G = Geo<L>;
    constexpr int N = G::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RG_C2 *x = reinterpret_cast<RG_C2 *>(smem_raw);   // [2][N]
    if (H.enabled && !__ldcg(&H.ctl[blockIdx.y].active)) return;
    const int tid = threadIdx.x, lane = tid & 31, l = lane >> 4, u = (tid >> 5) * 16 + (lane & 15), p = blockIdx.x * 2 + l;
    RG_C2 *xl = x + l * N;
    const RG_C2 *__restrict__ line = specT + (size_t)blockIdx.y * nx * N + (size_t)p * N;
    specN += (size_t)blockIdx.y * nx * N;
    RG_C2 a[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int m = u + G::S1 * j;
        a[j] = line[j < 8 ? 2 * m : 2 * (N - 1 - m) + 1];
    }
    fft_fwd<L>(a, x, (unsigned)(l * N) * ES, u, u, T);
    const RG_S cxp = (RG_S)cosx[p], ta = (RG_S)tau_alpha;   // (rgf: the eigenvalue in single precision as well)
#pragma unroll 4
    for (int c = 0; c < 8; c++) {   // pairs (k, N-k), k = 1 .. N/2-1 (and the self pair k = 0: thread 0, c = 0)
        const int k = u + G::TPL * c, nk = (N - k) & (N - 1);
        const int sk = RG_SWZ(G::pos(k)), sn = RG_SWZ(G::pos(nk));
        const RG_C2 qk = q[k];
        RG_C2 ak, an;
        post_pair(xl[sk], xl[sn], qk, ak, an);
        // OpticalFlowCurvature.cpp:24, :135-136: 1 / (1 + tau alpha lap^2) for k and N-k, with one division for the two
        const RG_S lk = -4 + cxp + (RG_S)cosy[k], ln = -4 + cxp + (RG_S)cosy[nk];
        const RG_S dk = 1.0f + ta * (lk * lk), dn = 1.0f + ta * (ln * ln);
        const RG_S r = rg_rcp(dk * dn), ek = r * dn, en = r * dk;
        { const RG_S fk = (RG_S)ek, fn = (RG_S)en; ak.x *= fk; ak.y *= fk; an.x *= fn; an.y *= fn; }
        RG_C2 oj, on;
        pre_pair(ak, an, qk, oj, on);
        if (k == 0) {                             // h_0 = X_0
            xl[sk] = ak;
        } else {
            xl[sk] = oj;
            xl[sn] = on;
        }
    }
    if (u == 0) {                                 // self pair k = N/2
        constexpr int k = N >> 1;
        const int sk = RG_SWZ(G::pos(k));
        const RG_C2 qk = q[k], z = xl[sk];
        RG_C2 ak, an;
        post_pair(z, z, qk, ak, an);
        const RG_S lk = -4 + cxp + (RG_S)cosy[k];
        const RG_S ek = rg_rcp(1.0f + ta * (lk * lk));
        { const RG_S fk = (RG_S)ek; ak.x *= fk; ak.y *= fk; }
        RG_C2 oj, on;
        pre_pair(ak, ak, qk, oj, on);
        xl[sk] = oj;
    }
    __syncthreads();
    fft_inv<L>(a, x, (unsigned)(l * N) * ES, u, u, T);
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int qq = slot_q(16, s), m = u + G::S1 * qq, y = qq < 8 ? 2 * m : 2 * (N - 1 - m) + 1;
        specN[(size_t)y * nx + p] = a[s];
    }
}

// ---- P3: LPC rows per CTA.  DCT-III along x, u' = rhs / (4 N), Logger epilogue ------------------------------------------
// FUSE: the kernel goes on with P1 of the NEXT iteration on the same rows -- rhs = u' - tau f(u') from the registers that
// hold u' (same pixel-pair mapping as k_rg_rows_fwd), DCT-II along x, spectrum written transposed -- so an iteration is
// two launches (columns, rows) and u' is not read back.  The forward part runs whatever the Logger decides: if the
// loop ends here its spectrum is simply never used.
template <class R, int L, int LPC, bool FUSE>
__global__ void __launch_bounds__(LPC * Geo<L>::TPL, LPC * Geo<L>::TPL <= 256 ? RG_MINB : 1) k_rg_rows_inv(int ny, const RG_C2 *__restrict__ specN, vec2_t<R> *est0, vec2_t<R> *est1, R fourN,
                                                                   const RG_C2 *__restrict__ q, Tw16 T, CurvHook H, const vec2_t<R> *__restrict__ gradI,
                                                                   const R *__restrict__ It, R tau, RG_C2 *__restrict__ specT) {
    pdl_enter();
    using G = Geo<L>;
    constexpr int N = G::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RG_C2 *x = reinterpret_cast<RG_C2 *>(smem_raw);   // [LPC][N]
    const int pair = blockIdx.y;
    const size_t pair_off = (size_t)pair * N * ny;
    vec2_t<R> *__restrict__ unew = est1;
    const vec2_t<R> *__restrict__ uold = est0;
    PairCtl *c = nullptr;
    if (H.enabled) {
        c = H.ctl + pair;
        if (!__ldcg(&c->active)) return;
        if (__ldcg(&c->sel)) { unew = est0; uold = est1; }
    }
    unew += pair_off; uold += pair_off; specN += pair_off;
    const int tid = threadIdx.x, l = tid / G::TPL, u = tid % G::TPL, j0 = blockIdx.x * LPC;
    RG_C2 a[16];
    {   // pairs (j, N-j), j = 1 .. N/2-1 of each line (contiguous in spec_N): 8 per thread, all 16 loads issued before the first use
        constexpr int NT = LPC * G::TPL, per_line = (N >> 1) - 1, items = LPC * per_line;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int e = tid + NT * c, ll = e / per_line, j = 1 + e % per_line;
            if (e < items) {
                const RG_C2 *ln = specN + (size_t)(j0 + ll) * N;
                a[2 * c] = ln[j];
                a[2 * c + 1] = ln[N - j];
            }
        }
        if (tid < 2 * LPC) {                                          // j = 0 (h_0 = X_0; pos(0) = 0 = RG_SWZ(0)) and the self pair j = N/2
            const int ll = tid % LPC, j = (tid / LPC) * (N >> 1);
            const RG_C2 Xj = specN[(size_t)(j0 + ll) * N + j];
            RG_C2 oj = Xj, on;
            if (j) pre_pair(Xj, Xj, q[j], oj, on);
            x[ll * N + RG_SWZ(G::pos(j))] = oj;
        }
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int e = tid + NT * c, ll = e / per_line, j = 1 + e % per_line;
            if (e < items) {
                RG_C2 oj, on;
                pre_pair(a[2 * c], a[2 * c + 1], q[j], oj, on);
                x[ll * N + RG_SWZ(G::pos(j))] = oj;
                x[ll * N + RG_SWZ(G::pos(N - j))] = on;
            }
        }
    }
    __syncthreads();
    const int k = G::k1(u);
    const size_t row = (size_t)(j0 + l) * N;
    // Logger.cpp:32-51: prev is the estimate this iteration started from; fp32 fields: fetched while a[] is dead
    vec2_t<R> pv[16];
    constexpr bool PREFETCH = sizeof(R) == 4;
    auto fetch_prev = [&]() {
        if (H.enabled) {
#pragma unroll
            for (int qq = 0; qq < 8; qq++) load_px_pair<R>(uold + row + 2 * (k + G::S1 * qq), pv[2 * qq], pv[2 * qq + 1]);
        }
    };
    if constexpr (PREFETCH) fft_inv<L>(a, x, (unsigned)(l * N) * ES, u, k, T, fetch_prev);
    else { fft_inv<L>(a, x, (unsigned)(l * N) * ES, u, k, T); fetch_prev(); }
    double sd = 0.0, sp = 0.0;
    float fsd = 0.0f, fsp = 0.0f;
    // 4 nx ny is a power of two on this path (both sizes are): multiplying by its reciprocal rounds exactly like the reference's division
    const R inv4N = (R)1 / fourN;
    vec2_t<R> ue[FUSE ? 8 : 1], uo[FUSE ? 8 : 1];   // FUSE: the new estimate at the thread's 8 pixel pairs
    {
#pragma unroll
        for (int qq = 0; qq < 8; qq++) {
            // t[m] (mine, m = k + S1 qq) is pixel 2m; pixel 2m+1 is t[N-1-m] = element 15-qq of lane ^ 16
            const RG_C2 ve = a[slot_q(16, qq)], vs = a[slot_q(16, 15 - qq)];
            const vec2_t<R> o0 = mk2<R>((R)ve.x * inv4N, (R)ve.y * inv4N);                      // OpticalFlowCurvature.cpp:116-117
            const vec2_t<R> os = mk2<R>((R)vs.x * inv4N, (R)vs.y * inv4N);
            const vec2_t<R> o1 = mk2<R>(__shfl_xor_sync(0xffffffffu, os.x, 16), __shfl_xor_sync(0xffffffffu, os.y, 16));
            store_px_pair<R>(unew + row + 2 * (k + G::S1 * qq), o0, o1);
            if (FUSE) { ue[qq] = o0; uo[qq] = o1; }
            if (H.enabled) {
                const vec2_t<R> p0 = pv[2 * qq], p1 = pv[2 * qq + 1];
                const vec2_t<R> d0 = mk2<R>(o0.x - p0.x, o0.y - p0.y), d1 = mk2<R>(o1.x - p1.x, o1.y - p1.y);
                if (sizeof(R) == 4) {
#ifdef RG_PACKED   // relaxed engine: the thread's 16 addends summed in single precision, widened once (below)
                    fsd += lg_sqrtf((float)(d0.x * d0.x + d0.y * d0.y)) + lg_sqrtf((float)(d1.x * d1.x + d1.y * d1.y));
                    fsp += lg_sqrtf((float)(p0.x * p0.x + p0.y * p0.y)) + lg_sqrtf((float)(p1.x * p1.x + p1.y * p1.y));
#else
                    sd += (double)lg_sqrtf((float)(d0.x * d0.x + d0.y * d0.y)) + (double)lg_sqrtf((float)(d1.x * d1.x + d1.y * d1.y));
                    sp += (double)lg_sqrtf((float)(p0.x * p0.x + p0.y * p0.y)) + (double)lg_sqrtf((float)(p1.x * p1.x + p1.y * p1.y));
#endif
                } else {
                    sd += sqrt((double)(d0.x * d0.x + d0.y * d0.y)) + sqrt((double)(d1.x * d1.x + d1.y * d1.y));
                    sp += sqrt((double)(p0.x * p0.x + p0.y * p0.y)) + sqrt((double)(p1.x * p1.x + p1.y * p1.y));
                }
            }
        }
    }
    sd += (double)fsd; sp += (double)fsp;
    if constexpr (FUSE) {
        // P1 of the next iteration (k_rg_rows_fwd) on the new estimate held in registers
        gradI += pair_off; It += pair_off; specT += pair_off;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const size_t g = row + 2 * (k + G::S1 * j);
            vec2_t<R> g0, g1; R t0, t1;
            load_px_pair<R>(gradI + g, g0, g1); load_s_pair<R>(It + g, t0, t1);
            const vec2_t<R> u0 = ue[j], u1 = uo[j];
            const vec2_t<R> f0 = lssd_force<R>(g0, t0, u0), f1 = lssd_force<R>(g1, t1, u1);                  // OpticalFlow.cpp:33
            const R ex = u0.x - tau * f0.x, ey = u0.y - tau * f0.y, ox = u1.x - tau * f1.x, oy = u1.y - tau * f1.y;   // OpticalFlowCurvature.cpp:90-91
            a[j] = RG_MK2((RG_S)ex, (RG_S)ey);
            a[15 - j] = RG_MK2((RG_S)__shfl_xor_sync(0xffffffffu, ox, 16), (RG_S)__shfl_xor_sync(0xffffffffu, oy, 16));
        }
        __syncthreads();   // every thread has left the last shared-memory phase of the inverse transform
        fft_fwd<L>(a, x, (unsigned)(l * N) * ES, u, k, T);
        constexpr int hp = (N >> 1) + 1;
        for (int e = tid; e < LPC * hp; e += LPC * G::TPL) {
            const int ll = e % LPC, kk = e / LPC, nk = (N - kk) & (N - 1);
            const RG_C2 *base = x + ll * N;
            RG_C2 ok, on;
            post_pair(base[RG_SWZ(G::pos(kk))], base[RG_SWZ(G::pos(nk))], q[kk], ok, on);
            specT[(size_t)kk * ny + j0 + ll] = ok;
            if (nk != kk) specT[(size_t)nk * ny + j0 + ll] = on;
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
            finalize_logger<R>(c, H.tr, pair, out[0], out[1], (unsigned)(N * ny), H.n_active);
        }
    }
}

// handle for the host launch code (dct.cu), which is written once for both instantiations of this file
struct Api {
    using C2 = RG_C2;
    using S = RG_S;
    using Tw = Tw16;
    template <int L> using G = Geo<L>;
    template <class R, int L, int LPC> static constexpr auto rows_fwd() { return &k_rg_rows_fwd<R, L, LPC>; }
    template <class R, int L, int LPC, bool FUSE> static constexpr auto rows_inv() { return &k_rg_rows_inv<R, L, LPC, FUSE>; }
    template <int L> static constexpr auto cols() { return &k_rg_cols<L>; }
};

}  // namespace RG_NS
}  // namespace

#undef RG_NS
#undef RG_C2
#undef RG_S
#undef RG_MK2
#undef RG_MINB
#undef RG_SWZ
#undef RG_PACKED
#undef RG_CS
