"""numpy prototype of the register-blocked (E = 16) in-place FFT used by csrc/fft_reg.cuh: checks the pass structure,
the digit-reversed output order, the DIT inverse, the Makhoul DCT-II/III wrapping and shared-memory bank conflicts
of the XOR swizzle."""
import numpy as np, sys
from scipy.fft import dct

def swz(i):
    x = i >> 3
    return i ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7)

def conflicts(idx):
    """idx: per-lane element index (16 B elements) of one warp access; returns wavefronts (ideal 4)"""
    w = 0
    for q in range(0, 32, 8):
        lanes = [swz(int(i)) for i in idx[q:q+8]]
        groups = {}
        for p in lanes: groups.setdefault(p % 8, set()).add(p)
        w += max(len(v) for v in groups.values())
    return w

def plan(L):
    d3 = L - 8
    return [16, 16, 1 << d3]

def fft_dif(x, L, stats):
    N = 1 << L; E = 16; TPL = N // E
    x = x.copy(); M = N
    for R in plan(L):
        S = M // R
        y = x.copy()
        for c in range(E // R):
            warp_idx = {}
            for u in range(TPL):
                g = u + TPL * c; b, k = divmod(g, S); base = b * M + k
                a = np.array([x[base + S * r] for r in range(R)])
                o = np.fft.fft(a) * np.exp(-2j * np.pi * k * np.arange(R) / M)
                for q in range(R): y[base + S * q] = o[q]
                for r in range(R): warp_idx.setdefault((u // 32, r), []).append(base + S * r)
            for key, idx in warp_idx.items():
                if len(idx) == 32: stats.append(conflicts(idx))
        x = y; M = S
    return x

def fft_dit_inv(x, L):
    N = 1 << L; E = 16; TPL = N // E
    x = x.copy()
    Ms = []; M = N
    for R in plan(L): Ms.append((R, M)); M //= R
    for R, M in reversed(Ms):
        S = M // R
        y = x.copy()
        for c in range(E // R):
            for u in range(TPL):
                g = u + TPL * c; b, k = divmod(g, S); base = b * M + k
                a = np.array([x[base + S * r] for r in range(R)]) * np.exp(+2j * np.pi * k * np.arange(R) / M)
                o = np.fft.ifft(a) * R
                for q in range(R): y[base + S * q] = o[q]
        x = y
    return x

def pos_of(k, L):
    N = 1 << L; S1 = N >> 4; S2 = N >> 8
    return (k & 15) * S1 + ((k >> 4) & 15) * S2 + (k >> 8)

for L in (9, 10, 11, 12):
    N = 1 << L
    rng = np.random.default_rng(L)
    z = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    st = []
    Z = fft_dif(z, L, st)
    ref = np.fft.fft(z)
    P = np.array([pos_of(k, L) for k in range(N)])
    assert sorted(P) == list(range(N))
    e1 = np.max(np.abs(Z[P] - ref))
    back = fft_dit_inv(Z, L)
    e2 = np.max(np.abs(back - N * z))
    # pair-stage access conflicts
    pc = [conflicts([pos_of(k, L) for k in range(k0, k0 + 32)]) for k0 in range(0, N // 2, 32)]
    pc2 = [conflicts([pos_of((N - k) % N, L) for k in range(k0, k0 + 32)]) for k0 in range(0, N // 2, 32)]
    # DCT-II / III via Makhoul
    a = rng.standard_normal(N); b = rng.standard_normal(N)
    v = np.zeros(N, complex)
    m = np.arange(N); mk = np.where(m & 1, N - 1 - (m >> 1), m >> 1)
    v[mk] = a + 1j * b
    Zs = fft_dif(v, L, [])[P]
    k = np.arange(N); nk = (N - k) % N
    Va = 0.5 * (Zs + np.conj(Zs[nk])); Vb = (Zs - np.conj(Zs[nk])) / 2j
    q = np.exp(-1j * np.pi * k / (2 * N))
    A = 2 * np.real(Va * q); B = 2 * np.real(Vb * q)
    e3 = max(np.max(np.abs(A - dct(a, 2))), np.max(np.abs(B - dct(b, 2))))
    # DCT-III of (A, B)
    XA, XB = A, B
    h = np.zeros(N, complex)
    j = np.arange(1, N)
    ha = (XA[j] - 1j * XA[N - j]) * np.conj(q[j]); hb = (XB[j] - 1j * XB[N - j]) * np.conj(q[j])
    h[j] = ha + 1j * hb; h[0] = XA[0] + 1j * XB[0]
    hin = np.zeros(N, complex); hin[P] = h
    t = fft_dit_inv(hin, L)
    oa = np.real(t)[mk]; ob = np.imag(t)[mk]
    e4 = max(np.max(np.abs(oa - dct(XA, 3))), np.max(np.abs(ob - dct(XB, 3))))
    print(L, "fft err", e1, "inv err", e2, "dct2 err", e3, "dct3 err", e4, "pass wavefronts max/mean", max(st), np.mean(st), "pair", max(pc), max(pc2))
