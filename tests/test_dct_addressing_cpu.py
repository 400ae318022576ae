"""Host-side check of the shared-memory addressing the register-blocked DCT kernels rely on (opticalflow2d_b200/csrc/dct_reg.cuh,
fft_fwd / fft_inv): every access of a super-pass forms its offset as swizzle(thread part) XOR swizzle(loop constant).  That is the
swizzled offset of (thread part + loop constant) only because (1) the swizzles are linear over GF(2) and (2) the two parts sit on
disjoint bits.  Both facts are restated here from the kernel source and checked exhaustively for the line lengths the path serves
(512 .. 4096), together with the bank-conflict freedom of the three super-passes under the kernels' thread mappings.
No GPU, no oracle: pure index arithmetic."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "opticalflow2d_b200", "csrc")


def swz16(i):   # dct_fast.cuh: swizzle of 16-byte elements (double2; namespace rg)
    x = i >> 3
    return i ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7)


def swz8(i):    # dct.cu: swizzle of 8-byte elements (float2; namespace rgf)
    x = i >> 4
    return i ^ ((x ^ (x >> 4) ^ (x >> 8)) & 15)


def slot_q(R, s):   # dct_reg.cuh
    return (s >> 2) + 4 * (s & 3) if R == 16 else (s >> 1) + 4 * (s & 1) if R == 8 else s


def test_python_restatement_matches_the_source():
    fast = open(os.path.join(CSRC, "dct_fast.cuh")).read()
    dct = open(os.path.join(CSRC, "dct.cu")).read()
    reg = open(os.path.join(CSRC, "dct_reg.cuh")).read()
    assert re.search(r"int swz\(int i\) \{\s*const int x = i >> 3;\s*return i \^ \(\(x \^ \(x >> 3\) \^ \(x >> 6\) \^ \(x >> 9\)\) & 7\);", fast)
    assert re.search(r"int swz8\(int i\) \{\s*const int x = i >> 4;\s*return i \^ \(\(x \^ \(x >> 4\) \^ \(x >> 8\)\) & 15\);", dct)
    assert "return R == 16 ? (s >> 2) + 4 * (s & 3) : R == 8 ? (s >> 1) + 4 * (s & 1) : s;" in reg
    # the access forms this file enumerates
    for form in ("lo + RG_CS(k)", "t ^ RG_CS(G::S1 * slot_q(16, s))", "lo + RG_CS(base)", "t ^ RG_CS(G::S2 * j)", "t ^ RG_CS(G::S2 * slot_q(16, s))",
                 "lo + RG_CS(16 * u)", "t ^ RG_CS(j)", "t ^ RG_CS((s & ~(G::M3 - 1)) + slot_q(G::M3, s & (G::M3 - 1)))", "t ^ RG_CS(G::S1 * j)"):
        assert form in reg, form


@pytest.mark.parametrize("swz", [swz8, swz16])
def test_swizzles_are_linear_bijections(swz):
    i = np.arange(4096)
    s = swz(i)
    assert sorted(s.tolist()) == list(range(4096))                      # a permutation of the line's slots ...
    for n in (512, 1024, 2048, 4096):
        assert s[:n].max() == n - 1                                      # ... that stays inside every power-of-two line
    a, b = np.meshgrid(i, i, indexing="ij")
    assert np.array_equal(swz(a ^ b), swz(a) ^ swz(b))                   # linear over GF(2)


def passes(L):
    """(thread part A(u, k1), loop constants B) of every super-pass access, as fft_fwd / fft_inv form them"""
    N = 1 << L
    S1, S2, M3 = N >> 4, N >> 8, N >> 8
    p1 = (lambda u, k1: k1, [S1 * slot_q(16, s) for s in range(16)] + [S1 * j for j in range(16)])
    p2 = (lambda u, k1: (u // S2) * S1 + u % S2, [S2 * j for j in range(16)] + [S2 * slot_q(16, s) for s in range(16)])
    p3 = (lambda u, k1: 16 * u, list(range(16)) + [(s & ~(M3 - 1)) + slot_q(M3, s & (M3 - 1)) for s in range(16)])
    return N, S1, [p1, p2, p3]


def k1_rows(u, S1):   # Geo<L>::k1: lanes l and l ^ 16 hold k and S1 - 1 - k
    lane, w = u & 31, u >> 5
    return 16 * w + lane if lane < 16 else S1 - 1 - 16 * w - (lane - 16)


@pytest.mark.parametrize("L", [9, 10, 11, 12])
@pytest.mark.parametrize("swz", [swz8, swz16])
def test_thread_part_and_loop_constant_are_disjoint_and_the_xor_form_is_exact(L, swz):
    N, S1, P = passes(L)
    for A, Bs in P:
        for u in range(S1):                       # threads per line = N / 16 = S1
            for k1 in (u, k1_rows(u, S1)):        # column kernel: k1 = u; row kernels: the mirrored mapping
                a = A(u, k1)
                for b in Bs:
                    assert a & b == 0 and a + b < N
                    assert swz(a + b) == swz(a) ^ swz(b)


def _conflict_degree(slots, group):
    """slots: element slot per lane of a warp (32); a wavefront serves `group` lanes; worst number of DIFFERENT addresses the lanes of
    one wavefront send to the same bank group (1 = conflict-free)"""
    worst = 1
    for g0 in range(0, 32, group):
        banks = {}
        for s in slots[g0:g0 + group]:
            banks.setdefault(s % group, set()).add(s)
        worst = max(worst, max(len(v) for v in banks.values()))
    return worst


@pytest.mark.parametrize("L", [9, 10, 11, 12])
@pytest.mark.parametrize("kernel", ["rows", "cols"])
@pytest.mark.parametrize("elem", [8, 16])
def test_bank_conflict_degree_of_the_super_passes(L, kernel, elem):
    # 8-byte elements (float2, the relaxed fp32 path): 16 lanes per 128-byte wavefront; 16-byte elements (double2): 8 lanes.
    # Conflict-free everywhere, except the second super-pass of the double-precision path for lines of 512 and 1024, where the
    # eight lanes of a wavefront spread over N/256 <= 4 consecutive k and the swizzle separates only half of the rest: 2-way.
    swz, group = (swz8, 16) if elem == 8 else (swz16, 8)
    expected = [1, 2 if (elem == 16 and L <= 10) else 1, 1]
    N, S1, P = passes(L)
    worst = [1, 1, 1]
    for w in range((2 * S1) // 32):               # two lines per CTA in both kernels
        if kernel == "rows":                      # l = tid / TPL, u = tid % TPL: a warp lies inside one line
            us = [(32 * w + lane) % S1 for lane in range(32)]
            k1s = [k1_rows(u, S1) for u in us]
        else:                                     # lanes 0-15: first line, lanes 16-31: second line, u = 16 * warp + (lane & 15)
            us = [16 * w + (lane & 15) for lane in range(32)]
            k1s = us
        for pi, (A, Bs) in enumerate(P):
            for b in Bs:
                slots = [int(swz(A(u, k1) + b)) for u, k1 in zip(us, k1s)]
                worst[pi] = max(worst[pi], _conflict_degree(slots, group))
    assert worst == expected
