"""Batched registration (-m gpu): every pair of a batch must get exactly what a fresh single-pair
registration object computes for it (same engine kernels, per-pair control blocks), and the batch
must agree with the CPU oracle within the north-star tolerance."""
import numpy as np
import pytest

import opticalflow2d_b200 as of
from gpu_common import maxdiff, oracle
from opticalflow2d_b200 import synthetic as S

# these tests pin the EXACT engine (arithmetic level 1) bit for bit; the default relaxed engine: tests/test_relaxed_gpu.py
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_engine")]

CASES = [
    (of.THIRION, [1.0, 0.25, 1.5, 1.5, 5, 0], 12),
    (of.DIFFEOMORPHIC, [1.0, 2.0, 1.5, 1.5, 5], 8),
    (of.FLUID, [0.1, 0.0], 14),
    (of.ELASTIC, [1.0, 0.25], 10),
    (of.DIFFUSION, [0.5], 10),
    (of.CURVATURE, [0.25, 1.0], 6),
]


def _pairs(n, dimx, dimy):
    R = np.zeros((n, dimy, dimx)); T = np.zeros((n, dimy, dimx))
    for k in range(n):
        R[k], T[k] = S.batch_pair(k, dimx, dimy)
    return R, T


@pytest.mark.parametrize("reg,params,niter", CASES, ids=[of.METHOD_NAMES[c[0]] for c in CASES])
def test_batch_equals_single_pair_sessions(reg, params, niter):
    dimx, dimy, n = 96, 64, 6
    of.set_strict(False, 32)
    R, T = _pairs(n, dimx, dimy)
    with of.Batch((dimx, dimy), n, niter, reg, params, wave=3) as b:
        b.set_images(R, T)
        b.estimate()
        got = b.motion()
        its, _ = b.iterations()
    for k in range(n):
        with of.Session((dimx, dimy), [niter], 0, reg, params, nrefine=1, verbose=0, bits=32) as s:
            s.set_images(R[k], T[k])
            s.estimate()
            want = s.motion()
            assert s.trace()["total_iterations"] == its[k]
        assert maxdiff(got[k], want) == 0.0, (k, maxdiff(got[k], want))


@pytest.mark.parametrize("reg,params,niter", CASES[:3], ids=[of.METHOD_NAMES[c[0]] for c in CASES[:3]])
def test_batch_matches_oracle(reg, params, niter):
    dimx, dimy, n = 64, 64, 4
    of.set_strict(False, 32)
    R, T = _pairs(n, dimx, dimy)
    with of.Batch((dimx, dimy), n, niter, reg, params, wave=4) as b:
        b.set_images(R, T)
        b.estimate()
        got = b.motion()
        its, _ = b.iterations()
    orc = oracle(32)
    for k in range(n):
        want = orc.register(R[k], T[k], reg, params, [niter], nscales=0, nrefine=1, verbose=1)
        assert len(want["err"]) == its[k]
        assert maxdiff(got[k], want["motion"]) <= 1e-3   # north-star fp32 tolerance, px


def test_batch_early_convergence_is_per_pair():
    """A pair of identical images stops after the reference's minimum of three iterations (zero motion, Logger error 0,
    SURVEY Q10); its neighbours in the batch keep iterating."""
    dimx, dimy = 64, 64
    of.set_strict(False, 32)
    R, T = _pairs(3, dimx, dimy)
    T[1] = R[1]
    with of.Batch((dimx, dimy), 3, 30, of.THIRION, [1.0, 0.25, 1.5, 1.5, 5, 0], wave=3) as b:
        b.set_images(R, T)
        b.estimate()
        its, _ = b.iterations()
    assert its[1] == 3 and its[0] > 3 and its[2] > 3, its


def test_batch_two_step_diffusion_breaks_per_pair():
    """Diffusion runs two Jacobi steps per launch (k_hs_pair).  The pair of identical images meets the break test on the
    FIRST step of the second launch (iteration index 2): that launch must leave its state alone and the next one redo the
    single step for this pair only, while its neighbours in the batch go on two steps at a time."""
    dimx, dimy, niter = 64, 64, 31
    of.set_strict(False, 32)
    R, T = _pairs(3, dimx, dimy)
    T[1] = R[1]
    with of.Batch((dimx, dimy), 3, niter, of.DIFFUSION, [0.5], wave=3) as b:
        b.set_images(R, T)
        b.estimate()
        got = b.motion()
        its, _ = b.iterations()
    assert its[1] == 3 and its[0] > 3 and its[2] > 3, its
    for k in range(3):
        with of.Session((dimx, dimy), [niter], 0, of.DIFFUSION, [0.5], nrefine=1, verbose=0, bits=32) as s:
            s.set_images(R[k], T[k])
            s.estimate()
            assert s.trace()["total_iterations"] == its[k]
            assert maxdiff(got[k], s.motion()) == 0.0
    want = oracle(32).register(R[0], T[0], of.DIFFUSION, [0.5], [niter], nscales=0, nrefine=1, verbose=1)
    assert len(want["err"]) == its[0] and maxdiff(got[0], want["motion"]) == 0.0


THIRION = (of.THIRION, [1.0, 0.25, 1.5, 1.5, 5, 0], 12)
FLUID = (of.FLUID, [0.1, 0.0], 14)


def _single(dimx, dimy, R, T, reg, params, niter):
    with of.Session((dimx, dimy), [niter], 0, reg, params, nrefine=1, verbose=0, bits=32) as s:
        s.set_images(R, T)
        s.estimate()
        return s.motion(), s.trace()["total_iterations"]


@pytest.mark.parametrize("reg,params,niter", [THIRION, FLUID], ids=["thirion", "fluid"])
def test_batch_of_any_size_pads_the_last_wave(reg, params, niter):
    """7 pairs in waves of 3: the waves are balanced (3 + 3 + 1) and the last one is padded with copies whose results are dropped."""
    dimx, dimy, n = 80, 64, 7
    of.set_strict(False, 32)
    R, T = _pairs(n, dimx, dimy)
    with of.Batch((dimx, dimy), n, niter, reg, params, wave=3) as b:
        assert b.wave() == 3
        b.set_images(R, T)
        b.estimate()
        got = b.motion()
        its, _ = b.iterations()
    for k in range(n):
        want, it = _single(dimx, dimy, R[k], T[k], reg, params, niter)
        assert it == its[k]
        assert maxdiff(got[k], want) == 0.0, k


@pytest.mark.parametrize("reg,params,niter", [THIRION, FLUID], ids=["thirion", "fluid"])
def test_streamed_protocol_equals_three_call_protocol(reg, params, niter):
    """of2d_batch_register: H2D of wave k+1 / D2H of wave k-1 under the solve of wave k, double-buffered: same bits as
    set_images + estimate + get_motion, also with a padded tail wave and when called twice on one object."""
    dimx, dimy, n = 80, 64, 8
    of.set_strict(False, 32)
    R, T = _pairs(n, dimx, dimy)
    with of.Batch((dimx, dimy), n, niter, reg, params, wave=3) as b:
        b.set_images(R, T)
        b.estimate()
        want = b.motion()
        its, rgs = b.iterations()
        got = b.register(R, T)
        its2, rgs2 = b.iterations()
        assert np.array_equal(its, its2) and np.array_equal(rgs, rgs2)
        assert maxdiff(got, want) == 0.0
        got2 = b.register(R[::-1].copy(), T[::-1].copy())     # buffers are reused: a second, different batch
        assert maxdiff(got2, want[::-1]) == 0.0


@pytest.mark.parametrize("reg,params,niter", [THIRION, FLUID, (of.DIFFUSION, [0.5], 9)], ids=["thirion", "fluid", "diffusion"])
def test_cine_chain_equals_repeated_estimates_on_one_object(reg, params, niter):
    """Frame f of a sequence warm-starts from frame f-1 (motion, and Fluid's velocity) exactly like a second estimate_motion()
    on one reference object (src/ImageRegistration.cpp:135-139)."""
    dimx, dimy, S_, F = 72, 64, 2, 3
    of.set_strict(False, 32)
    n = S_ * F
    R = np.empty((n, dimy, dimx)); T = np.empty((n, dimy, dimx))
    for f in range(F):
        for s in range(S_):
            R[f * S_ + s], T[f * S_ + s] = S.make_pair(dimx, dimy, "lattice", shift=(0.8 + 0.5 * f + 0.3 * s, -0.4 - 0.2 * f), sigma_b=6.0 + s)
    with of.Batch((dimx, dimy), n, niter, reg, params, frames=F) as b:
        assert b.wave() == S_
        b.set_images(R, T)
        b.estimate()
        got = b.motion()
        its, _ = b.iterations()
        streamed = b.register(R, T)
        assert maxdiff(streamed, got) == 0.0
    for s in range(S_):
        with of.Session((dimx, dimy), [niter], 0, reg, params, nrefine=1, verbose=0, bits=32) as ses:
            for f in range(F):
                ses.set_images(R[f * S_ + s], T[f * S_ + s])
                ses.estimate()
                assert ses.trace()["total_iterations"] == its[f * S_ + s]
                assert maxdiff(got[f * S_ + s], ses.motion()) == 0.0, (s, f)
    # and it is a warm start: frame 1 differs from a cold registration of the same pair
    cold, _ = _single(dimx, dimy, R[S_], T[S_], reg, params, niter)
    assert maxdiff(got[S_], cold) > 0.0


@pytest.mark.parametrize("reg,params,niter", [THIRION, FLUID], ids=["thirion", "fluid"])
def test_multi_shard_batch_equals_single_shard(reg, params, niter):
    """of2d_batch_create_multi: contiguous shards, each with its own context, streams and host thread.  With one GPU in
    the box the three shards share device 0 (the sharding, threading and per-shard contexts are what is under test)."""
    import torch
    dimx, dimy, n = 80, 64, 7
    of.set_strict(False, 32)
    R, T = _pairs(n, dimx, dimy)
    ndev = torch.cuda.device_count()
    devices = [k % ndev for k in range(3)]
    with of.Batch((dimx, dimy), n, niter, reg, params, wave=2) as b:
        want = b.register(R, T)
        wits, _ = b.iterations()
    with of.Batch((dimx, dimy), n, niter, reg, params, wave=2, devices=devices) as b:
        sh = b.shards()
        assert [(lo, hi) for _, lo, hi in sh] == [of.shard_pairs(n, 3, r) for r in range(3)]
        assert [d for d, _, _ in sh] == devices
        got = b.register(R, T)
        its, _ = b.iterations()
        assert np.array_equal(its, wits)
        assert maxdiff(got, want) == 0.0
        b.set_images(R, T)
        b.estimate()
        assert maxdiff(b.motion(), want) == 0.0
