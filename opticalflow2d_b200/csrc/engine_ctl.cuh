// engine_ctl.cuh -- device-resident control block of the iteration engine.
//
// The reference's driver loops take three data-dependent decisions per iteration on the host:
// the convergence break (err < 0.001 && iter > 1), the fluid time step (skip when dt >= 65) and the
// fluid regrid (min Jacobian < 0.5).  Here they are taken ON THE DEVICE by the last CTA of the kernel
// that produces the statistic ("last-block" reduction: every CTA publishes its partial, takes a
// ticket, and the CTA drawing the last ticket reduces all partials in a fixed order, so results are
// deterministic).  Every kernel of an iteration starts with `if (!ctl.active) return;`, so the host
// can enqueue iterations ahead without synchronising: iterations past the break are empty launches
// and the state stays exactly where the reference's loop would have left it.
#pragma once

#include "common.cuh"

struct alignas(16) PairCtl {
    int active;          // 1 while the inner loop of this pair is running
    int iter;            // iterations completed so far in this refine (Logger::iter)
    int niter;           // cap
    unsigned flags;      // OF2D_FLAG_* (sticky until the host reads them)
    int sel;             // which of the two estimate buffers holds the current estimate
    int regrid;          // fluid: regrid requested by the iteration that just finished
    int skip;            // fluid: integration skipped (dt >= 65)
    int nsquares;        // diffeomorphic: squarings of the current iteration
    int nregrid;         // fluid: regrid events so far in this refine
    int msel;            // fluid: which of the two level-motion buffers is current
    int overflow;        // nsquares above the enqueued squarings: host must not trust the result
    int vsel;            // fluid: which velocity buffer is current (persists across refines and calls, like the velocity)
    int prev_other;      // fluid: Logger's prev is the other estimate buffer (set by a regrid, which zeroes the estimate)
    int redo;            // diffusion: the two-step kernel met the break test after its FIRST step; the single-step kernel that follows redoes it
    int pad0[2];
    unsigned ticket[4];  // last-block tickets (one per reduction kind)
    double err;          // last Logger error
    double maxabs;       // fluid / diffeo: sqrt(max(2 y^2))
    double dt;           // fluid time step
    double scale;        // diffeo: 2^-nsquares
    double minjac;
};

// per-iteration traces, [batch][cap] each
struct TraceDev {
    double *err;      // Logger error (rounded to `real` first)
    double *maxabs;   // fluid increment maxabs / diffeo correspondence maxabs
    double *dt;       // fluid
    double *minjac;   // fluid (meaningful where regrid == 1)
    int *regrid;      // fluid: 1 if a regrid followed this iteration
    int *nsq;         // diffeo
    int cap;
};

// Publishes NV per-CTA partial values (valid in thread 0) and returns true, for every thread of the
// CTA, in the CTA that arrived last; that CTA may then read all `nblocks` partials of this pair.
template <int NV>
__device__ __forceinline__ bool publish_partials(const double (&vals)[NV], double *__restrict__ partials, unsigned *ticket, int nblocks, int bid) {
    __shared__ int s_last;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    if (tid == 0) {
#pragma unroll
        for (int v = 0; v < NV; v++) __stcg(&partials[(size_t)bid * NV + v], vals[v]);
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == (unsigned)nblocks - 1u);
        if (s_last) { *ticket = 0u; __threadfence(); }
    }
    __syncthreads();
    return s_last != 0;
}

// fixed-order reduction of the published partials by the last CTA: max for value indices in max_mask,
// min for min_mask, sum otherwise.  Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void reduce_partials(const double *__restrict__ partials, int nblocks, double (&out)[NV], unsigned max_mask, unsigned min_mask) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nt = blockDim.x * blockDim.y;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = (max_mask >> v & 1u) ? -INFINITY : ((min_mask >> v & 1u) ? INFINITY : 0.0);
    for (int k = tid; k < nblocks; k += nt) {
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const double p = __ldcg(&partials[(size_t)k * NV + v]);
            if (max_mask >> v & 1u) acc[v] = p > acc[v] ? p : acc[v];
            else if (min_mask >> v & 1u) acc[v] = p < acc[v] ? p : acc[v];
            else acc[v] += p;
        }
    }
#pragma unroll
    for (int v = 0; v < NV; v++) {
        if (max_mask >> v & 1u) out[v] = block_extreme<double, true>(acc[v]);
        else if (min_mask >> v & 1u) out[v] = block_extreme<double, false>(acc[v]);
        else { double dummy = 0.0; double a = acc[v]; block_sum2(a, dummy); out[v] = a; }
    }
}

// Logger::update_error + the drivers' break test (Logger.cpp:32-51, ImageRegistrationOpticalFlow.cpp:131-134).
// `sum_diff`, `sum_prev` are the double sums of |u - prev| and |prev| over the n pixels of the pair.
// Called by ONE thread.  Returns true if the loop of this pair goes on.
template <class R>
__device__ __forceinline__ bool finalize_logger(PairCtl *c, const TraceDev &tr, int pair, double sum_diff, double sum_prev, unsigned n, int *n_active) {
    const R diffnorm = (R)sum_diff / (R)n, prevnorm = (R)sum_prev / (R)n;     // Motion.cpp:47
    const R err = prevnorm == 0 ? (R)0.0f : diffnorm / prevnorm;               // Logger.cpp:39
    const int it = c->iter;
    c->err = (double)err;
    if (it < tr.cap) tr.err[(size_t)pair * tr.cap + it] = (double)err;
    c->iter = it + 1;
    if ((err < (R)0.001f && it > 1) || it + 1 >= c->niter) {
        c->active = 0;
        atomicSub(n_active, 1);
        return false;
    }
    return true;
}
