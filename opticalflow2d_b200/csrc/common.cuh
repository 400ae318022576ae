// common.cuh -- shared device/host helpers for libof2d_cuda (sm_100a only).
//
// Arithmetic policy: the library is compiled with -fmad=false, so every expression written with
// plain * and + rounds exactly like the reference's baseline-x86-64 build (no FMA contraction,
// SURVEY Q20).  Kernels that are flop-bound additionally take a FAST template flag and use explicit
// fma() where that is allowed to differ from the reference by rounding only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/of2d_cuda.h"

struct of2d_ctx {
    int device;
    int sm_count;
    cudaStream_t own_stream;
    cudaStream_t stream;
    int fast_math;           // arithmetic level: 0 strict (per-step kernels, the reference loop literally), 1 exact engine, 2 relaxed engine (default)
    uint64_t launches;
    // scratch for reductions: per-block partials + a pinned mailbox for scalar results
    double *d_partials;      // [kMaxPartialBlocks * 4]
    unsigned *d_status;      // per-pair status words [kMaxBatchStatus]
    unsigned *d_progress;    // wavefront progress counters
    size_t progress_cap;
    unsigned progress_epoch;
    void *h_mailbox;         // pinned, 4 KB
    void *d_mailbox;         // device, 4 KB
    void *d_kernel;          // device copy of convolution weights
    size_t kernel_cap;
    struct of2d_profiler *prof;   // per-kernel CUDA-event timing (NULL unless enabled)
};

constexpr int kMaxPartialBlocks = 4096;
constexpr int kMaxBatchStatus = 8192;

void of2d_set_error(const char *fmt, ...);

// per-kernel timing with CUDA events on the launching stream (ctx.cu); no-ops unless profiling is enabled
void of2d_prof_begin(of2d_ctx *ctx, const char *name);
void of2d_prof_end(of2d_ctx *ctx);
struct ProfScope {
    of2d_ctx *ctx;
    ProfScope(of2d_ctx *c, const char *name) : ctx(c) { if (c->prof) of2d_prof_begin(c, name); }
    ~ProfScope() { if (ctx->prof) of2d_prof_end(ctx); }
};

#define OF2D_CUDA_TRY(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            of2d_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return OF2D_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

#define OF2D_REQUIRE(cond, msg)                                    \
    do {                                                           \
        if (!(cond)) {                                             \
            of2d_set_error("%s: %s", __func__, msg);               \
            return OF2D_ERR_INVALID;                               \
        }                                                          \
    } while (0)

#define OF2D_LAUNCH_CHECK(ctx)                                     \
    do {                                                           \
        (ctx)->launches++;                                         \
        OF2D_CUDA_TRY(cudaGetLastError());                         \
    } while (0)

// ---- programmatic dependent launch (sm_90+) -------------------------------------------------------
// Every kernel of the iteration engine starts with pdl_enter(): wait until the previous kernel of the stream has
// completed and its writes are visible (griddepcontrol.wait), then let the NEXT kernel of the stream be launched
// (griddepcontrol.launch_dependents): its CTAs are scheduled while this kernel runs and sit in their own wait, so the
// launch latency between the ~4 000 dependent kernels of a step is hidden.  Without the launch attribute (pdl_launch below
// its level, or a plain <<<>>> launch) both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
int of2d_pdl_level();   // OF2D_PDL: 0 = off, 1 = on except for the SOR sweep (default), 2 = every engine kernel
template <int LEVEL = 1, class... KArgs, class... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = of2d_pdl_level() >= LEVEL ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- vector types -------------------------------------------------------------------------------
template <class R> struct Vec2T;
template <> struct Vec2T<float> { using type = float2; };
template <> struct Vec2T<double> { using type = double2; };
template <class R> using vec2_t = typename Vec2T<R>::type;

template <class R> __host__ __device__ __forceinline__ vec2_t<R> mk2(R x, R y) {
    vec2_t<R> v; v.x = x; v.y = y; return v;
}

// precision-selected libm (the reference calls the float overloads on float data)
__device__ __forceinline__ float r_floor(float x) { return floorf(x); }
__device__ __forceinline__ double r_floor(double x) { return floor(x); }
__device__ __forceinline__ float r_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double r_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float r_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double r_fma(double a, double b, double c) { return fma(a, b, c); }

// ---- reductions -----------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <class R> __device__ __forceinline__ R warp_max(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { R w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}
template <class R> __device__ __forceinline__ R warp_min(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { R w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    return v;
}

// block-wide sum of two doubles; result valid in thread 0. blockDim <= 1024.
__device__ __forceinline__ void block_sum2(double &a, double &b) {
    __shared__ double sa[32], sb[32];
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { sa[wid] = a; sb[wid] = b; }
    __syncthreads();
    if (wid == 0) {
        a = lane < nw ? sa[lane] : 0.0;
        b = lane < nw ? sb[lane] : 0.0;
        a = warp_sum(a); b = warp_sum(b);
    }
    __syncthreads();
}
template <class R, bool IS_MAX> __device__ __forceinline__ R block_extreme(R v) {
    __shared__ R sv[32];
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    v = IS_MAX ? warp_max(v) : warp_min(v);
    if (lane == 0) sv[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = sv[lane < nw ? lane : 0];
        v = IS_MAX ? warp_max(v) : warp_min(v);
    }
    __syncthreads();
    return v;
}

// raises (never lowers) a kernel's dynamic shared-memory limit; one record per kernel function per process
int of2d_ensure_dynamic_smem(const void *kernel, size_t bytes);

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
