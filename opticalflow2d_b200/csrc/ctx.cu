#include <stdlib.h>
// ctx.cu -- context, memory and stream plumbing of libof2d_cuda.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

struct of2d_profiler {
    struct Rec { std::string name; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t take() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
};

int of2d_ensure_dynamic_smem(const void *kernel, size_t bytes) {
    // cudaFuncSetAttribute is per device: one record per (device, kernel), guarded for concurrent host threads
    static std::map<std::pair<int, const void *>, size_t> configured;
    static std::mutex mu;
    if (bytes <= 40 * 1024) return OF2D_SUCCESS;   // the 48 KiB default covers static + dynamic: opt in with some room for the static part
    int dev = 0;
    OF2D_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = configured[std::make_pair(dev, kernel)];
    if (bytes > cur) {
        OF2D_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
    return OF2D_SUCCESS;
}

void of2d_prof_begin(of2d_ctx *c, const char *name) {
    of2d_profiler::Rec r; r.name = name; r.a = c->prof->take(); r.b = c->prof->take();
    cudaEventRecord(r.a, c->stream);
    c->prof->recs.push_back(r);
}
void of2d_prof_end(of2d_ctx *c) { cudaEventRecord(c->prof->recs.back().b, c->stream); }

static thread_local char g_error[512] = "";

void of2d_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" {

const char *of2d_last_error(void) { return g_error; }

int of2d_device_count(int *count) {
    *count = 0;
    OF2D_CUDA_TRY(cudaGetDeviceCount(count));
    return OF2D_SUCCESS;
}

int of2d_ctx_create(int device, of2d_ctx **out) {
    *out = nullptr;
    int n = 0;
    OF2D_CUDA_TRY(cudaGetDeviceCount(&n));
    if (n == 0 || device < 0 || device >= n) {
        of2d_set_error("of2d_ctx_create: no usable CUDA device %d (count %d); this library has no CPU fallback", device, n);
        return OF2D_ERR_CUDA;
    }
    OF2D_CUDA_TRY(cudaSetDevice(device));
    of2d_ctx *c = new of2d_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->fast_math = 2;
    {   // OF2D_MATH = strict | exact | relaxed (or 0 | 1 | 2): the arithmetic level new contexts start in
        const char *e = getenv("OF2D_MATH");
        if (e && *e) c->fast_math = !strcmp(e, "strict") ? 0 : !strcmp(e, "exact") ? 1 : !strcmp(e, "relaxed") ? 2 : (atoi(e) < 0 ? 0 : atoi(e) > 2 ? 2 : atoi(e));
    }
    OF2D_CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    OF2D_CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    {   // keep freed blocks in the pool instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        OF2D_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        OF2D_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    c->stream = c->own_stream;
    OF2D_CUDA_TRY(cudaMalloc(&c->d_partials, sizeof(double) * kMaxPartialBlocks * 4));
    OF2D_CUDA_TRY(cudaMalloc(&c->d_status, sizeof(unsigned) * kMaxBatchStatus));
    OF2D_CUDA_TRY(cudaMemset(c->d_status, 0, sizeof(unsigned) * kMaxBatchStatus));
    OF2D_CUDA_TRY(cudaHostAlloc(&c->h_mailbox, 4096, cudaHostAllocDefault));
    OF2D_CUDA_TRY(cudaMalloc(&c->d_mailbox, 4096));
    OF2D_CUDA_TRY(cudaMemset(c->d_mailbox, 0, 4096));
    *out = c;
    return OF2D_SUCCESS;
}

void of2d_ctx_destroy(of2d_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_partials);
    cudaFree(c->d_status);
    cudaFree(c->d_progress);
    cudaFree(c->d_mailbox);
    cudaFree(c->d_kernel);
    cudaFreeHost(c->h_mailbox);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

int of2d_ctx_set_stream(of2d_ctx *c, void *s) {
    c->stream = (cudaStream_t)s;
    return OF2D_SUCCESS;
}
int of2d_ctx_use_own_stream(of2d_ctx *c) {
    c->stream = c->own_stream;
    return OF2D_SUCCESS;
}
void *of2d_ctx_get_stream(of2d_ctx *c) { return (void *)c->stream; }
int of2d_ctx_sync(of2d_ctx *c) {
    OF2D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OF2D_SUCCESS;
}
int of2d_ctx_make_current(of2d_ctx *c) {
    OF2D_CUDA_TRY(cudaSetDevice(c->device));
    return OF2D_SUCCESS;
}
int of2d_ctx_device(of2d_ctx *c) { return c->device; }
int of2d_ctx_wait_for(of2d_ctx *waiter, of2d_ctx *signaller) {
    // everything enqueued on the signaller's stream so far happens before whatever the waiter's stream is given next
    cudaEvent_t ev;
    OF2D_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    OF2D_CUDA_TRY(cudaEventRecord(ev, signaller->stream));
    OF2D_CUDA_TRY(cudaStreamWaitEvent(waiter->stream, ev, 0));
    OF2D_CUDA_TRY(cudaEventDestroy(ev));   // released once the recorded work completes
    return OF2D_SUCCESS;
}
int of2d_ctx_set_fast_math(of2d_ctx *c, int on) {
    c->fast_math = on < 0 ? 0 : on > 2 ? 2 : on;
    return OF2D_SUCCESS;
}
int of2d_ctx_get_fast_math(of2d_ctx *c) { return c->fast_math; }
uint64_t of2d_ctx_launch_count(of2d_ctx *c) { return c->launches; }

int of2d_ctx_profile_enable(of2d_ctx *c, int on) {
    if (on && !c->prof) c->prof = new of2d_profiler();
    if (c->prof) {
        cudaStreamSynchronize(c->stream);
        for (auto &r : c->prof->recs) { c->prof->pool.push_back(r.a); c->prof->pool.push_back(r.b); }
        c->prof->recs.clear();
    }
    if (!on && c->prof) {
        for (cudaEvent_t e : c->prof->pool) cudaEventDestroy(e);
        delete c->prof;
        c->prof = nullptr;
    }
    return OF2D_SUCCESS;
}

// "name count total_ms\n" per kernel since profiling was enabled (or last read); synchronises the stream
int of2d_ctx_profile_read(of2d_ctx *c, char *buf, size_t cap) {
    if (!buf || cap == 0) return OF2D_ERR_INVALID;
    buf[0] = 0;
    if (!c->prof) return OF2D_SUCCESS;
    OF2D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::map<std::string, std::pair<long, double>> agg;
    for (auto &r : c->prof->recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { auto &e = agg[r.name]; e.first++; e.second += ms; }
        c->prof->pool.push_back(r.a); c->prof->pool.push_back(r.b);
    }
    c->prof->recs.clear();
    std::string out;
    for (auto &kv : agg) { char line[256]; snprintf(line, sizeof(line), "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second); out += line; }
    strncpy(buf, out.c_str(), cap - 1);
    buf[cap - 1] = 0;
    return OF2D_SUCCESS;
}

int of2d_malloc(of2d_ctx *c, size_t bytes, void **p) {
    *p = nullptr;
    // stream-ordered allocation from the device's default pool (kept cached: see of2d_ctx_create), so the
    // temporaries the reference API creates per call (Image / Motion copies) cost no cudaMalloc / cudaFree
    OF2D_CUDA_TRY(cudaMallocAsync(p, bytes ? bytes : 1, c->stream));
    return OF2D_SUCCESS;
}
int of2d_free(of2d_ctx *c, void *p) {
    if (!p) return OF2D_SUCCESS;
    OF2D_CUDA_TRY(cudaFreeAsync(p, c->stream));
    return OF2D_SUCCESS;
}
int of2d_host_alloc(size_t bytes, void **p) {
    *p = nullptr;
    OF2D_CUDA_TRY(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
    return OF2D_SUCCESS;
}
int of2d_host_free(void *p) {
    if (p) OF2D_CUDA_TRY(cudaFreeHost(p));
    return OF2D_SUCCESS;
}
int of2d_memset(of2d_ctx *c, void *p, int byte, size_t bytes) {
    OF2D_CUDA_TRY(cudaMemsetAsync(p, byte, bytes, c->stream));
    return OF2D_SUCCESS;
}
int of2d_h2d(of2d_ctx *c, void *d, const void *h, size_t bytes) {
    OF2D_CUDA_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
    return OF2D_SUCCESS;
}
int of2d_d2h(of2d_ctx *c, void *h, const void *d, size_t bytes) {
    OF2D_CUDA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    OF2D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OF2D_SUCCESS;
}
int of2d_d2h_async(of2d_ctx *c, void *h, const void *d, size_t bytes) {
    OF2D_CUDA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    return OF2D_SUCCESS;
}
int of2d_d2d(of2d_ctx *c, void *dst, const void *src, size_t bytes) {
    OF2D_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return OF2D_SUCCESS;
}

}  // extern "C"

int of2d_pdl_level() {
    static int lvl = -1;
    if (lvl < 0) { const char *e = getenv("OF2D_PDL"); lvl = e ? atoi(e) : 1; if (lvl < 0) lvl = 0; }
    return lvl;
}
