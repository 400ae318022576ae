#include <src/DeviceRuntime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace of2d {

namespace {
of2d_ctx* g_ctx = nullptr;
std::mutex g_mutex;
thread_local of2d_ctx* tl_ctx = nullptr;   // per-thread override (multi-GPU batches: one host thread and one context per device)

int pick_device() {
    const char* names[] = {"OF2D_DEVICE", "LOCAL_RANK"};
    for (const char* n : names) {
        const char* v = std::getenv(n);
        if (v && *v) return std::atoi(v);
    }
    return 0;
}
}  // namespace

void set_thread_context(of2d_ctx* ctx) {
    tl_ctx = ctx;
    if (ctx) check(of2d_ctx_make_current(ctx));
}

of2d_ctx* context() {
    if (tl_ctx) return tl_ctx;
    if (g_ctx) return g_ctx;
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_ctx) {
        of2d_ctx* c = nullptr;
        const int st = of2d_ctx_create(pick_device(), &c);
        if (st != OF2D_SUCCESS)
            throw std::runtime_error(std::string("OpticalFlow2d needs a CUDA device and has no CPU path: ") + of2d_last_error());
        const char* strict = std::getenv("OF2D_STRICT");
        if (strict && *strict && std::atoi(strict) != 0) of2d_ctx_set_fast_math(c, 0);
        g_ctx = c;
    }
    return g_ctx;
}

void release_context() {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_ctx) of2d_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

void check(int status) {
    switch (status) {
        case OF2D_SUCCESS: return;
        case OF2D_ERR_INVALID: throw std::invalid_argument(of2d_last_error());
        case OF2D_ERR_DIVZERO: throw std::runtime_error("Divide by zero exception");
        default: throw std::runtime_error(std::string("of2d: ") + of2d_last_error());
    }
}

void poll_divzero() {
    unsigned flags = 0;
    check(of2d_poll_status(context(), 1, &flags));
    if (flags & OF2D_FLAG_DIVZERO) throw std::runtime_error("Divide by zero exception");
}

// ---------------------------------------------------------------------------------------------
Buffer::Buffer(size_t bytes) : bytes_(bytes), dptr_(nullptr), hptr_(nullptr), device_valid_(true), host_valid_(false) {
    check(of2d_malloc(context(), bytes_, &dptr_));
    check(of2d_memset(context(), dptr_, 0, bytes_));
}

Buffer::Buffer(const Buffer& other) : bytes_(other.bytes_), dptr_(nullptr), hptr_(nullptr), device_valid_(true), host_valid_(false) {
    check(of2d_malloc(context(), bytes_, &dptr_));
    check(of2d_d2d(context(), dptr_, other.device_ro(), bytes_));
}

Buffer::~Buffer() {
    if (dptr_) of2d_free(context(), dptr_);
    if (hptr_) of2d_host_free(hptr_);
}

const void* Buffer::device_ro() const {
    if (!device_valid_) {
        check(of2d_h2d(context(), dptr_, hptr_, bytes_));
        device_valid_ = true;
    }
    return dptr_;
}

void* Buffer::device_rw() {
    device_ro();
    host_valid_ = false;
    return dptr_;
}

void* Buffer::device_discard() {
    device_valid_ = true;
    host_valid_ = false;
    return dptr_;
}

void* Buffer::host() const {
    if (!hptr_) check(of2d_host_alloc(bytes_, &hptr_));
    else if (host_valid_ && device_valid_) check(of2d_ctx_sync(context()));   // an upload from this mirror may still be in flight: the caller may write
    if (!host_valid_) {
        check(of2d_d2h(context(), hptr_, dptr_, bytes_));
        host_valid_ = true;
    }
    // the caller gets a mutable pointer: assume it writes
    device_valid_ = false;
    return hptr_;
}

void Buffer::zero() {
    check(of2d_memset(context(), dptr_, 0, bytes_));
    device_valid_ = true;
    host_valid_ = false;
}

void Buffer::copy_from(const Buffer& other) {
    check(of2d_d2d(context(), dptr_, other.device_ro(), bytes_));
    device_valid_ = true;
    host_valid_ = false;
}

void Buffer::swap(Buffer& other) {
    std::swap(bytes_, other.bytes_);
    std::swap(dptr_, other.dptr_);
    std::swap(hptr_, other.hptr_);
    std::swap(device_valid_, other.device_valid_);
    std::swap(host_valid_, other.host_valid_);
}

}  // namespace of2d
