// mex_harness.cpp -- in-process stand-in for the interpreter side of the MEX API plus the C entry
// points declared in include/of2d_host.h.  Lets tests and bench.py drive mexFunction() and the C++
// classes through ctypes exactly as Octave would, and read back what the reference only prints.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <exception>
#include <functional>
#include <stdexcept>
#include <thread>
#include <string>
#include <vector>

#include <mex.h>
#include <of2d_host.h>

#include <src/BatchRegistration.h>
#include <src/DeviceRuntime.h>
#include <src/Image.h>
#include <src/Kernel.h>
#include <src/ImageRegistrationDemons.h>
#include <src/ImageRegistrationFluid.h>
#include <src/ImageRegistrationOpticalFlow.h>
#include <src/Motion.h>

struct mxArray_tag {
    double* data;
    mwSize ndim;
    mwSize dims[4];
    mwSize numel;
};

ImageRegistration* of2d_wrapper_registration();

namespace {
std::string g_error;
std::string g_printed;
bool g_capture = false;

template <class F>
int guarded(F&& body) {
    try {
        body();
        return OF2D_HOST_OK;
    } catch (const std::invalid_argument& e) {
        g_error = e.what();
        return OF2D_HOST_EINVAL;
    } catch (const std::runtime_error& e) {
        g_error = e.what();
        return OF2D_HOST_ERUNTIME;
    } catch (const std::exception& e) {
        g_error = e.what();
        return OF2D_HOST_EOTHER;
    }
}

const RegistrationTrace* trace_of(const ImageRegistration* r) { return r ? &r->get_trace() : nullptr; }

int copy_out(const std::vector<double>& v, double* out, int cap) {
    const int n = (int)v.size() < cap ? (int)v.size() : cap;
    if (out && n > 0) memcpy(out, v.data(), sizeof(double) * (size_t)n);
    return (int)v.size();
}
}  // namespace

// ---- interpreter side of the MEX API -----------------------------------------------------------
extern "C" {

double* mxGetPr(const mxArray* a) { return a->data; }

mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID, mxComplexity) {
    mxArray* a = new mxArray_tag();
    a->ndim = ndim;
    a->numel = 1;
    for (int d = 0; d < 4; d++) a->dims[d] = 1;
    for (mwSize d = 0; d < ndim && d < 4; d++) { a->dims[d] = dims[d]; a->numel *= dims[d]; }
    a->data = static_cast<double*>(calloc(a->numel ? a->numel : 1, sizeof(double)));
    return a;
}

void mxDestroyArray(mxArray* a) {
    if (!a) return;
    free(a->data);
    delete a;
}

int mexPrintf(const char* fmt, ...) {
    if (!g_capture) return 0;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (g_printed.size() < (1u << 22)) g_printed += buf;
    return n;
}

void mexErrMsgTxt(const char* msg) { throw std::runtime_error(std::string("mexErrMsgTxt: ") + msg); }

// ---- include/of2d_host.h -------------------------------------------------------------------------
int of2d_host_real_bits(void) { return (int)sizeof(of2d_real) * 8; }
const char* of2d_host_last_error(void) { return g_error.c_str(); }
void of2d_host_capture_printf(int on) { g_capture = on != 0; g_printed.clear(); }
const char* of2d_host_printed(void) { return g_printed.c_str(); }
static int default_math_level() {
    const char* e = std::getenv("OF2D_MATH");
    if (!e || !*e) return 2;
    const std::string s(e);
    if (s == "strict") return 0;
    if (s == "exact") return 1;
    if (s == "relaxed") return 2;
    const int v = std::atoi(e);
    return v < 0 ? 0 : v > 2 ? 2 : v;
}
int of2d_host_set_strict(int strict) {
    return guarded([&] { of2d::check(of2d_ctx_set_fast_math(of2d::context(), strict ? 0 : default_math_level())); });
}
int of2d_host_set_math(int level) {
    return guarded([&] { of2d::check(of2d_ctx_set_fast_math(of2d::context(), level)); });
}
int of2d_host_get_math(void) {
    int level = -1;
    guarded([&] { level = of2d_ctx_get_fast_math(of2d::context()); });
    return level;
}
int of2d_host_set_stream(void* cuda_stream) {
    return guarded([&] { of2d::check(of2d_ctx_set_stream(of2d::context(), cuda_stream)); });
}
int of2d_host_use_own_stream(void) {
    return guarded([&] { of2d::check(of2d_ctx_use_own_stream(of2d::context())); });
}
int of2d_host_sync(void) {
    return guarded([&] { of2d::check(of2d_ctx_sync(of2d::context())); });
}
unsigned long long of2d_host_launch_count(void) {
    unsigned long long n = 0;
    guarded([&] { n = of2d_ctx_launch_count(of2d::context()); });
    return n;
}
void of2d_host_shutdown(void) { of2d::release_context(); }
int of2d_host_profile_enable(int on) {
    return guarded([&] { of2d::check(of2d_ctx_profile_enable(of2d::context(), on)); });
}
int of2d_host_profile_read(char* buf, size_t cap) {
    return guarded([&] { of2d::check(of2d_ctx_profile_read(of2d::context(), buf, cap)); });
}

void* of2d_mx_create(int ndim, const size_t* dims) { return mxCreateNumericArray((mwSize)ndim, dims, mxDOUBLE_CLASS, mxREAL); }
double* of2d_mx_data(void* mx) { return static_cast<mxArray*>(mx)->data; }
size_t of2d_mx_numel(void* mx) { return static_cast<mxArray*>(mx)->numel; }
int of2d_mx_ndim(void* mx) { return (int)static_cast<mxArray*>(mx)->ndim; }
size_t of2d_mx_dim(void* mx, int d) { return static_cast<mxArray*>(mx)->dims[d]; }
void of2d_mx_free(void* mx) { mxDestroyArray(static_cast<mxArray*>(mx)); }

int of2d_mex_call(int nlhs, void** plhs, int nrhs, void** prhs) {
    return guarded([&] { mexFunction(nlhs, reinterpret_cast<mxArray**>(plhs), nrhs, const_cast<const mxArray**>(reinterpret_cast<mxArray**>(prhs))); });
}

// ---- session API: the C++ classes without the singleton ---------------------------------------------
struct of2d_session {
    std::unique_ptr<ImageRegistration> reg;
    dim grid;
};

int of2d_session_create(int dimx, int dimy, int nscales, const int* niter, int nrefine, int reg, const double* regparams, int nparams, int verbose,
                        of2d_session** out) {
    *out = nullptr;
    return guarded([&] {
        std::vector<of2d_real> p((size_t)(nparams > 0 ? nparams : 1));
        for (int k = 0; k < nparams; k++) p[(size_t)k] = (of2d_real)regparams[k];
        const dim grid((unsigned int)dimx, (unsigned int)dimy);
        const Regularisation r = static_cast<Regularisation>(reg);
        const Verbose v = static_cast<Verbose>(verbose);
        std::unique_ptr<of2d_session> s(new of2d_session());
        s->grid = grid;
        if (reg >= 0 && reg <= 2) s->reg.reset(new ImageRegistrationOpticalFlow(grid, nscales, niter, nrefine, r, p.data(), (unsigned)nparams, v));
        else if (reg == 3 || reg == 4) s->reg.reset(new ImageRegistrationDemons(grid, nscales, niter, nrefine, r, p.data(), (unsigned)nparams, v));
        else if (reg == 5) s->reg.reset(new ImageRegistrationFluid(grid, nscales, niter, nrefine, r, p.data(), (unsigned)nparams, v));
        else mexErrMsgTxt("Error: invalid regularisation given\n");
        *out = s.release();
    });
}
void of2d_session_destroy(of2d_session* s) { delete s; }

int of2d_session_set_images(of2d_session* s, const double* Iref, const double* Imov) {
    return guarded([&] {
        Image r(s->grid), m(s->grid);
        r.set_image(Iref);
        s->reg->set_reference_image(r);
        m.set_image(Imov);
        s->reg->set_moving_image(m);
    });
}
int of2d_session_reset(of2d_session* s) {
    return guarded([&] { s->reg->reset_state(); });
}
int of2d_session_estimate(of2d_session* s) {
    return guarded([&] { s->reg->estimate_motion(); });
}
int of2d_session_get_motion(of2d_session* s, double* planar_out) {
    return guarded([&] {
        Motion m(s->grid);
        s->reg->copy_estimated_motion(m);
        m.copy_motion_to_input(planar_out);
    });
}
// ---- several registrations in one call, copies under the solves -------------------------------------------------------
// set_images / estimate_motion / copy_estimated_motion of sessions[0 .. n-1] with three streams: `in` (host -> device copy and
// double -> real cast of job k+1, pyramids included), the context's compute stream (job k), `out` (real -> planar double and
// device -> host copy of job k-1).  Results are the ones the three separate calls give; the sessions must be distinct.
namespace {
struct SidePipes {
    of2d_ctx* main = nullptr;
    of2d_ctx* in = nullptr;
    of2d_ctx* out = nullptr;
    of2d::Buffer* s_in[2] = {nullptr, nullptr};
    of2d::Buffer* s_out[2] = {nullptr, nullptr};
    size_t cap = 0;   // pixels the staging buffers hold
};
thread_local SidePipes tl_pipes;

void ensure_pipes(size_t npix) {
    SidePipes& P = tl_pipes;
    of2d_ctx* ctx = of2d::context();
    if (P.main != ctx) {   // (contexts live as long as the process: a changed main context just gets new side streams)
        P = SidePipes();
        P.main = ctx;
        of2d::check(of2d_ctx_create(of2d_ctx_device(ctx), &P.in));
        of2d::check(of2d_ctx_create(of2d_ctx_device(ctx), &P.out));
        of2d::check(of2d_ctx_make_current(ctx));
    }
    if (P.cap < npix) {
        of2d::check(of2d_ctx_sync(P.in));
        of2d::check(of2d_ctx_sync(P.out));
        for (int k = 0; k < 2; k++) {
            delete P.s_in[k]; delete P.s_out[k];
            P.s_in[k] = new of2d::Buffer(sizeof(double) * 2 * npix);
            P.s_out[k] = new of2d::Buffer(sizeof(double) * 2 * npix);
        }
        P.cap = npix;
        of2d::check(of2d_ctx_sync(ctx));   // allocations and clears are ordered on the compute stream: visible to the copy streams from here on
    }
}
}  // namespace

int of2d_sessions_register(of2d_session* const* sessions, int n, const double* const* Iref, const double* const* Imov, double* const* planar_out) {
    return guarded([&] {
        if (n <= 0) return;
        size_t cap = 0;
        for (int k = 0; k < n; k++) {
            if (!sessions[k] || !Iref[k] || !Imov[k] || !planar_out[k]) throw std::invalid_argument("of2d_sessions_register: null argument");
            for (int q = 0; q < k; q++) if (sessions[q] == sessions[k]) throw std::invalid_argument("of2d_sessions_register: the sessions must be distinct");
            const size_t np = (size_t)sessions[k]->grid.x * sessions[k]->grid.y;
            cap = np > cap ? np : cap;
        }
        ensure_pipes(cap);
        SidePipes& P = tl_pipes;
        of2d_ctx* ctx = P.main;
        auto npix = [&](int k) { return (size_t)sessions[k]->grid.x * sessions[k]->grid.y; };
        auto upload = [&](int k) {   // `in` stream
            const size_t np = npix(k);
            double* st = (double*)P.s_in[k & 1]->device_discard();
            ImageRegistration* reg = sessions[k]->reg.get();
            of2d::check(of2d_h2d(P.in, st, Iref[k], sizeof(double) * np));
            of2d::check(of2d_h2d(P.in, st + np, Imov[k], sizeof(double) * np));
            of2d_real* dr = reg->reference_level0()->device_overwrite();
            of2d_real* dm = reg->moving_level0()->device_overwrite();
            of2d::check(sizeof(of2d_real) == 8 ? of2d_image_from_double_f64(P.in, np, st, (double*)dr) : of2d_image_from_double_f32(P.in, np, st, (float*)dr));
            of2d::check(sizeof(of2d_real) == 8 ? of2d_image_from_double_f64(P.in, np, st + np, (double*)dm) : of2d_image_from_double_f32(P.in, np, st + np, (float*)dm));
            of2d::set_thread_context(P.in);      // the coarser levels on the same stream
            try { reg->rebuild_image_pyramids(); } catch (...) { of2d::set_thread_context(nullptr); of2d_ctx_make_current(ctx); throw; }
            of2d::set_thread_context(nullptr);
            of2d::check(of2d_ctx_make_current(ctx));
        };
        auto download = [&](int k) {   // `out` stream, after the compute stream has finished job k
            const size_t np = npix(k);
            double* st = (double*)P.s_out[k & 1]->device_discard();
            const of2d_real* mo = reinterpret_cast<const of2d_real*>(sessions[k]->reg->get_estimated_motion()->device());
            of2d::check(of2d_ctx_wait_for(P.out, ctx));
            of2d::check(sizeof(of2d_real) == 8 ? of2d_motion_to_planar_double_f64(P.out, np, (const double*)mo, st) : of2d_motion_to_planar_double_f32(P.out, np, (const float*)mo, st));
            of2d::check(of2d_d2h_async(P.out, planar_out[k], st, sizeof(double) * 2 * np));
        };
        struct Drain {   // the caller's buffers must not be in flight when the call returns, error or not
            SidePipes& P;
            ~Drain() { of2d_ctx_sync(P.in); of2d_ctx_sync(P.out); }
        } drain{P};
        upload(0);
        for (int k = 0; k < n; k++) {
            of2d::check(of2d_ctx_wait_for(ctx, P.in));   // job k's images are in place (job k+1's copies are enqueued after this point)
            if (k + 1 < n) upload(k + 1);
            if (k >= 1) download(k - 1);
            sessions[k]->reg->estimate_motion();
        }
        download(n - 1);
    });
}

int of2d_session_get_motion_aos(of2d_session* s, void* out_real) {
    return guarded([&] {
        const Motion* m = s->reg->get_estimated_motion();
        memcpy(out_real, m->get_motion(), sizeof(vector2d) * m->get_size());
    });
}
int of2d_session_warp(of2d_session* s, const double* img, double* out) {
    return guarded([&] {
        Image m(s->grid);
        m.set_image(img);
        m.warp2d(*s->reg->get_estimated_motion());
        m.copy_image_to_input(out);
    });
}

// ---- the public Image / Motion / Kernel methods that no driver calls, for callers without C++ ------------
// Images cross as column-major doubles (Image::set_image / copy_image_to_input), motions as AoS doubles {x, y} per pixel.
int of2d_host_image_op(int op, int dimx, int dimy, const double* in, double* out, double* scalars, int kernel_w, double sigma) {
    return guarded([&] {
        Image im(dim((unsigned int)dimx, (unsigned int)dimy));
        im.set_image(in);
        switch (op) {
            case 0:   // Image::sum / max / min, src/Image.cpp:78-104
                scalars[0] = (double)im.sum(); scalars[1] = (double)im.max(); scalars[2] = (double)im.min();
                break;
            case 1:   // Image::normalize, src/Image.cpp:107-116
                im.normalize();
                im.copy_image_to_input(out);
                break;
            case 2: {   // Image::convolute, src/Image.cpp:184-187, with Kernel::set_gaussian (sigma > 0) or Kernel::set_average
                Kernel k((unsigned int)kernel_w);
                if (sigma > 0) k.set_gaussian((of2d_real)sigma); else k.set_average();
                im.convolute(k);
                im.copy_image_to_input(out);
                break;
            }
            default: throw std::invalid_argument("of2d_host_image_op: unknown operation");
        }
    });
}
int of2d_host_motion_boundary(int kind, int dimx, int dimy, const double* aos_in, double* aos_out) {
    return guarded([&] {
        Motion mo(dim((unsigned int)dimx, (unsigned int)dimy));
        vector2d* u = mo.get_motion();
        const size_t n = (size_t)dimx * (size_t)dimy;
        for (size_t k = 0; k < n; k++) { u[k].x = (of2d_real)aos_in[2 * k]; u[k].y = (of2d_real)aos_in[2 * k + 1]; }
        if (kind) mo.Dirichlet_boundaryconditions(); else mo.Neumann_boundaryconditions();   // src/Motion.cpp:181-251
        u = mo.get_motion();
        for (size_t k = 0; k < n; k++) { aos_out[2 * k] = (double)u[k].x; aos_out[2 * k + 1] = (double)u[k].y; }
    });
}
int of2d_host_kernel(int kind, int w, double sigma, double* out) {
    return guarded([&] {
        Kernel k((unsigned int)w);
        if (kind == 0) k.set_gaussian((of2d_real)sigma); else k.set_average();   // src/Kernel.cpp:45-82
        memcpy(out, k.get_kernel(), sizeof(double) * k.get_size());
    });
}

// ---- batch API (extension): independent pairs registered together ----------------------------------
// A batch is one or more SHARDS: contiguous pair ranges, each on its own device behind its own context and driven by its
// own host thread (one host thread, one CUDA context, compute + copy streams per GPU; no collective in the
// solve).  The default is a single shard on the process context, run inline in the calling thread.
struct of2d_batch {
    struct Shard {
        of2d_ctx* ctx = nullptr;                 // nullptr: the process context (of2d::context())
        std::unique_ptr<BatchRegistration> reg;
        int lo = 0, hi = 0;                      // pair range [lo, hi)
    };
    std::vector<Shard> shards;
    size_t npix = 0;
    int batch = 0;

    // fn(shard) on every shard, each in its own thread with its context current; the first exception is rethrown here
    void each(const std::function<void(Shard&)>& fn) {
        if (shards.size() == 1 && !shards[0].ctx) { fn(shards[0]); return; }
        std::vector<std::thread> th;
        std::vector<std::exception_ptr> err(shards.size());
        for (size_t k = 0; k < shards.size(); k++)
            th.emplace_back([&, k] {
                try {
                    of2d::set_thread_context(shards[k].ctx);
                    fn(shards[k]);
                } catch (...) { err[k] = std::current_exception(); }
                try { of2d::set_thread_context(nullptr); } catch (...) {}
            });
        for (auto& t : th) t.join();
        for (auto& e : err) if (e) std::rethrow_exception(e);
    }
    ~of2d_batch() {
        try { each([](Shard& s) { s.reg.reset(); }); } catch (...) {}
        for (auto& s : shards) if (s.ctx) of2d_ctx_destroy(s.ctx);
    }
};

// contiguous shard [lo, hi) of `total` pairs for `rank` of `world` devices (needs no GPU): the whole multi-GPU protocol of the path
int of2d_shard_range(int total, int world, int rank, int* lo, int* hi) {
    if (total < 0 || world < 1 || rank < 0 || rank >= world) return 2;
    const int base = total / world, rem = total % world;
    const int l = rank * base + (rank < rem ? rank : rem);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (rank < rem ? 1 : 0);
    return 0;
}

static int batch_create(int dimx, int dimy, int batch, int frames, int niter, int nrefine, int reg, const double* regparams, int nparams, int wave,
                        const int* devices, int ndevices, of2d_batch** out) {
    *out = nullptr;
    return guarded([&] {
        std::vector<of2d_real> p((size_t)(nparams > 0 ? nparams : 1));
        for (int k = 0; k < nparams; k++) p[(size_t)k] = (of2d_real)regparams[k];
        if (reg < 0 || reg > 5) mexErrMsgTxt("Error: invalid regularisation given\n");
        if (batch <= 0 || dimx <= 0 || dimy <= 0) throw std::invalid_argument("of2d_batch_create: bad batch / dimensions");
        if (ndevices < 0 || (ndevices > 0 && !devices)) throw std::invalid_argument("of2d_batch_create_multi: bad device list");
        if (frames > 1 && batch % frames != 0) throw std::invalid_argument("of2d_batch_create_chain: the batch must be a multiple of the number of frames");
        std::unique_ptr<of2d_batch> b(new of2d_batch());
        b->npix = (size_t)dimx * (size_t)dimy;
        b->batch = batch;
        const int nsh = ndevices > 0 ? (ndevices < batch ? ndevices : batch) : 1;
        b->shards.resize((size_t)nsh);
        const int level = ndevices > 0 ? of2d_ctx_get_fast_math(of2d::context()) : 0;
        for (int k = 0; k < nsh; k++) {   // contiguous ranges, the same partition as one process per GPU uses
            of2d_shard_range(batch, nsh, k, &b->shards[(size_t)k].lo, &b->shards[(size_t)k].hi);
            if (ndevices > 0) {
                of2d::check(of2d_ctx_create(devices[k], &b->shards[(size_t)k].ctx));
                of2d::check(of2d_ctx_set_fast_math(b->shards[(size_t)k].ctx, level));
            }
        }
        if (ndevices > 0) of2d::check(of2d_ctx_make_current(of2d::context()));
        b->each([&](of2d_batch::Shard& s) {
            s.reg.reset(new BatchRegistration(dim((unsigned int)dimx, (unsigned int)dimy), s.hi - s.lo, niter, nrefine, static_cast<Regularisation>(reg), p.data(),
                                              (unsigned)nparams, wave, frames));
        });
        *out = b.release();
    });
}

int of2d_batch_create(int dimx, int dimy, int batch, int niter, int nrefine, int reg, const double* regparams, int nparams, int wave, of2d_batch** out) {
    return batch_create(dimx, dimy, batch, 1, niter, nrefine, reg, regparams, nparams, wave, nullptr, 0, out);
}
int of2d_batch_create_chain(int dimx, int dimy, int batch, int frames, int niter, int nrefine, int reg, const double* regparams, int nparams, of2d_batch** out) {
    return batch_create(dimx, dimy, batch, frames, niter, nrefine, reg, regparams, nparams, 0, nullptr, 0, out);
}
int of2d_batch_create_multi(int dimx, int dimy, int batch, int niter, int nrefine, int reg, const double* regparams, int nparams, int wave, const int* devices,
                            int ndevices, of2d_batch** out) {
    if (ndevices <= 0) { *out = nullptr; return guarded([] { throw std::invalid_argument("of2d_batch_create_multi: no devices given"); }); }
    return batch_create(dimx, dimy, batch, 1, niter, nrefine, reg, regparams, nparams, wave, devices, ndevices, out);
}
void of2d_batch_destroy(of2d_batch* b) { delete b; }
int of2d_batch_set_images(of2d_batch* b, const double* Iref, const double* Imov) {
    return guarded([&] { b->each([&](of2d_batch::Shard& s) { s.reg->set_images(Iref + b->npix * (size_t)s.lo, Imov + b->npix * (size_t)s.lo); }); });
}
int of2d_batch_estimate(of2d_batch* b) {
    return guarded([&] { b->each([&](of2d_batch::Shard& s) { s.reg->estimate_motion(); }); });
}
int of2d_batch_get_motion(of2d_batch* b, double* planar_out) {
    return guarded([&] { b->each([&](of2d_batch::Shard& s) { s.reg->copy_estimated_motion(planar_out + 2 * b->npix * (size_t)s.lo); }); });
}
int of2d_batch_register(of2d_batch* b, const double* Iref, const double* Imov, double* planar_out) {
    return guarded([&] {
        b->each([&](of2d_batch::Shard& s) { s.reg->register_pairs(Iref + b->npix * (size_t)s.lo, Imov + b->npix * (size_t)s.lo, planar_out + 2 * b->npix * (size_t)s.lo); });
    });
}
int of2d_batch_iterations(of2d_batch* b, int* iterations, int* regrids) {
    return guarded([&] {
        for (const auto& s : b->shards)
            for (int k = 0; k < s.reg->size(); k++) {
                if (iterations) iterations[s.lo + k] = s.reg->iterations()[(size_t)k];
                if (regrids) regrids[s.lo + k] = s.reg->regrids()[(size_t)k];
            }
    });
}
int of2d_batch_wave(of2d_batch* b) { return b->shards[0].reg->wave_size(); }
int of2d_batch_num_shards(of2d_batch* b) { return (int)b->shards.size(); }
int of2d_batch_shard_info(of2d_batch* b, int shard, int* device, int* lo, int* hi) {
    return guarded([&] {
        if (shard < 0 || shard >= (int)b->shards.size()) throw std::invalid_argument("of2d_batch_shard_info: shard out of range");
        const auto& s = b->shards[(size_t)shard];
        if (device) *device = of2d_ctx_device(s.ctx ? s.ctx : of2d::context());
        if (lo) *lo = s.lo;
        if (hi) *hi = s.hi;
    });
}

// ---- trace access: `s` may be NULL to address the MEX singleton -------------------------------------
static const RegistrationTrace* pick_trace(of2d_session* s) { return trace_of(s ? s->reg.get() : of2d_wrapper_registration()); }

int of2d_trace_num_levels(of2d_session* s) {
    const RegistrationTrace* t = pick_trace(s);
    return t ? (int)t->levels.size() : 0;
}
long of2d_trace_total_iterations(of2d_session* s) {
    const RegistrationTrace* t = pick_trace(s);
    return t ? t->total_iterations() : 0;
}
int of2d_trace_level_info(of2d_session* s, int level, int* scale, int* refine, int* iterations, int* nregrid) {
    const RegistrationTrace* t = pick_trace(s);
    if (!t || level < 0 || level >= (int)t->levels.size()) return OF2D_HOST_EINVAL;
    const RegistrationTrace::Level& l = t->levels[(size_t)level];
    if (scale) *scale = l.scale;
    if (refine) *refine = l.refine;
    if (iterations) *iterations = l.iterations;
    if (nregrid) *nregrid = (int)l.regrid_iteration.size();
    return OF2D_HOST_OK;
}
/* which: 0 error, 1 regrid iteration, 2 regrid min-Jacobian, 3 fluid maxabs, 4 fluid dt; returns the full length */
int of2d_trace_level_series(of2d_session* s, int level, int which, double* out, int cap) {
    const RegistrationTrace* t = pick_trace(s);
    if (!t || level < 0 || level >= (int)t->levels.size()) return 0;
    const RegistrationTrace::Level& l = t->levels[(size_t)level];
    switch (which) {
        case 0: return copy_out(l.error, out, cap);
        case 1: {
            std::vector<double> v(l.regrid_iteration.begin(), l.regrid_iteration.end());
            return copy_out(v, out, cap);
        }
        case 2: return copy_out(l.regrid_minjac, out, cap);
        case 3: return copy_out(l.fluid_maxabs, out, cap);
        case 4: return copy_out(l.fluid_dt, out, cap);
    }
    return 0;
}

}  // extern "C"
