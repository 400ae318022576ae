#include <src/BatchRegistration.h>

#include <cstdlib>
#include <cstring>

#include <src/Kernel.h>

namespace {
bool valid_params(const Regularisation reg, const unsigned int n) {   // arities of the three drivers' valid_regularisation_parameters
    switch (reg) {
        case Regularisation::Diffusion: return n == 1;
        case Regularisation::Curvature: return n >= 1 && n <= 2;
        case Regularisation::Elastic: return n >= 2 && n <= 3;
        case Regularisation::ThirionsDemons: return n == 6;
        case Regularisation::DiffeomorphicDemons: return n == 5;
        case Regularisation::Fluid: return n >= 2 && n <= 3;
    }
    return false;
}
}  // namespace

BatchRegistration::BatchRegistration(const dim dimin, const int batch_, const int niter_, const int nrefine_, const Regularisation reg, const of2d_real* p,
                                     const unsigned int nparams, const int wave_, const int frames_)
    : grid(dimin), batch(batch_), niter(niter_), nrefine(nrefine_), wave(wave_), nwaves(1), frames(frames_ < 1 ? 1 : frames_), npix((size_t)dimin.x * dimin.y),
      engine(nullptr), Iref(nullptr), Imov(nullptr), motion(nullptr), staging(nullptr), ctx_in(nullptr), ctx_out(nullptr) {
    for (int k = 0; k < 2; k++) s_in[k] = s_ref[k] = s_mov[k] = s_mot[k] = s_out[k] = nullptr;
    if (batch <= 0 || niter < 0 || nrefine < 1) throw std::invalid_argument("BatchRegistration: bad batch / niter / nrefine");
    if (!valid_params(reg, nparams)) throw std::invalid_argument("Invalid number of regularisation parameters for given regularisation method.\n");
    if (frames > 1) {   // cine chains: one wave per frame
        if (batch % frames != 0) throw std::invalid_argument("BatchRegistration: the batch must be a multiple of the number of frames");
        wave = batch / frames;
        nwaves = frames;
    } else {
        if (wave <= 0) {
            const char* e = std::getenv("OF2D_BATCH_WAVE");
            wave = e && std::atoi(e) > 0 ? std::atoi(e) : 128;
        }
        if (wave > batch) wave = batch;
        nwaves = (batch + wave - 1) / wave;
        wave = (batch + nwaves - 1) / nwaves;   // balanced: the last wave is short by less than nwaves pairs (padded with copies)
    }

    of2d_engine_desc d;
    std::memset(&d, 0, sizeof(d));
    d.method = (int)reg;
    d.dimx = (int)grid.x; d.dimy = (int)grid.y; d.batch = wave;
    d.real_is_double = sizeof(of2d_real) == 8;
    d.max_iter = niter;
    // the reference's defaults: Elastic `omega = 0.66f` (OpticalFlowElastic.h:9), Fluid `omega = 0.66` (OpticalFlowFluid.h:10):
    // they differ in the fp64 build
    const of2d_real omega_elastic = 0.66f, omega_fluid = 0.66;
    Kernel kf(1u), kd(1u);
    switch (reg) {
        case Regularisation::Diffusion: d.alpha = p[0]; break;
        case Regularisation::Curvature: d.alpha = p[0]; d.tau = nparams > 1 ? p[1] : (of2d_real)1; break;
        case Regularisation::Elastic: d.mu = p[0]; d.lambda = p[1]; d.omega = nparams > 2 ? p[2] : omega_elastic; break;
        case Regularisation::Fluid: d.mu = p[0]; d.lambda = p[1]; d.omega = nparams > 2 ? p[2] : omega_fluid; break;
        case Regularisation::ThirionsDemons:
        case Regularisation::DiffeomorphicDemons: {
            d.sigma_i = p[0]; d.sigma_x = p[1];
            const unsigned int width = static_cast<unsigned int>(p[4]);
            d.kernel_w = (int)width;
            kd = Kernel(width); kd.set_gaussian(p[2]);
            kf = Kernel(width); kf.set_gaussian(p[3]);
            d.kernel_diffusion = kd.get_kernel();
            d.kernel_fluid = kf.get_kernel();
            d.accumulation = reg == Regularisation::ThirionsDemons ? (int)p[5] : 0;
            break;
        }
    }
    of2d::check(of2d_engine_create(of2d::context(), &d, &engine));
    iters.assign((size_t)batch, 0);
    nregrid.assign((size_t)batch, 0);
}

BatchRegistration::~BatchRegistration() {
    if (engine) of2d_engine_destroy(engine);
    delete Iref;
    delete Imov;
    delete motion;
    delete staging;
    for (int k = 0; k < 2; k++) { delete s_in[k]; delete s_ref[k]; delete s_mov[k]; delete s_mot[k]; delete s_out[k]; }
    if (ctx_in) of2d_ctx_destroy(ctx_in);
    if (ctx_out) of2d_ctx_destroy(ctx_out);
}

// a partial last wave is filled up with copies of its first pair (valid images: no spurious divide-by-zero); their results are dropped
void BatchRegistration::pad_wave(of2d_ctx* ctx, of2d_real* img, int m) const {
    for (int k = m; k < wave; k++) of2d::check(of2d_d2d(ctx, img + npix * (size_t)k, img, sizeof(of2d_real) * npix));
}

// Image::set_image for every pair: doubles cross the bus, the cast to `real` happens on the device
void BatchRegistration::set_images(const double* ref, const double* mov) {
    of2d_ctx* ctx = of2d::context();
    const size_t rb = sizeof(of2d_real), cap = npix * (size_t)wave * (size_t)nwaves;
    if (!Iref) {
        Iref = new of2d::Buffer(rb * cap);
        Imov = new of2d::Buffer(rb * cap);
        motion = new of2d::Buffer(2 * rb * cap);
        staging = new of2d::Buffer(sizeof(double) * 2 * npix * (size_t)wave);
    }
    double* st = (double*)staging->device_discard();
    of2d_real* dr = (of2d_real*)Iref->device_discard();
    of2d_real* dm = (of2d_real*)Imov->device_discard();
    for (int w = 0; w < nwaves; w++) {
        const int p0 = w * wave, m = batch - p0 < wave ? batch - p0 : wave;
        const size_t cnt = npix * (size_t)m, off = npix * (size_t)p0;
        of2d::check(of2d_h2d(ctx, st, ref + off, sizeof(double) * cnt));
        of2d::check(of2d::image_from_double(cnt, st, dr + off));
        of2d::check(of2d_h2d(ctx, st + cnt, mov + off, sizeof(double) * cnt));
        of2d::check(of2d::image_from_double(cnt, st + cnt, dm + off));
        if (m < wave) { pad_wave(ctx, dr + off, m); pad_wave(ctx, dm + off, m); }
    }
    of2d::check(of2d_ctx_sync(ctx));   // the copies read the caller's buffers: complete before returning (as Image::set_image)
}

// one wave on the engine: cold start per pair (or the previous frame's state in a cine chain), nrefine passes
void BatchRegistration::solve_wave(int w, const of2d_real* dr, const of2d_real* dm, of2d_real* mo, int m) {
    if (frames == 1 || w == 0) of2d::check(of2d_engine_reset_state(engine));   // fresh solver state per pair; chains keep it
    const int p0 = w * wave;
    for (int refine = 0; refine < nrefine; refine++) {
        const int st = sizeof(of2d_real) == 8 ? of2d_engine_refine_f64(engine, (const double*)dr, (const double*)dm, (double*)mo, niter)
                                              : of2d_engine_refine_f32(engine, (const float*)dr, (const float*)dm, (float*)mo, niter);
        of2d::check(st);
        for (int k = 0; k < m; k++) {
            int it = 0, rg = 0;
            of2d::check(of2d_engine_pair_result(engine, k, &it, &rg, nullptr));
            if (refine == 0) { iters[(size_t)(p0 + k)] = 0; nregrid[(size_t)(p0 + k)] = 0; }
            iters[(size_t)(p0 + k)] += it;
            nregrid[(size_t)(p0 + k)] += rg;
        }
    }
}

void BatchRegistration::estimate_motion() {
    if (!Iref) throw std::invalid_argument("BatchRegistration: set_images() first");
    of2d_ctx* ctx = of2d::context();
    motion->zero();
    of2d_real* mo = (of2d_real*)motion->device_rw();
    const of2d_real* dr = (const of2d_real*)Iref->device_ro();
    const of2d_real* dm = (const of2d_real*)Imov->device_ro();
    for (int w = 0; w < nwaves; w++) {
        const int p0 = w * wave, m = batch - p0 < wave ? batch - p0 : wave;
        const size_t off = npix * (size_t)p0;
        if (frames > 1 && w > 0)   // cine chain: frame w starts from the motion frame w - 1 ended with
            of2d::check(of2d_d2d(ctx, mo + 2 * off, mo + 2 * (off - npix * (size_t)wave), 2 * sizeof(of2d_real) * npix * (size_t)wave));
        solve_wave(w, dr + off, dm + off, mo + 2 * off, m);
    }
}

void BatchRegistration::copy_estimated_motion(double* out) const {
    if (!motion) throw std::invalid_argument("BatchRegistration: nothing estimated yet");
    of2d_ctx* ctx = of2d::context();
    const of2d_real* mo = (const of2d_real*)motion->device_ro();
    double* st = (double*)staging->device_discard();
    for (int w = 0; w < nwaves; w++) {
        const int p0 = w * wave, m = batch - p0 < wave ? batch - p0 : wave;
        of2d::check(sizeof(of2d_real) == 8 ? of2d_motion_to_planar_double_batch_f64(ctx, npix, m, (const double*)(mo + 2 * npix * (size_t)p0), st)
                                           : of2d_motion_to_planar_double_batch_f32(ctx, npix, m, (const float*)(mo + 2 * npix * (size_t)p0), st));
        of2d::check(of2d_d2h(ctx, out + 2 * npix * (size_t)p0, st, sizeof(double) * 2 * npix * (size_t)m));
    }
}

// Streamed protocol.  Three streams: `in` (host -> device copies and the double -> real casts of the NEXT wave), the
// context's compute stream (the engine), `out` (real -> planar double and the device -> host copy of the PREVIOUS wave).
// of2d_engine_refine returns when its wave is complete, so the host enqueues the neighbours' copies just before it calls the
// solve of wave k; the streams are ordered by of2d_ctx_wait_for at the buffer hand-overs.
void BatchRegistration::register_pairs(const double* ref, const double* mov, double* out) {
    of2d_ctx* ctx = of2d::context();
    const size_t rb = sizeof(of2d_real), wn = npix * (size_t)wave;
    if (!ctx_in) {
        of2d::check(of2d_ctx_create(of2d_ctx_device(ctx), &ctx_in));
        of2d::check(of2d_ctx_create(of2d_ctx_device(ctx), &ctx_out));
        of2d::check(of2d_ctx_make_current(ctx));
        for (int k = 0; k < 2; k++) {
            s_in[k] = new of2d::Buffer(sizeof(double) * 2 * wn);
            s_ref[k] = new of2d::Buffer(rb * wn);
            s_mov[k] = new of2d::Buffer(rb * wn);
            s_mot[k] = new of2d::Buffer(2 * rb * wn);
            s_out[k] = new of2d::Buffer(sizeof(double) * 2 * wn);
        }
        of2d::check(of2d_ctx_sync(ctx));   // allocations and clears are ordered on the compute stream: visible to the copy streams from here on
    }
    auto count = [&](int w) { const int p0 = w * wave; return batch - p0 < wave ? batch - p0 : wave; };
    auto issue_in = [&](int w) {   // `in` stream: images of wave w -> buffer set w & 1
        const int b = w & 1, m = count(w);
        const size_t cnt = npix * (size_t)m, off = npix * (size_t)(w * wave);
        double* st = (double*)s_in[b]->device_discard();
        of2d_real* dr = (of2d_real*)s_ref[b]->device_discard();
        of2d_real* dm = (of2d_real*)s_mov[b]->device_discard();
        of2d::check(of2d_h2d(ctx_in, st, ref + off, sizeof(double) * cnt));
        of2d::check(of2d_h2d(ctx_in, st + wn, mov + off, sizeof(double) * cnt));
        of2d::check(sizeof(of2d_real) == 8 ? of2d_image_from_double_f64(ctx_in, cnt, st, (double*)dr) : of2d_image_from_double_f32(ctx_in, cnt, st, (float*)dr));
        of2d::check(sizeof(of2d_real) == 8 ? of2d_image_from_double_f64(ctx_in, cnt, st + wn, (double*)dm) : of2d_image_from_double_f32(ctx_in, cnt, st + wn, (float*)dm));
        if (m < wave) { pad_wave(ctx_in, dr, m); pad_wave(ctx_in, dm, m); }
    };
    auto issue_out = [&](int w) {   // `out` stream: motion of wave w (solved: the engine call has returned) -> caller
        const int b = w & 1, m = count(w);
        const of2d_real* mo = (const of2d_real*)s_mot[b]->device_ro();
        double* st = (double*)s_out[b]->device_discard();
        // one launch for the wave (a launch per pair waited for a gap between the solve's kernels each time: 11 % of the streamed rate)
        of2d::check(sizeof(of2d_real) == 8 ? of2d_motion_to_planar_double_batch_f64(ctx_out, npix, m, (const double*)mo, st)
                                           : of2d_motion_to_planar_double_batch_f32(ctx_out, npix, m, (const float*)mo, st));
        of2d::check(of2d_d2h_async(ctx_out, out + 2 * npix * (size_t)(w * wave), st, sizeof(double) * 2 * npix * (size_t)m));
    };
    issue_in(0);
    for (int w = 0; w < nwaves; w++) {
        const int b = w & 1;
        // the compute stream takes over buffer set b: after the copies of wave w (in) and after wave w - 2 has left it (out)
        of2d::check(of2d_ctx_wait_for(ctx, ctx_in));
        of2d::check(of2d_ctx_wait_for(ctx, ctx_out));
        of2d_real* mo = (of2d_real*)s_mot[b]->device_discard();
        if (frames > 1 && w > 0) of2d::check(of2d_d2d(ctx, mo, s_mot[b ^ 1]->device_ro(), 2 * rb * wn));   // cine chain
        else of2d::check(of2d_memset(ctx, mo, 0, 2 * rb * wn));
        // neighbours' copies go under this wave's solve: set b ^ 1 is free (wave w - 1 is solved, wave w + 1 not started)
        if (w + 1 < nwaves) issue_in(w + 1);
        if (w >= 1) issue_out(w - 1);
        solve_wave(w, (const of2d_real*)s_ref[b]->device_ro(), (const of2d_real*)s_mov[b]->device_ro(), mo, count(w));
    }
    issue_out(nwaves - 1);
    of2d::check(of2d_ctx_sync(ctx_out));
    of2d::check(of2d_ctx_sync(ctx_in));
}
