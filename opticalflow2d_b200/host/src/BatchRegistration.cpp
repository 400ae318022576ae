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
                                     const unsigned int nparams, const int wave_)
    : grid(dimin), batch(batch_), niter(niter_), nrefine(nrefine_), wave(wave_), npix((size_t)dimin.x * dimin.y), engine(nullptr), Iref(nullptr),
      Imov(nullptr), motion(nullptr), staging(nullptr) {
    if (batch <= 0 || niter < 0 || nrefine < 1) throw std::invalid_argument("BatchRegistration: bad batch / niter / nrefine");
    if (!valid_params(reg, nparams)) throw std::invalid_argument("Invalid number of regularisation parameters for given regularisation method.\n");
    if (wave <= 0) {
        const char* e = std::getenv("OF2D_BATCH_WAVE");
        wave = e && std::atoi(e) > 0 ? std::atoi(e) : 256;
    }
    if (wave > batch) wave = batch;

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
    const size_t rb = sizeof(of2d_real);
    Iref = new of2d::Buffer(rb * npix * (size_t)batch);
    Imov = new of2d::Buffer(rb * npix * (size_t)batch);
    motion = new of2d::Buffer(2 * rb * npix * (size_t)batch);
    staging = new of2d::Buffer(sizeof(double) * 2 * npix * (size_t)wave);
    iters.assign((size_t)batch, 0);
    nregrid.assign((size_t)batch, 0);
}

BatchRegistration::~BatchRegistration() {
    if (engine) of2d_engine_destroy(engine);
    delete Iref;
    delete Imov;
    delete motion;
    delete staging;
}

// Image::set_image for every pair: doubles cross the bus, the cast to `real` happens on the device
void BatchRegistration::set_images(const double* ref, const double* mov) {
    of2d_ctx* ctx = of2d::context();
    double* st = (double*)staging->device_discard();
    of2d_real* dr = (of2d_real*)Iref->device_discard();
    of2d_real* dm = (of2d_real*)Imov->device_discard();
    for (int p0 = 0; p0 < batch; p0 += wave) {
        const int m = batch - p0 < wave ? batch - p0 : wave;
        const size_t cnt = npix * (size_t)m, off = npix * (size_t)p0;
        of2d::check(of2d_h2d(ctx, st, ref + off, sizeof(double) * cnt));
        of2d::check(of2d::image_from_double(cnt, st, dr + off));
        of2d::check(of2d_h2d(ctx, st + cnt, mov + off, sizeof(double) * cnt));
        of2d::check(of2d::image_from_double(cnt, st + cnt, dm + off));
    }
}

void BatchRegistration::estimate_motion() {
    motion->zero();
    of2d_real* mo = (of2d_real*)motion->device_rw();
    const of2d_real* dr = (const of2d_real*)Iref->device_ro();
    const of2d_real* dm = (const of2d_real*)Imov->device_ro();
    for (int p0 = 0; p0 < batch; p0 += wave) {
        const int m = batch - p0 < wave ? batch - p0 : wave;
        if (m != wave) throw std::invalid_argument("BatchRegistration: the batch must be a multiple of the wave size");
        of2d::check(of2d_engine_reset_state(engine));   // fresh solver state per pair
        const size_t off = npix * (size_t)p0;
        for (int refine = 0; refine < nrefine; refine++) {
            const int st = sizeof(of2d_real) == 8 ? of2d_engine_refine_f64(engine, (const double*)(dr + off), (const double*)(dm + off), (double*)(mo + 2 * off), niter)
                                                  : of2d_engine_refine_f32(engine, (const float*)(dr + off), (const float*)(dm + off), (float*)(mo + 2 * off), niter);
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
}

void BatchRegistration::copy_estimated_motion(double* out) const {
    of2d_ctx* ctx = of2d::context();
    const of2d_real* mo = (const of2d_real*)motion->device_ro();
    double* st = (double*)staging->device_discard();
    for (int p0 = 0; p0 < batch; p0 += wave) {
        const int m = batch - p0 < wave ? batch - p0 : wave;
        for (int k = 0; k < m; k++)
            of2d::check(of2d::motion_to_planar(npix, mo + 2 * npix * (size_t)(p0 + k), st + 2 * npix * (size_t)k));
        of2d::check(of2d_d2h(ctx, out + 2 * npix * (size_t)p0, st, sizeof(double) * 2 * npix * (size_t)m));
    }
}
