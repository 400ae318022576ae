#include <src/ImageRegistration.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include <mex.h>

#include <src/Logger.h>
#include <src/regularization/OpticalFlow/OpticalFlowFluid.h>

namespace {
const char* kRule =
    "%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%\n";

const char* method_name(Regularisation reg) {
    switch (reg) {
        case Regularisation::Diffusion: return "Diffusion";
        case Regularisation::Curvature: return "Curvature";
        case Regularisation::Elastic: return "Elastic";
        case Regularisation::ThirionsDemons: return "Thirions Demons";
        case Regularisation::DiffeomorphicDemons: return "Diffeomorphic Demons";
        case Regularisation::Fluid: return "Fluid";
    }
    return "?";
}
}  // namespace

// same text as the reference banner (src/ImageRegistration.cpp:6-47)
void ImageRegistration::display_registration_parameters(const Regularisation reg, const of2d_real* regparams, const unsigned int nparams) const {
    mexPrintf(kRule);
    mexPrintf("Optical flow image registration started... (2D C++ implementation)...\n");
    mexPrintf("Registration parameters:\n");
    mexPrintf("dimensions:\t\t\t\t(%d %d)\n", dimin[0].x, dimin[0].y);
    mexPrintf("niter:\t\t\t\t\t(%d", niter[0]);
    for (int s = 1; s < nscales + 1; s++) mexPrintf(" %d", niter[s]);
    mexPrintf(")\n");
    mexPrintf("nscales:\t\t\t\t%d\n", nscales);
    mexPrintf("nrefine:\t\t\t\t%d\n", nrefine);
    mexPrintf("regularisation:\t\t\t\t%s\n", method_name(reg));
    if (nparams == 1) {
        mexPrintf("reg. param:\t\t\t\t%.2f\n", regparams[0]);
    } else {
        mexPrintf("reg. params:\t\t\t\t(%.2f", regparams[0]);
        for (unsigned int p = 1; p < nparams; p++) mexPrintf(" %.2f", regparams[p]);
        mexPrintf(")\n");
    }
    mexPrintf(kRule);
    mexPrintf("\n");
}

// reference src/ImageRegistration.cpp:49-84
ImageRegistration::ImageRegistration(const dim dimin_, const int nscales_, const int* niter_, const int nrefine_, const Regularisation reg,
                                     const of2d_real* regparams, const unsigned int nparams, const Verbose verbose_)
    : nscales(nscales_), nrefine(nrefine_), solver(nullptr), verbose(verbose_), method(reg), method_params(regparams, regparams + nparams) {
    const int levels = nscales + 1;
    engines.assign((size_t)levels, nullptr);
    engine_refused.assign((size_t)levels, 0);
    dimin = new dim[levels];
    sizein = new int[levels];
    for (int s = nscales; s >= 0; s--) {
        const of2d_real scale = std::pow(2, s);   // truncating division as in the reference (:57-59)
        dimin[s] = dim((unsigned int)(dimin_.x / scale), (unsigned int)(dimin_.y / scale));
        sizein[s] = (int)(dimin[s].x * dimin[s].y);
    }
    niter = new int[levels];
    std::memcpy(niter, niter_, sizeof(int) * levels);

    Iref = new Image*[levels];
    Imov = new Image*[levels];
    motion = new Motion*[levels];
    for (int s = nscales; s >= 0; s--) {
        Iref[s] = new Image(dimin[s]);
        Imov[s] = new Image(dimin[s]);
        motion[s] = new Motion(dimin[s]);
    }
    display_registration_parameters(reg, regparams, nparams);
}

void ImageRegistration::release_solvers() {
    if (!solver) return;
    for (int s = nscales; s >= 0; s--) delete solver[s];
    delete[] solver;
    solver = nullptr;
}

ImageRegistration::~ImageRegistration() {
    release_solvers();
    release_engines();
    for (int s = nscales; s >= 0; s--) {
        delete Iref[s];
        delete Imov[s];
        delete motion[s];
    }
    delete[] Iref;
    delete[] Imov;
    delete[] motion;
    delete[] dimin;
    delete[] sizein;
    delete[] niter;
}

// reference :103-121
void ImageRegistration::set_reference_image(const Image& im) {
    *Iref[0] = im;
    for (int s = nscales; s >= 1; s--) Iref[s]->downSample(*Iref[0]);
}
void ImageRegistration::set_moving_image(const Image& im) {
    *Imov[0] = im;
    for (int s = nscales; s >= 1; s--) Imov[s]->downSample(*Imov[0]);
}

void ImageRegistration::rebuild_image_pyramids() {
    for (int s = nscales; s >= 1; s--) { Iref[s]->downSample(*Iref[0]); Imov[s]->downSample(*Imov[0]); }
}

Motion* ImageRegistration::get_estimated_motion() const { return motion[0]; }
void ImageRegistration::copy_estimated_motion(Motion& mo) const { mo = *motion[0]; }

void ImageRegistration::reset_state() {
    for (int s = nscales; s >= 0; s--) {
        motion[s]->reset();
        if (solver && solver[s]) solver[s]->reset_state();
        if (engines[(size_t)s]) of2d::check(of2d_engine_reset_state(engines[(size_t)s]));
    }
}

// reference :133-156.  motion[nscales] is neither reset nor re-downsampled, so a second call on the
// same object warm-starts from the previous result, exactly as the reference does.
void ImageRegistration::estimate_motion() {
    trace = RegistrationTrace();
    for (int s = nscales; s >= 0; s--) {
        if (s > 0 && s < nscales) motion[s]->downSample(*motion[0]);
        current_scale = s;
        estimate_motion_at_current_resolution(motion[s], Iref[s], Imov[s], solver[s], niter[s], dimin[s], sizein[s]);
        if (s > 0) motion[0]->upSample(*motion[s]);
    }
}

// The loop of ImageRegistrationOpticalFlow.cpp:97-151, ImageRegistrationDemons.cpp:86-137 and
// ImageRegistrationFluid.cpp:67-142.  Differences between the families: OpticalFlow and Fluid fix the
// derivatives once per refine, Demons hands (Iref, Iaux) to every get_update; Fluid regrids when the
// Jacobian of the running estimate drops below 0.5.
void ImageRegistration::release_engines() {
    for (of2d_engine*& e : engines) {
        if (e) of2d_engine_destroy(e);
        e = nullptr;
    }
}

// Engine description from the registration parameters, with the defaults of the solver constructors
// (OpticalFlowCurvature.h:10 tau = 1; OpticalFlowElastic.h:9 / OpticalFlowFluid.h:10 omega = 0.66).
of2d_engine* ImageRegistration::engine_for_level(int level) {
    const size_t L = (size_t)level;
    if (engines[L] || engine_refused[L]) return engines[L];
    const std::vector<of2d_real>& p = method_params;
    of2d_engine_desc d;
    std::memset(&d, 0, sizeof(d));
    d.method = (int)method;
    d.dimx = (int)dimin[level].x; d.dimy = (int)dimin[level].y; d.batch = 1;
    d.real_is_double = sizeof(of2d_real) == 8;
    d.max_iter = niter[level];
    // the reference's defaults: Elastic `omega = 0.66f` (OpticalFlowElastic.h:9), Fluid `omega = 0.66` (OpticalFlowFluid.h:10):
    // they differ in the fp64 build
    const of2d_real omega_elastic = 0.66f, omega_fluid = 0.66;
    Kernel kf(1u), kd(1u);
    switch (method) {
        case Regularisation::Diffusion: d.alpha = p[0]; break;
        case Regularisation::Curvature: d.alpha = p[0]; d.tau = p.size() > 1 ? p[1] : (of2d_real)1; break;
        case Regularisation::Elastic: d.mu = p[0]; d.lambda = p[1]; d.omega = p.size() > 2 ? p[2] : omega_elastic; break;
        case Regularisation::Fluid: d.mu = p[0]; d.lambda = p[1]; d.omega = p.size() > 2 ? p[2] : omega_fluid; break;
        case Regularisation::ThirionsDemons:
        case Regularisation::DiffeomorphicDemons: {
            d.sigma_i = p[0]; d.sigma_x = p[1];
            const unsigned int width = static_cast<unsigned int>(p[4]);
            d.kernel_w = (int)width;
            kd = Kernel(width); kd.set_gaussian(p[2]);    // sigma_diffusion smooths the motion
            kf = Kernel(width); kf.set_gaussian(p[3]);    // sigma_fluid smooths the correspondence
            d.kernel_diffusion = kd.get_kernel();
            d.kernel_fluid = kf.get_kernel();
            d.accumulation = method == Regularisation::ThirionsDemons ? (int)p[5] : 0;
            break;
        }
    }
    of2d_engine* e = nullptr;
    const int st = of2d_engine_create(of2d::context(), &d, &e);
    if (st == OF2D_ERR_UNSUPPORTED) {
        // the engine declined this configuration (e.g. SOR parameters whose sweep does not contract fast enough for overlapped
        // tiles, kernels wider than 15): the per-iteration path runs instead -- same results, but one launch group and one host
        // round trip per iteration (Elastic / Fluid: the exact wavefront sweep, ~150x slower per sweep at 2048^2).  Say so once.
        engine_refused[L] = 1;
        std::fprintf(stderr, "OpticalFlow2d: level %d (%u x %u) runs on the per-iteration path, not on the device-resident engine: %s\n", level, dimin[level].x, dimin[level].y,
                     of2d_last_error());
        return nullptr;
    }
    of2d::check(st);
    engines[L] = e;
    return e;
}

// One level on the device-resident engine; returns false when the per-iteration path has to run instead.
// mexPrintf output is reconstructed from the device traces in the reference's order: the Fluid time-step
// line (OpticalFlowFluid.cpp:94), the Logger line when verbose (Logger.cpp:62-69), the regrid line
// (ImageRegistrationFluid.cpp:110).
bool ImageRegistration::run_level_on_engine(LoopKind kind, Motion* level_motion, const Image* ref, Image* mov, const int iterations, const dim d) {
    if (!of2d_ctx_get_fast_math(of2d::context())) return false;
    if (method == Regularisation::ThirionsDemons && method_params.size() > 5 && (int)method_params[5] != 0 && (int)method_params[5] != 1) return false;
    of2d_engine* eng = engine_for_level(current_scale);
    if (!eng) return false;
    for (int refine = 0; refine < nrefine; refine++) {
        RegistrationTrace::Level rec;
        rec.scale = current_scale;
        rec.refine = refine;
        const int st = sizeof(of2d_real) == 8
                           ? of2d_engine_refine_f64(eng, (const double*)ref->device(), (const double*)mov->device(), (double*)level_motion->device_mut(), iterations)
                           : of2d_engine_refine_f32(eng, (const float*)ref->device(), (const float*)mov->device(), (float*)level_motion->device_mut(), iterations);
        of2d::check(st);
        int its = 0, nreg = 0;
        of2d::check(of2d_engine_pair_result(eng, 0, &its, &nreg, nullptr));
        rec.iterations = its;
        std::vector<double> err((size_t)its), ma, dt, mj, rg;
        of2d::check(of2d_engine_trace(eng, 0, 0, err.data(), its));
        if (kind == LoopKind::Fluid) {
            ma.resize((size_t)its); dt.resize((size_t)its); mj.resize((size_t)its); rg.resize((size_t)its);
            of2d::check(of2d_engine_trace(eng, 0, 1, ma.data(), its));
            of2d::check(of2d_engine_trace(eng, 0, 2, dt.data(), its));
            of2d::check(of2d_engine_trace(eng, 0, 3, mj.data(), its));
            of2d::check(of2d_engine_trace(eng, 0, 4, rg.data(), its));
        }
        for (int it = 0; it < its; it++) {
            if (kind == LoopKind::Fluid) {
                mexPrintf("Dumax: %.3f\tMaxabs increment: %.3f\t Timestep: %.3f\n", (of2d_real)0.65f, (of2d_real)ma[(size_t)it], (of2d_real)dt[(size_t)it]);
                rec.fluid_maxabs.push_back(ma[(size_t)it]);
                rec.fluid_dt.push_back(dt[(size_t)it]);
            }
            if (verbose == Verbose::On) mexPrintf("Iteration: %d\tError:%.4f\n", it, (of2d_real)err[(size_t)it]);
            rec.error.push_back(err[(size_t)it]);
            if (kind == LoopKind::Fluid && rg[(size_t)it] != 0) {
                mexPrintf("Regridding on iteration: %d\tMin Jacobian: %.3f\n", it, (of2d_real)mj[(size_t)it]);
                rec.regrid_iteration.push_back(it);
                rec.regrid_minjac.push_back(mj[(size_t)it]);
            }
        }
        trace.levels.push_back(rec);
    }
    return true;
}

void ImageRegistration::run_level(LoopKind kind, Motion* level_motion, const Image* ref, Image* mov, IterativeSolver* slv, const int iterations, const dim d) {
    if (run_level_on_engine(kind, level_motion, ref, mov, iterations, d)) return;
    Image aux(d);
    Motion estimate(d);
    for (int refine = 0; refine < nrefine; refine++) {
        RegistrationTrace::Level rec;
        rec.scale = current_scale;
        rec.refine = refine;

        aux = *mov;
        aux.warp2d(*level_motion);
        Logger log(d, (unsigned int)iterations, verbose);
        if (kind != LoopKind::Demons) slv->set_derivatives(ref, &aux);

        for (int iter = 0; iter < iterations; iter++) {
            if (kind == LoopKind::Demons) slv->get_update(&estimate, ref, &aux);
            else slv->get_update(&estimate);
            rec.iterations++;
            if (kind == LoopKind::Fluid) {
                const OpticalFlowFluid* fl = static_cast<const OpticalFlowFluid*>(slv);
                rec.fluid_dt.push_back((double)fl->last_timestep());
                rec.fluid_maxabs.push_back((double)fl->last_maxabs());
            }

            log.update_error(&estimate);
            rec.error.push_back((double)log.get_error_at_current_iteration());
            if (log.get_error_at_current_iteration() < 0.001f && iter > 1) break;

            if (kind == LoopKind::Fluid) {
                of2d_real minjac = 0;
                of2d::check(of2d::jacobian((int)d.x, (int)d.y, estimate.device(), (of2d_real*)nullptr, &minjac));
                if (minjac < 0.5) {
                    mexPrintf("Regridding on iteration: %d\tMin Jacobian: %.3f\n", iter, minjac);
                    rec.regrid_iteration.push_back(iter);
                    rec.regrid_minjac.push_back((double)minjac);
                    level_motion->accumulate(estimate);
                    estimate.reset();
                    aux = *mov;
                    aux.warp2d(*level_motion);
                    slv->set_derivatives(ref, &aux);
                }
            }
        }
        level_motion->accumulate(estimate);
        estimate.reset();
        trace.levels.push_back(rec);
    }
}
