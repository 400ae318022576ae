#include <src/ImageRegistration.h>

#include <cmath>
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
    : nscales(nscales_), nrefine(nrefine_), solver(nullptr), verbose(verbose_) {
    const int levels = nscales + 1;
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

Motion* ImageRegistration::get_estimated_motion() const { return motion[0]; }
void ImageRegistration::copy_estimated_motion(Motion& mo) const { mo = *motion[0]; }

void ImageRegistration::reset_state() {
    for (int s = nscales; s >= 0; s--) {
        motion[s]->reset();
        if (solver && solver[s]) solver[s]->reset_state();
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
void ImageRegistration::run_level(LoopKind kind, Motion* level_motion, const Image* ref, Image* mov, IterativeSolver* slv, const int iterations, const dim d) {
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
