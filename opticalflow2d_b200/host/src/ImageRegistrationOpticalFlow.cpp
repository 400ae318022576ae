#include <src/ImageRegistrationOpticalFlow.h>

#include <src/regularization/OpticalFlow/OpticalFlowCurvature.h>
#include <src/regularization/OpticalFlow/OpticalFlowDiffusion.h>
#include <src/regularization/OpticalFlow/OpticalFlowElastic.h>

// arities as reference ImageRegistrationOpticalFlow.cpp:8-12
bool ImageRegistrationOpticalFlow::valid_regularisation_parameters(const Regularisation reg, const unsigned int nparams) const {
    if (reg == Regularisation::Diffusion) return nparams == 1;
    if (reg == Regularisation::Curvature) return nparams >= 1 && nparams <= 2;
    if (reg == Regularisation::Elastic) return nparams >= 2 && nparams <= 3;
    return false;
}

// reference :14-69
void ImageRegistrationOpticalFlow::set_solver(const Regularisation reg, const of2d_real* p, const unsigned int nparams) {
    if (!valid_regularisation_parameters(reg, nparams))
        throw std::invalid_argument("Invalid number of regularisation parameters for given regularisation method.\n");
    solver = new IterativeSolver*[nscales + 1]();
    for (int s = nscales; s >= 0; s--) {
        if (reg == Regularisation::Diffusion) {
            solver[s] = new OpticalFlowDiffusion(dimin[s], p[0]);
        } else if (reg == Regularisation::Curvature) {
            solver[s] = nparams == 1 ? new OpticalFlowCurvature(dimin[s], p[0]) : new OpticalFlowCurvature(dimin[s], p[0], p[1]);
        } else {
            solver[s] = nparams != 3 ? new OpticalFlowElastic(dimin[s], p[0], p[1]) : new OpticalFlowElastic(dimin[s], p[0], p[1], p[2]);
        }
    }
}

ImageRegistrationOpticalFlow::ImageRegistrationOpticalFlow(const dim dimin_, const int nscales_, const int* niter_, const int nrefine_,
                                                           const Regularisation reg, const of2d_real* regparams, const unsigned int nparams,
                                                           const Verbose verbose_)
    : ImageRegistration(dimin_, nscales_, niter_, nrefine_, reg, regparams, nparams, verbose_) {
    set_solver(reg, regparams, nparams);
}

ImageRegistrationOpticalFlow::~ImageRegistrationOpticalFlow() { release_solvers(); }

void ImageRegistrationOpticalFlow::estimate_motion_at_current_resolution(Motion* m, const Image* ref, Image* mov, IterativeSolver* slv, const int iterations,
                                                                         const dim d, const int) {
    run_level(LoopKind::OpticalFlow, m, ref, mov, slv, iterations, d);
}
