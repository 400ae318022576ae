#include <src/ImageRegistrationFluid.h>

#include <src/regularization/OpticalFlow/OpticalFlowFluid.h>

// reference ImageRegistrationFluid.cpp:5-7
bool ImageRegistrationFluid::valid_regularisation_parameters(const Regularisation reg, const unsigned int nparams) const {
    return reg == Regularisation::Fluid && nparams >= 2 && nparams <= 3;
}

// reference :9-37
void ImageRegistrationFluid::set_solver(const Regularisation reg, const of2d_real* p, const unsigned int nparams) {
    if (!valid_regularisation_parameters(reg, nparams))
        throw std::invalid_argument("Invalid number of regularisation parameters for given regularisation method.\n");
    solver = new IterativeSolver*[nscales + 1]();
    for (int s = nscales; s >= 0; s--)
        solver[s] = nparams != 3 ? new OpticalFlowFluid(dimin[s], p[0], p[1]) : new OpticalFlowFluid(dimin[s], p[0], p[1], p[2]);
}

ImageRegistrationFluid::ImageRegistrationFluid(const dim dimin_, const int nscales_, const int* niter_, const int nrefine_, const Regularisation reg,
                                               const of2d_real* regparams, const unsigned int nparams, const Verbose verbose_)
    : ImageRegistration(dimin_, nscales_, niter_, nrefine_, reg, regparams, nparams, verbose_) {
    set_solver(reg, regparams, nparams);
}

ImageRegistrationFluid::~ImageRegistrationFluid() { release_solvers(); }

void ImageRegistrationFluid::estimate_motion_at_current_resolution(Motion* m, const Image* ref, Image* mov, IterativeSolver* slv, const int iterations,
                                                                   const dim d, const int) {
    run_level(LoopKind::Fluid, m, ref, mov, slv, iterations, d);
}
