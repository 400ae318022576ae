#include <src/ImageRegistrationDemons.h>

#include <src/regularization/Demons/DemonsDiffeomorphic.h>
#include <src/regularization/Demons/DemonsThirions.h>

// arities as reference ImageRegistrationDemons.cpp:7-10
bool ImageRegistrationDemons::valid_regularisation_parameters(const Regularisation reg, const unsigned int nparams) const {
    if (reg == Regularisation::ThirionsDemons) return nparams == 6;
    if (reg == Regularisation::DiffeomorphicDemons) return nparams == 5;
    return false;
}

// reference :12-58: regparams = {sigma_i, sigma_x, sigma_diffusion, sigma_fluid, kernel width[, accumulation]}
void ImageRegistrationDemons::set_solver(const Regularisation reg, const of2d_real* p, const unsigned int nparams) {
    if (!valid_regularisation_parameters(reg, nparams))
        throw std::invalid_argument("Invalid number of regularisation parameters for given regularisation method.\n");
    solver = new IterativeSolver*[nscales + 1]();
    for (int s = nscales; s >= 0; s--) {
        const unsigned int width = static_cast<unsigned int>(p[4]);
        if (reg == Regularisation::ThirionsDemons)
            solver[s] = new DemonsThirions(dimin[s], p[0], p[1], p[2], p[3], width, static_cast<MotionAccumulation>((int)p[5]));
        else
            solver[s] = new DemonsDiffeomorphic(dimin[s], p[0], p[1], p[2], p[3], width);
    }
}

ImageRegistrationDemons::ImageRegistrationDemons(const dim dimin_, const int nscales_, const int* niter_, const int nrefine_, const Regularisation reg,
                                                 const of2d_real* regparams, const unsigned int nparams, const Verbose verbose_)
    : ImageRegistration(dimin_, nscales_, niter_, nrefine_, reg, regparams, nparams, verbose_) {
    set_solver(reg, regparams, nparams);
}

ImageRegistrationDemons::~ImageRegistrationDemons() { release_solvers(); }

void ImageRegistrationDemons::estimate_motion_at_current_resolution(Motion* m, const Image* ref, Image* mov, IterativeSolver* slv, const int iterations,
                                                                    const dim d, const int) {
    run_level(LoopKind::Demons, m, ref, mov, slv, iterations, d);
}
