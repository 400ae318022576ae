// ImageRegistrationOpticalFlow.h -- registration driver of the OpticalFlow family
// (reference src/ImageRegistrationOpticalFlow.h): validates the parameter count, builds one solver per
// pyramid level and runs the per-level loop.
#ifndef OF2D_HOST_IMAGE_REGISTRATION_OPTICALFLOW_H
#define OF2D_HOST_IMAGE_REGISTRATION_OPTICALFLOW_H

#include <src/ImageRegistration.h>

class ImageRegistrationOpticalFlow : public ImageRegistration {
public:
    ImageRegistrationOpticalFlow(const dim dimin, const int nscales, const int* niter, const int nrefine, const Regularisation reg,
                            const of2d_real* regparams, const unsigned int nparams, const Verbose verbose);
    ~ImageRegistrationOpticalFlow();

private:
    bool valid_regularisation_parameters(const Regularisation reg, const unsigned int nparams) const;
    void set_solver(const Regularisation reg, const of2d_real* regparams, const unsigned int nparams);
    void estimate_motion_at_current_resolution(Motion* motion, const Image* Iref, Image* Imov, IterativeSolver* solver, const int niter,
                                               const dim dimin, const int sizein);
};

#endif
