// Demons.h -- base of the Demons family (reference src/regularization/Demons/Demons.h:7-51): the
// correspondence update c = -gradI It / (|gradI|^2 + It^2 sigma_i^2 / sigma_x^2) and the two Gaussian
// kernels (fluid-like smoothing of c, diffusion-like smoothing of u).
#ifndef OF2D_HOST_DEMONS_H
#define OF2D_HOST_DEMONS_H

#include <src/Kernel.h>
#include <src/regularization/IterativeSolver.h>

class Demons : public IterativeSolver {
public:
    Demons(const dim dimin, const of2d_real sigma_i = 1.0, const of2d_real sigma_x = 0.25, const of2d_real sigma_diffusion = 2.0,
           const of2d_real sigma_fluid = 2.0, const unsigned int kernelwidth = 5);
    ~Demons();

    virtual void get_update(Motion* motion, const Image* Iref, const Image* Imov) {}

protected:
    // c from the derivatives currently held in gradI / It
    void demons_iteration(Motion* motion);

    // warp + derivatives + force in one kernel, then c <- K_fluid * c; result in `correspondence`
    void smoothed_correspondence(const Motion* motion, const Image* Iref, const Image* Imov);
    // u <- K_diffusion * u
    void smooth_motion(Motion* motion);

    Image* Iwar;
    Motion* correspondence;
    Motion* scratch;

    of2d_real sigma_i;
    of2d_real sigma_x;
    of2d_real sigma_diffusion;
    of2d_real sigma_fluid;

    Kernel* kernel_diffusion;
    Kernel* kernel_fluid;
};

#endif
