#include <src/regularization/Demons/Demons.h>

Demons::Demons(const dim dimin_, const of2d_real sigma_i_, const of2d_real sigma_x_, const of2d_real sigma_diffusion_,
               const of2d_real sigma_fluid_, const unsigned int kernelwidth)
    : IterativeSolver(dimin_), sigma_i(sigma_i_), sigma_x(sigma_x_), sigma_diffusion(sigma_diffusion_), sigma_fluid(sigma_fluid_) {
    Iwar = new Image(dimin);
    correspondence = new Motion(dimin);
    scratch = new Motion(dimin);
    kernel_diffusion = new Kernel(kernelwidth);
    kernel_fluid = new Kernel(kernelwidth);
    kernel_diffusion->set_gaussian(sigma_diffusion);
    kernel_fluid->set_gaussian(sigma_fluid);
}

Demons::~Demons() {
    delete Iwar;
    delete correspondence;
    delete scratch;
    delete kernel_diffusion;
    delete kernel_fluid;
}

// reference Demons.cpp:34-63 on the derivatives currently stored in gradI / It (set_derivatives);
// get_update() uses the fused smoothed_correspondence() instead.
void Demons::demons_iteration(Motion*) {
    of2d::check(of2d::demons_correspondence((size_t)sizein, gradI->device(), It->device(), correspondence->device_overwrite(), sigma_i, sigma_x));
}

// DemonsThirions.cpp:18-30 / DemonsDiffeomorphic.cpp:15-27
void Demons::smoothed_correspondence(const Motion* motion, const Image* Iref, const Image* Imov) {
    of2d::check(of2d::demons_force((int)dimin.x, (int)dimin.y, Iref->device(), Imov->device(), motion->device(), scratch->device_overwrite(), sigma_i, sigma_x));
    const dim kd = kernel_fluid->get_dimensions();
    of2d::check(of2d::convolute(2, (int)dimin.x, (int)dimin.y, scratch->device(), correspondence->device_overwrite(), kernel_fluid->get_kernel(), (int)kd.x, (int)kd.y));
}

void Demons::smooth_motion(Motion* motion) {
    const dim kd = kernel_diffusion->get_dimensions();
    of2d::check(of2d::convolute(2, (int)dimin.x, (int)dimin.y, motion->device(), scratch->device_overwrite(), kernel_diffusion->get_kernel(), (int)kd.x, (int)kd.y));
    motion->swap_storage(*scratch);
}
