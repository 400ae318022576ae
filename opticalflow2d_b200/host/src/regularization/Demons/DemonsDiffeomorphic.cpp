#include <src/regularization/Demons/DemonsDiffeomorphic.h>

DemonsDiffeomorphic::DemonsDiffeomorphic(const dim dimin_, const of2d_real sigma_i_, const of2d_real sigma_x_, const of2d_real sigma_diffusion_,
                                         const of2d_real sigma_fluid_, const unsigned int kernelwidth)
    : Demons(dimin_, sigma_i_, sigma_x_, sigma_diffusion_, sigma_fluid_, kernelwidth) {}

DemonsDiffeomorphic::~DemonsDiffeomorphic() {}

// reference DemonsDiffeomorphic.cpp:15-35
void DemonsDiffeomorphic::get_update(Motion* motion, const Image* Iref, const Image* Imov) {
    smoothed_correspondence(motion, Iref, Imov);
    of2d::check(of2d::motion_exp((int)dimin.x, (int)dimin.y, correspondence->device_mut(), scratch->device_overwrite(), &nsquares));
    of2d::check(of2d::compose((int)dimin.x, (int)dimin.y, motion->device(), correspondence->device(), scratch->device_overwrite()));
    motion->swap_storage(*scratch);
    smooth_motion(motion);
}
