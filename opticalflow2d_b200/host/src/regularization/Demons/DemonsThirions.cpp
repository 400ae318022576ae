#include <src/regularization/Demons/DemonsThirions.h>

DemonsThirions::DemonsThirions(const dim dimin_, const of2d_real sigma_i_, const of2d_real sigma_x_, const of2d_real sigma_diffusion_,
                               const of2d_real sigma_fluid_, const unsigned int kernelwidth, const MotionAccumulation method)
    : Demons(dimin_, sigma_i_, sigma_x_, sigma_diffusion_, sigma_fluid_, kernelwidth), motion_accumulation_method(method) {}

DemonsThirions::~DemonsThirions() {}

// reference DemonsThirions.cpp:18-42
void DemonsThirions::get_update(Motion* motion, const Image* Iref, const Image* Imov) {
    smoothed_correspondence(motion, Iref, Imov);
    if (motion_accumulation_method == MotionAccumulation::Composition) {
        of2d::check(of2d::compose((int)dimin.x, (int)dimin.y, motion->device(), correspondence->device(), scratch->device_overwrite()));
        motion->swap_storage(*scratch);
    } else if (motion_accumulation_method == MotionAccumulation::Addition) {
        *motion += *correspondence;
    }
    smooth_motion(motion);
}
