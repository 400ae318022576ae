#include <src/regularization/OpticalFlow/OpticalFlowDiffusion.h>

OpticalFlowDiffusion::OpticalFlowDiffusion(const dim dimin_, const of2d_real alpha_) : OpticalFlow(dimin_), alpha(alpha_) {}

OpticalFlowDiffusion::~OpticalFlowDiffusion() {}

// reference OpticalFlowDiffusion.cpp:43-84: its three passes (neighbour mean, force, update) are one
// kernel here; it is a Jacobi step, so it runs out of place into `force` and the buffers trade places.
void OpticalFlowDiffusion::get_update(Motion* motion, const Image*, const Image*) {
    of2d::check(of2d::diffusion_step((int)dimin.x, (int)dimin.y, motion->device(), force->device_overwrite(), gradI->device(), It->device(), alpha));
    motion->swap_storage(*force);
}
