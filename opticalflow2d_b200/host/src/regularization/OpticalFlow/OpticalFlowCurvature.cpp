#include <src/regularization/OpticalFlow/OpticalFlowCurvature.h>

OpticalFlowCurvature::OpticalFlowCurvature(const dim dimin_, const of2d_real alpha_, const of2d_real tau_)
    : OpticalFlow(dimin_), alpha(alpha_), tau(tau_), plan(nullptr) {
    of2d::check(of2d_curvature_plan_create(of2d::context(), (int)dimin.x, (int)dimin.y, (double)alpha, (double)tau,
                                           sizeof(of2d_real) == sizeof(double), &plan));
}

OpticalFlowCurvature::~OpticalFlowCurvature() { of2d_curvature_plan_destroy(plan); }

// reference OpticalFlowCurvature.cpp:144-167
void OpticalFlowCurvature::get_update(Motion* motion, const Image*, const Image*) {
    of2d::check(of2d::curvature_step(plan, motion->device(), force->device_overwrite(), gradI->device(), It->device()));
    motion->swap_storage(*force);
}
