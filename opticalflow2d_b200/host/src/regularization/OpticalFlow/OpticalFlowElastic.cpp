#include <src/regularization/OpticalFlow/OpticalFlowElastic.h>

OpticalFlowElastic::OpticalFlowElastic(const dim dimin_, const of2d_real mu_, const of2d_real lambda_, const of2d_real omega_)
    : OpticalFlow(dimin_), mu(mu_), lambda(lambda_), omega(omega_) {}

OpticalFlowElastic::~OpticalFlowElastic() {}

// reference OpticalFlowElastic.cpp:13-55: force + sweep in one wavefront kernel (the force of a cell
// is evaluated from its still-old value when its row enters the wavefront)
void OpticalFlowElastic::get_update(Motion* motion, const Image*, const Image*) {
    of2d::check(of2d::elastic_step((int)dimin.x, (int)dimin.y, motion->device_mut(), gradI->device(), It->device(), mu, lambda, omega));
}
