#include <src/regularization/OpticalFlow/OpticalFlowFluid.h>

#include <mex.h>

OpticalFlowFluid::OpticalFlowFluid(const dim dimin_, const of2d_real mu_, const of2d_real lambda_, const of2d_real omega_)
    : OpticalFlow(dimin_), mu(mu_), lambda(lambda_), omega(omega_), timestep(0) {
    velocity = new Motion(dimin);
    increment = new Motion(dimin);
}

OpticalFlowFluid::~OpticalFlowFluid() {
    delete velocity;
    delete increment;
}

// reference OpticalFlowFluid.cpp:123-140; the time-step line is printed on every iteration whatever
// the verbosity, as the reference does (:94)
void OpticalFlowFluid::get_update(Motion* motion, const Image*, const Image*) {
    of2d_real maxabs = 0, dt = 0;
    of2d::check(of2d::fluid_step((int)dimin.x, (int)dimin.y, motion->device_mut(), velocity->device_mut(), increment->device_mut(),
                                 gradI->device(), It->device(), mu, lambda, omega, &maxabs, &dt));
    timestep = dt;
    maxabs_increment = maxabs;
    mexPrintf("Dumax: %.3f\tMaxabs increment: %.3f\t Timestep: %.3f\n", dumax, maxabs, timestep);
}
