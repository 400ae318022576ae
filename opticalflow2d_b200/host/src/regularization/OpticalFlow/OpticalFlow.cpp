#include <src/regularization/OpticalFlow/OpticalFlow.h>

OpticalFlow::OpticalFlow(const dim dimin_) : IterativeSolver(dimin_) { force = new Motion(dimin); }

OpticalFlow::~OpticalFlow() { delete force; }

// reference OpticalFlow.cpp:15-39.  The solvers fuse the force into their update kernels; this
// stand-alone form is kept because get_force() is part of the public class.
void OpticalFlow::get_force(Motion* f, const Motion* motion) const {
    of2d::check(of2d::lssd_force((int)dimin.x, (int)dimin.y, gradI->device(), It->device(), motion->device(), f->device_overwrite()));
}
