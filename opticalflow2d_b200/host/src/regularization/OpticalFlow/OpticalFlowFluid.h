// OpticalFlowFluid.h -- viscous-fluid registration (reference OpticalFlowFluid.h:8-37): the SOR
// sweep acts on a persistent velocity field; the displacement follows the material derivative with
// an adaptive explicit Euler step dt = 0.65 / maxabs(R).
#ifndef OF2D_HOST_OPTICAL_FLOW_FLUID_H
#define OF2D_HOST_OPTICAL_FLOW_FLUID_H

#include <src/regularization/OpticalFlow/OpticalFlow.h>

class OpticalFlowFluid : public OpticalFlow {
public:
    OpticalFlowFluid(const dim dimin, const of2d_real mu, const of2d_real lambda, const of2d_real omega = 0.66);
    ~OpticalFlowFluid();

    void get_update(Motion* motion, const Image* Iref = NULL, const Image* Imov = NULL);

    void reset_state() { velocity->reset(); }

    of2d_real last_timestep() const { return timestep; }
    of2d_real last_maxabs() const { return maxabs_increment; }

private:
    of2d_real mu;
    of2d_real lambda;
    of2d_real omega;

    of2d_real timestep;
    of2d_real maxabs_increment = 0;
    const of2d_real dumax = 0.65f;

    Motion* velocity;    // never reset: carried across regrids, refines and calls, as in the reference
    Motion* increment;
};

#endif
