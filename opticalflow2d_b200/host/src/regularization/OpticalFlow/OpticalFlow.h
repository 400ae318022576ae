// OpticalFlow.h -- base of the optical-flow family: owns the L-SSD force buffer
// f = gradI (It + u . gradI) (reference src/regularization/OpticalFlow/OpticalFlow.h:9-24).
#ifndef OF2D_HOST_OPTICAL_FLOW_H
#define OF2D_HOST_OPTICAL_FLOW_H

#include <src/Image.h>
#include <src/Motion.h>
#include <src/coord2d.h>
#include <src/regularization/IterativeSolver.h>

class OpticalFlow : public IterativeSolver {
public:
    OpticalFlow(const dim dimin);
    ~OpticalFlow();

    void get_force(Motion* force, const Motion* motion) const;

    virtual void get_update(Motion* motion, const Image* Iref = NULL, const Image* Imov = NULL) {}

protected:
    Motion* force;   // also the ping-pong partner of the out-of-place update kernels
};

#endif
