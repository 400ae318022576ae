// OpticalFlowElastic.h -- Navier-Lame (linear elastic) regularisation (reference OpticalFlowElastic.h:7-22):
// per iteration the force is taken from the current u, then ONE in-place lexicographic SOR sweep.
#ifndef OF2D_HOST_OPTICAL_FLOW_ELASTIC_H
#define OF2D_HOST_OPTICAL_FLOW_ELASTIC_H

#include <src/regularization/OpticalFlow/OpticalFlow.h>

class OpticalFlowElastic : public OpticalFlow {
public:
    OpticalFlowElastic(const dim dimin, const of2d_real mu, const of2d_real lambda, const of2d_real omega = 0.66f);
    ~OpticalFlowElastic();

    void get_update(Motion* motion, const Image* Iref = NULL, const Image* Imov = NULL);

private:
    of2d_real mu;
    of2d_real lambda;
    of2d_real omega;
};

#endif
