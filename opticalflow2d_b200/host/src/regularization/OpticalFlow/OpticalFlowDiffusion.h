// OpticalFlowDiffusion.h -- Horn-Schunck with Jacobi relaxation (reference OpticalFlowDiffusion.h:9-30):
// ubar = mean of the 4 neighbours (0 on the border), f = force(ubar), u <- ubar - f / (alpha^2 + |gradI|^2).
#ifndef OF2D_HOST_OPTICAL_FLOW_DIFFUSION_H
#define OF2D_HOST_OPTICAL_FLOW_DIFFUSION_H

#include <src/regularization/OpticalFlow/OpticalFlow.h>

class OpticalFlowDiffusion : public OpticalFlow {
public:
    OpticalFlowDiffusion(const dim dimin, const of2d_real alpha);
    ~OpticalFlowDiffusion();

    void get_update(Motion* motion, const Image* Iref = NULL, const Image* Imov = NULL);

    of2d_real regularisation() const { return alpha; }

private:
    of2d_real alpha;
};

#endif
