// OpticalFlowCurvature.h -- semi-implicit curvature registration (reference OpticalFlowCurvature.h:8-42):
// u <- DCT^-1[ DCT[u - tau f] / (1 + tau alpha lap^2) ].  The reference's four fftw r2r plans are
// replaced by shared-memory DCT kernels; the plan object owns eigenvalue factors and twiddles.
#ifndef OF2D_HOST_OPTICAL_FLOW_CURVATURE_H
#define OF2D_HOST_OPTICAL_FLOW_CURVATURE_H

#include <src/regularization/OpticalFlow/OpticalFlow.h>

class OpticalFlowCurvature : public OpticalFlow {
public:
    OpticalFlowCurvature(const dim dimin, const of2d_real alpha, const of2d_real tau = 1.0f);
    ~OpticalFlowCurvature();

    void get_update(Motion* motion, const Image* Iref = NULL, const Image* Imov = NULL);

private:
    of2d_real alpha;
    of2d_real tau;
    of2d_curvature_plan* plan;
};

#endif
