// DemonsThirions.h -- Thirion's Demons with compositive or additive accumulation
// (reference src/regularization/Demons/DemonsThirions.h:9-24).
#ifndef OF2D_HOST_DEMONS_THIRIONS_H
#define OF2D_HOST_DEMONS_THIRIONS_H

#include <src/SolverOptions.h>
#include <src/regularization/Demons/Demons.h>

class DemonsThirions : public Demons {
public:
    DemonsThirions(const dim dimin, const of2d_real sigma_i = 1.0, const of2d_real sigma_x = 0.25, const of2d_real sigma_diffusion = 2.0,
                   const of2d_real sigma_fluid = 2.0, const unsigned int kernelwidth = 5,
                   const MotionAccumulation motion_accumulation_method = MotionAccumulation::Composition);
    ~DemonsThirions();

    void get_update(Motion* motion, const Image* Iref, const Image* Imov);

private:
    MotionAccumulation motion_accumulation_method;
};

#endif
