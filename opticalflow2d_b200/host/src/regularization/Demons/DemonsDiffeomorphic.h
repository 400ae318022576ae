// DemonsDiffeomorphic.h -- diffeomorphic Demons: the smoothed update is exponentiated by scaling
// and squaring before it is composed (reference src/regularization/Demons/DemonsDiffeomorphic.h:6-17).
#ifndef OF2D_HOST_DEMONS_DIFFEOMORPHIC_H
#define OF2D_HOST_DEMONS_DIFFEOMORPHIC_H

#include <src/regularization/Demons/Demons.h>

class DemonsDiffeomorphic : public Demons {
public:
    DemonsDiffeomorphic(const dim dimin, const of2d_real sigma_i = 1.0, const of2d_real sigma_x = 0.25, const of2d_real sigma_diffusion = 2.0,
                        const of2d_real sigma_fluid = 2.0, const unsigned int kernelwidth = 5);
    ~DemonsDiffeomorphic();

    void get_update(Motion* motion, const Image* Iref, const Image* Imov);

    int last_nsquares() const { return nsquares; }

private:
    int nsquares = 0;
};

#endif
