// IterativeSolver.h -- the operator seam between the registration drivers and the per-iteration
// solvers (reference src/regularization/IterativeSolver.h:8-31): drivers only call
// set_derivatives() and the virtual get_update().
#ifndef OF2D_HOST_ITERATIVE_SOLVER_H
#define OF2D_HOST_ITERATIVE_SOLVER_H

#include <src/Image.h>
#include <src/Motion.h>
#include <src/coord2d.h>

class IterativeSolver {
public:
    IterativeSolver(const dim dimin);
    virtual ~IterativeSolver();   // virtual here (the reference's is not, which leaks derived state: SURVEY Q16)

    void spatial_derivative(Motion* grad_image, const Image* image) const;
    void temporal_derivative(Image* It, const Image* Iref, const Image* Imov) const;
    void set_derivatives(const Image* Iref, const Image* Imov) const;   // gradI = grad(Imov), It = Imov - Iref

    // one iteration of the scheme, in place on `motion`
    virtual void get_update(Motion* motion, const Image* Iref = NULL, const Image* Imov = NULL) {}

    // extension: forget state carried between calls (only the fluid velocity)
    virtual void reset_state() {}

protected:
    dim dimin;
    dim step;
    unsigned int sizein;

    Motion* gradI;
    Image* It;
};

#endif
