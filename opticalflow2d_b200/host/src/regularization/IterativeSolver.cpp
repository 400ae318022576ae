#include <src/regularization/IterativeSolver.h>

IterativeSolver::IterativeSolver(const dim dimin_) : dimin(dimin_), step(1, dimin_.x), sizein(dimin_.x * dimin_.y) {
    gradI = new Motion(dimin);
    It = new Image(dimin);
}

IterativeSolver::~IterativeSolver() {
    delete gradI;
    delete It;
}

// reference IterativeSolver.cpp:22-44; the device kernel always produces the temporal term too, so
// the two public halves each run it into a scratch for the part they do not own
void IterativeSolver::spatial_derivative(Motion* grad_image, const Image* image) const {
    Image scratch(dimin);
    of2d::check(of2d::derivatives((int)dimin.x, (int)dimin.y, image->device(), image->device(), grad_image->device_overwrite(), scratch.device_overwrite()));
}

// reference IterativeSolver.cpp:46-51
void IterativeSolver::temporal_derivative(Image* It_, const Image* Iref, const Image* Imov) const {
    *It_ = *Imov - *Iref;
}

// reference IterativeSolver.cpp:53-56, one fused kernel
void IterativeSolver::set_derivatives(const Image* Iref, const Image* Imov) const {
    of2d::check(of2d::derivatives((int)dimin.x, (int)dimin.y, Iref->device(), Imov->device(), gradI->device_overwrite(), It->device_overwrite()));
}
