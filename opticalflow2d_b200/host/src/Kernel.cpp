#include <src/Kernel.h>

#include <cmath>

Kernel::Kernel(const dim dimkernel) : extent_(dimkernel), weights_((size_t)dimkernel.x * dimkernel.y, 0.0) {}

Kernel::Kernel(const unsigned int kernelwidth) : extent_(kernelwidth, kernelwidth), weights_((size_t)kernelwidth * kernelwidth, 0.0) {}

Kernel::~Kernel() {}

dim Kernel::get_dimensions() const { return extent_; }
dim Kernel::get_step() const { return dim(1, extent_.x); }
unsigned int Kernel::get_size() const { return extent_.x * extent_.y; }
double* Kernel::get_kernel() const { return weights_.data(); }

// Gaussian weights as reference src/Kernel.cpp:45-73: the exponential is evaluated in the field
// precision (float exp on a float argument in the fp32 build), stored and normalised in double.
void Kernel::set_gaussian(const of2d_real sigma) {
    const int w = (int)extent_.x, h = (int)extent_.y;
    const int cx = (int)((extent_.x - 1) / 2), cy = (int)((extent_.y - 1) / 2);
    double total = 0;
    for (int i = 0; i < w; i++) {
        for (int j = 0; j < h; j++) {
            const of2d_real arg = -((i - cx) * (i - cx) + (j - cy) * (j - cy)) / (2 * sigma * sigma);
            const double v = std::exp(arg);
            weights_[(size_t)i + (size_t)j * w] = v;
            total += v;
        }
    }
    for (double& v : weights_) v /= total;
}

// reference src/Kernel.cpp:75-82
void Kernel::set_average() {
    const of2d_real v = 1.0f / (of2d_real)get_size();
    for (double& wgt : weights_) wgt = v;
}
