// Kernel.h -- small dense convolution kernel with double weights (reference src/Kernel.h:6-28).
// Lives on the host; Field::convolute hands the weights to the CUDA convolution.
#ifndef OF2D_HOST_KERNEL_H
#define OF2D_HOST_KERNEL_H

#include <vector>

#include <src/coord2d.h>

class Kernel {
public:
    Kernel(const unsigned int kernelwidth);
    Kernel(const dim dimkernel);
    Kernel(const Kernel&) = default;
    ~Kernel();

    dim get_dimensions() const;
    dim get_step() const;
    unsigned int get_size() const;
    double* get_kernel() const;

    void set_gaussian(const of2d_real sigma);
    void set_average();

private:
    dim extent_;
    mutable std::vector<double> weights_;
};

#endif
