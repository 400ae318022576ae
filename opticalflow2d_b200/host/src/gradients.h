// gradients.h -- finite-difference stencils on a host array (reference src/gradients.h:6-81).  The
// solve itself evaluates these stencils inside the CUDA kernels (csrc/device_math.cuh); the host
// versions exist because the header is part of the public API surface and operate on the host
// mirrors handed out by Image::get_image() / Motion::get_motion().
#ifndef OF2D_HOST_GRADIENTS_H
#define OF2D_HOST_GRADIENTS_H

#include <src/coord2d.h>

namespace gradients {

// first derivatives: central inside, one-sided on the first / last sample
template <typename T>
inline T partial_x(T* field, const unsigned int idx, const unsigned int i, const dim& dimin) {
    if (i == 0) return field[idx + 1] - field[idx];
    if (i == dimin.x - 1) return field[idx] - field[idx - 1];
    return (field[idx + 1] - field[idx - 1]) / 2.0f;
}

template <typename T>
inline T partial_y(T* field, const unsigned int idx, const unsigned int j, const dim& dimin) {
    const unsigned int s = dimin.x;
    if (j == 0) return field[idx + s] - field[idx];
    if (j == dimin.y - 1) return field[idx] - field[idx - s];
    return (field[idx + s] - field[idx - s]) / 2.0f;
}

// second derivatives: 3-point inside, 4-point one-sided at the ends
template <typename T>
inline T partial_xx(T* field, const unsigned int idx, const unsigned int i, const dim& dimin) {
    if (i == 0) return field[idx] * 2 - field[idx + 1] * 5 + field[idx + 2] * 4 - field[idx + 3];
    if (i == dimin.x - 1) return field[idx - 3] * -1 + field[idx - 2] * 4 - field[idx - 1] * 5 + field[idx] * 2;
    return field[idx + 1] - field[idx] * 2 + field[idx - 1];
}

template <typename T>
inline T partial_yy(T* field, const unsigned int idx, const unsigned int j, const dim& dimin) {
    const unsigned int s = dimin.x;
    if (j == 0) return field[idx] * 2 - field[idx + s] * 5 + field[idx + 2 * s] * 4 - field[idx + 3 * s];
    if (j == dimin.y - 1) return field[idx - 3 * s] * -1 + field[idx - 2 * s] * 4 - field[idx - s] * 5 + field[idx] * 2;
    return field[idx + s] - field[idx] * 2 + field[idx - s];
}

template <typename T>
inline T partial_xy(T* field, const unsigned int idx, const unsigned int i, const unsigned int j, const dim& dimin) {
    const unsigned int s = dimin.x;
    if (i == 0 || j == 0 || i == dimin.x - 1 || j == dimin.y - 1) return T(0.0f);
    return (field[idx + 1 + s] - field[idx + 1 - s] - field[idx - 1 + s] + field[idx - 1 - s]) / 4.0f;
}

// mean of the four neighbours, zero on the border (the Jacobi average of Horn-Schunck)
template <typename T>
inline T qlaplacian(T* field, const unsigned int idx, const unsigned int i, const unsigned int j, const dim& dimin) {
    const unsigned int s = dimin.x;
    if (i == 0 || i == dimin.x - 1 || j == 0 || j == dimin.y - 1) return T(0.0f);
    return (field[idx - 1] + field[idx + 1] + field[idx - s] + field[idx + s]) / 4.0f;
}

}  // namespace gradients

#endif
