// Image.h -- scalar image on the registration grid (reference src/Image.h:8-53): MEX I/O in
// double, bilinear warping by a Motion, Jacobian determinant map, reductions, resampling.
#ifndef OF2D_HOST_IMAGE_H
#define OF2D_HOST_IMAGE_H

#include <src/Field.h>
#include <src/Kernel.h>
#include <src/Motion.h>

class Image : public Field<of2d_real> {
public:
    Image(const dim dimin);
    Image(const Image& im);
    ~Image();

    void set_image(const double* im);          // caller-owned column-major doubles -> field
    of2d_real* get_image() const;              // mutable host mirror
    void copy_image_to_input(double* im) const;

    void upSample(const Image& im);
    void downSample(const Image& im);

    of2d_real sum() const;
    of2d_real max() const;
    of2d_real min() const;
    void normalize();

    void warp2d(const Motion& mo);             // I(x) <- I(x + u(x)), bilinear, out-of-range pixels keep their value
    void convolute(const Kernel& kernel);
    void jacobian(const Motion& mo);           // det(Id + grad u) per pixel

    Image& operator=(const Image& im);
    Image operator+(const Image& im) const;
    Image& operator+=(const Image& im);
    Image operator-(const Image& im) const;
    Image& operator-=(const Image& im);
    Image& operator*=(const of2d_real& val);
};

#endif
