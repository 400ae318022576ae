// Motion.h -- displacement field u(x) in pixels, array-of-structs {x, y} (reference
// src/Motion.h:7-52): composition, scaling-and-squaring exponential, norms, smoothing, resampling.
#ifndef OF2D_HOST_MOTION_H
#define OF2D_HOST_MOTION_H

#include <src/Field.h>
#include <src/Kernel.h>

class Motion : public Field<vector2d> {
public:
    Motion(const dim dimin);
    Motion(const Motion& mo);
    ~Motion();

    vector2d* get_motion() const;               // mutable host mirror
    void reset();
    void copy_motion_to_input(double* mo) const;   // planar doubles: x plane then y plane

    of2d_real norm() const;                     // mean Euclidean length
    of2d_real maxabs() const;                   // sqrt(max(y^2 + y^2)): the reference ignores x (src/Motion.cpp:54)

    void upSample(const Motion& mo);
    void downSample(const Motion& mo);

    void accumulate(const Motion& mo);          // u <- mo + u o (id + mo)

    void Neumann_boundaryconditions();
    void Dirichlet_boundaryconditions();

    void exp();                                 // scaling and squaring
    void convolute(const Kernel& kernel);

    Motion& operator=(const Motion& mo);
    Motion operator+(const Motion& mo) const;
    Motion& operator+=(const Motion& mo);
    Motion operator-(const Motion& mo) const;
    Motion& operator-=(const Motion& mo);
    Motion& operator*=(const of2d_real& val);
};

#endif
