#include <src/Motion.h>

#include <string>

#include <mex.h>

Motion::Motion(const dim dimin) : Field<vector2d>(dimin) {}
Motion::Motion(const Motion& mo) : Field<vector2d>(mo) {}
Motion::~Motion() {}

vector2d* Motion::get_motion() const { return get_field(); }

void Motion::reset() { clear(); }

// src/Motion.cpp:23-39
void Motion::copy_motion_to_input(double* mo) const {
    const size_t bytes = sizeof(double) * 2 * (size_t)sizein;
    of2d::Buffer staging(bytes);
    of2d::check(of2d::motion_to_planar(sizein, device(), static_cast<double*>(staging.device_discard())));
    of2d::check(of2d_d2h(of2d::context(), mo, staging.device_ro(), bytes));
}

// src/Motion.cpp:42-58
of2d_real Motion::norm() const {
    of2d_real v = 0;
    of2d::check(of2d::motion_norm(sizein, device(), &v));
    return v;
}
of2d_real Motion::maxabs() const {
    of2d_real v = 0;
    of2d::check(of2d::motion_maxabs(sizein, device(), &v));
    return v;
}

// src/Motion.cpp:61-111: resample, then rescale displacement magnitudes by the grid ratio
void Motion::upSample(const Motion& mo) {
    try {
        Field<vector2d>::upSample(mo);
    } catch (const std::invalid_argument& e) {
        const std::string msg = std::string("Error in Motion::upSample(const Motion& mo): ") + e.what() + "\n";
        mexErrMsgTxt(msg.c_str());
    }
    const dim src = mo.get_dimensions();
    of2d::check(of2d::scale_xy(sizein, (of2d_real)dimin.x / (of2d_real)src.x, (of2d_real)dimin.y / (of2d_real)src.y, device_mut()));
}
void Motion::downSample(const Motion& mo) {
    try {
        Field<vector2d>::downSample(mo);
    } catch (const std::invalid_argument& e) {
        const std::string msg = std::string("Error in Motion::downSample(const Motion& im): ") + e.what() + "\n";
        mexErrMsgTxt(msg.c_str());
    }
    const dim src = mo.get_dimensions();
    of2d::check(of2d::scale_xy(sizein, (of2d_real)dimin.x / (of2d_real)src.x, (of2d_real)dimin.y / (of2d_real)src.y, device_mut()));
}

// src/Motion.cpp:113-178
void Motion::accumulate(const Motion& mo) {
    if (dimin != mo.get_dimensions())
        throw std::invalid_argument("Error in Motion::accumulate(const Motion& mo): input dimensions should match target dimensions");
    Motion composed(dimin);
    of2d::check(of2d::compose((int)dimin.x, (int)dimin.y, device(), mo.device(), composed.device_overwrite()));
    swap_storage(composed);
}

// src/Motion.cpp:181-251 (never called by the reference's own drivers)
void Motion::Neumann_boundaryconditions() { of2d::check(of2d::boundary_conditions((int)dimin.x, (int)dimin.y, 0, device_mut())); }
void Motion::Dirichlet_boundaryconditions() { of2d::check(of2d::boundary_conditions((int)dimin.x, (int)dimin.y, 1, device_mut())); }

// src/Motion.cpp:253-277
void Motion::exp() {
    Motion scratch(dimin);
    of2d::check(of2d::motion_exp((int)dimin.x, (int)dimin.y, device_mut(), scratch.device_overwrite(), nullptr));
}

void Motion::convolute(const Kernel& kernel) { Field<vector2d>::convolute(kernel); }

Motion& Motion::operator=(const Motion& mo) {
    if (dimin != mo.get_dimensions())
        throw std::invalid_argument("Motion::operator=(const Motion& mo) input argument has to have same dimensions as target");
    if (this != &mo) assign(mo);
    return *this;
}

Motion Motion::operator+(const Motion& mo) const {
    Motion out(*this);
    out.Field<vector2d>::operator+=(mo);
    return out;
}
Motion& Motion::operator+=(const Motion& mo) {
    Field<vector2d>::operator+=(mo);
    return *this;
}
Motion Motion::operator-(const Motion& mo) const {
    Motion out(*this);
    out.Field<vector2d>::operator-=(mo);
    return out;
}
Motion& Motion::operator-=(const Motion& mo) {
    Field<vector2d>::operator-=(mo);
    return *this;
}
Motion& Motion::operator*=(const of2d_real& val) {
    Field<vector2d>::operator*=(val);
    return *this;
}
