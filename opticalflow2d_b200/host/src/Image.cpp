#include <src/Image.h>

#include <string>
#include <vector>

#include <mex.h>

namespace {
// the reference converts dimension mismatches inside the arithmetic wrappers into mexErrMsgTxt
// (src/Image.cpp:241-293) but lets them escape as std::invalid_argument from warp2d / jacobian / =
template <class F>
void mex_guard(const char* where, F&& body) {
    try {
        body();
    } catch (const std::invalid_argument& e) {
        const std::string msg = std::string("Error in ") + where + " " + e.what() + "\n";
        mexErrMsgTxt(msg.c_str());
    }
}
}  // namespace

Image::Image(const dim dimin) : Field<of2d_real>(dimin) {}
Image::Image(const Image& im) : Field<of2d_real>(im) {}
Image::~Image() {}

// src/Image.cpp:15-29: the doubles are staged in HBM and narrowed there
void Image::set_image(const double* im) {
    of2d::Buffer staging(sizeof(double) * (size_t)sizein);
    of2d::check(of2d_h2d(of2d::context(), staging.device_discard(), im, sizeof(double) * (size_t)sizein));
    of2d::check(of2d::image_from_double(sizein, static_cast<const double*>(staging.device_ro()), device_overwrite()));
    of2d::check(of2d_ctx_sync(of2d::context()));   // `im` is caller-owned and pageable: do not outlive the call
}

of2d_real* Image::get_image() const { return get_field(); }

// src/Image.cpp:36-50
void Image::copy_image_to_input(double* im) const {
    of2d::Buffer staging(sizeof(double) * (size_t)sizein);
    of2d::check(of2d::image_to_double(sizein, device(), static_cast<double*>(staging.device_discard())));
    of2d::check(of2d_d2h(of2d::context(), im, staging.device_ro(), sizeof(double) * (size_t)sizein));
}

void Image::upSample(const Image& im) {
    mex_guard("Image::upSample(const Image& im):", [&] { Field<of2d_real>::upSample(im); });
}
void Image::downSample(const Image& im) {
    mex_guard("Image::downSample(const Image& im):", [&] { Field<of2d_real>::downSample(im); });
}

// src/Image.cpp:78-104.  sum accumulates in a wider type than the reference's float (parallel order).
of2d_real Image::sum() const {
    of2d_real s = 0;
    of2d::check(of2d::image_stats(sizein, device(), &s, nullptr, nullptr));
    return s;
}
of2d_real Image::max() const {
    of2d_real m = 0;
    of2d::check(of2d::image_stats(sizein, device(), nullptr, &m, nullptr));
    return m;
}
of2d_real Image::min() const {
    of2d_real m = 0;
    of2d::check(of2d::image_stats(sizein, device(), nullptr, nullptr, &m));
    return m;
}

// src/Image.cpp:107-116
void Image::normalize() {
    of2d_real mx = 0, mn = 0;
    of2d::check(of2d::image_stats(sizein, device(), nullptr, &mx, &mn));
    of2d::check(of2d::image_normalize(sizein, mn, mx, device_mut()));
}

// src/Image.cpp:119-182
void Image::warp2d(const Motion& mo) {
    if (dimin != mo.get_dimensions())
        throw std::invalid_argument("Error in Image::warp2d(const Motion& mo): input dimensions have to be the same as target");
    Image warped(dimin);
    of2d::check(of2d::warp2d((int)dimin.x, (int)dimin.y, device(), mo.device(), warped.device_overwrite()));
    swap_storage(warped);
}

void Image::convolute(const Kernel& kernel) { Field<of2d_real>::convolute(kernel); }

// src/Image.cpp:189-218
void Image::jacobian(const Motion& mo) {
    if (dimin != mo.get_dimensions())
        throw std::invalid_argument("Error in Image::warp2d(const Motion& mo): input dimensions have to be the same as target");
    of2d::check(of2d::jacobian((int)dimin.x, (int)dimin.y, mo.device(), device_overwrite(), nullptr));
}

Image& Image::operator=(const Image& im) {
    if (dimin != im.get_dimensions())
        throw std::invalid_argument("Image::operator=(const Image& im) input argument has to have same dimensions as target");
    if (this != &im) assign(im);
    return *this;
}

Image Image::operator+(const Image& im) const {
    Image out(*this);
    mex_guard("Image::operator+(const Image& im)", [&] { out.Field<of2d_real>::operator+=(im); });
    return out;
}
Image& Image::operator+=(const Image& im) {
    mex_guard("Image::operator+=(const Image& im)", [&] { Field<of2d_real>::operator+=(im); });
    return *this;
}
Image Image::operator-(const Image& im) const {
    Image out(*this);
    mex_guard("Image::operator-(const Image& im)", [&] { out.Field<of2d_real>::operator-=(im); });
    return out;
}
Image& Image::operator-=(const Image& im) {
    mex_guard("Image::operator-=(const Image& im)", [&] { Field<of2d_real>::operator-=(im); });
    return *this;
}
Image& Image::operator*=(const of2d_real& val) {
    Field<of2d_real>::operator*=(val);
    return *this;
}
