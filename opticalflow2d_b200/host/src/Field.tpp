// Field.tpp -- template bodies of Field<T>.
#include <string>

template <class T>
Field<T>::Field(const dim dimin_)
    : dimin(dimin_), sizein(dimin_.x * dimin_.y), step(1, dimin_.x), storage(sizeof(T) * (size_t)dimin_.x * dimin_.y) {}

template <class T>
Field<T>::Field(const Field<T>& other)
    : dimin(other.dimin), sizein(other.sizein), step(other.step), storage(other.storage) {}

template <class T>
Field<T>::~Field() {}

template <class T>
T* Field<T>::get_field() const { return static_cast<T*>(storage.host()); }

template <class T>
dim Field<T>::get_dimensions() const { return dimin; }

template <class T>
dim Field<T>::get_step() const { return step; }

template <class T>
unsigned int Field<T>::get_size() const { return sizein; }

template <class T>
void Field<T>::require_same_grid(const Field<T>& other, const char* what) const {
    if (dimin != other.dimin) throw std::invalid_argument(what);
}

template <class T>
void Field<T>::swap_storage(Field<T>& other) {
    require_same_grid(other, "swap_storage: grids differ");
    storage.swap(other.storage);
}

template <class T>
void Field<T>::assign(const Field<T>& other) {
    require_same_grid(other, "input argument has to have same dimensions as target");
    storage.copy_from(other.storage);
}

template <class T>
void Field<T>::clear() { storage.zero(); }

template <class T>
void Field<T>::downSample(const Field<T>& src) {
    if (dimin.x > src.dimin.x || dimin.y > src.dimin.y)
        throw std::invalid_argument("Error in Field<T>::downSample(const FIeld<T>& fieldin): input has to have same dimensions as target");
    if (dimin.x == 0 || dimin.y == 0) throw std::runtime_error("Divide by zero exception");
    of2d::check(of2d::downsample(components, (int)src.dimin.x, (int)src.dimin.y, src.device(), (int)dimin.x, (int)dimin.y, device_mut()));
}

template <class T>
void Field<T>::upSample(const Field<T>& src) {
    if (dimin.x < src.dimin.x || dimin.y < src.dimin.y)
        throw std::invalid_argument("Error in Field<T>::downSample(const FIeld<T>& fieldin): input has to have same dimensions as target");
    of2d::check(of2d::upsample(components, (int)src.dimin.x, (int)src.dimin.y, src.device(), (int)dimin.x, (int)dimin.y, device_mut()));
}

template <class T>
void Field<T>::convolute(const Kernel& kernel) {
    Field<T> result(dimin);
    const dim kd = kernel.get_dimensions();
    of2d::check(of2d::convolute(components, (int)dimin.x, (int)dimin.y, device(), result.device_overwrite(), kernel.get_kernel(), (int)kd.x, (int)kd.y));
    storage.swap(result.storage);
}

template <class T>
Field<T>& Field<T>::operator+=(const Field<T>& rhs) {
    require_same_grid(rhs, "input argument has to have same dimensions as target");
    of2d::check(of2d::axpy((size_t)sizein * components, (of2d_real)1, rhs.device(), device_mut()));
    return *this;
}

template <class T>
Field<T> Field<T>::operator+(const Field<T>& rhs) const {
    require_same_grid(rhs, "input argument has to have same dimensions as target");
    Field<T> out(*this);
    out += rhs;
    return out;
}

template <class T>
Field<T>& Field<T>::operator-=(const Field<T>& rhs) {
    require_same_grid(rhs, "input argument has to have same dimensions as target");
    of2d::check(of2d::axpy((size_t)sizein * components, (of2d_real)-1, rhs.device(), device_mut()));
    return *this;
}

template <class T>
Field<T> Field<T>::operator-(const Field<T>& rhs) const {
    require_same_grid(rhs, "input argument has to have same dimensions as target");
    Field<T> out(*this);
    out -= rhs;
    return out;
}

template <class T>
Field<T>& Field<T>::operator*=(const of2d_real& val) {
    of2d::check(of2d::scale((size_t)sizein * components, val, device_mut()));
    return *this;
}
