// Field.h -- owning 2-D array of T over a dimx x dimy grid, column-major with x fastest
// (idx = i + j*dimx), the base of Image (T = real) and Motion (T = vector2d).  Public surface as
// the reference's src/Field.h:8-43; storage lives in HBM behind of2d::Buffer and every operation
// is a kernel launch through the C ABI.
#ifndef OF2D_HOST_FIELD_H
#define OF2D_HOST_FIELD_H

#include <src/DeviceRuntime.h>
#include <src/Kernel.h>
#include <src/coord2d.h>

template <class T>
class Field {
public:
    Field(const dim dimin);
    Field(const Field<T>& fieldin);
    virtual ~Field();

    dim get_dimensions() const;
    dim get_step() const;
    unsigned int get_size() const;

    // bilinear up-sampling / box down-sampling from another grid (reference src/Field.tpp:76-206)
    void upSample(const Field<T>& fieldin);
    void downSample(const Field<T>& fieldin);

    // dense correlation with `kernel`, bounds tested on the linear index (reference src/Field.tpp:210-269)
    void convolute(const Kernel& kernel);

    virtual Field<T> operator+(const Field<T>& fieldin) const;
    virtual Field<T>& operator+=(const Field<T>& fieldin);
    virtual Field<T> operator-(const Field<T>& fieldin) const;
    virtual Field<T>& operator-=(const Field<T>& fieldin);
    virtual Field<T>& operator*=(const of2d_real& val);

    // ---- device-side access (extension; used by the solvers and drivers) ----
    static constexpr int components = (int)(sizeof(T) / sizeof(of2d_real));
    const of2d_real* device() const { return static_cast<const of2d_real*>(storage.device_ro()); }
    of2d_real* device_mut() { return static_cast<of2d_real*>(storage.device_rw()); }
    of2d_real* device_overwrite() { return static_cast<of2d_real*>(storage.device_discard()); }
    void swap_storage(Field<T>& other);
    void assign(const Field<T>& other);   // deep copy, dimensions must match
    void clear();

protected:
    T* get_field() const;   // host mirror (synchronised on demand)

    void require_same_grid(const Field<T>& other, const char* what) const;

    dim dimin;
    unsigned int sizein;
    dim step;
    mutable of2d::Buffer storage;
};

#include <src/Field.tpp>

#endif
