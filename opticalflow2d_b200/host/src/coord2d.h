// coord2d.h -- 2-component value type of the registration API (dim = unsigned extents,
// vector2d = displacement in pixels).  Mirrors the public surface of the reference's
// src/coord2d.h:7-149, including its observable behaviour that every division by a zero
// component throws std::runtime_error("Divide by zero exception") (src/coord2d.h:95-127).
//
// OF2D_REAL selects the scalar type of all fields: float reproduces the reference as written,
// double is the fp64 mode (same effect as the float->double build used as the fp64 oracle).
#ifndef OF2D_HOST_COORD2D_H
#define OF2D_HOST_COORD2D_H

#include <stdexcept>

#ifndef OF2D_REAL
#define OF2D_REAL float
#endif
typedef OF2D_REAL of2d_real;

template <class T>
class coord2d {
public:
    T x;
    T y;

    coord2d() : x(T(0)), y(T(0)) {}
    coord2d(const T x_, const T y_) : x(x_), y(y_) {}
    coord2d(const T a) : x(a), y(a) {}

    coord2d<T>& operator=(const coord2d<T>& c) = default;
    coord2d(const coord2d<T>& c) = default;
    coord2d<T>& operator=(const T& a) { x = a; y = a; return *this; }

    coord2d<T> operator+(const coord2d<T>& c) const { return coord2d<T>(x + c.x, y + c.y); }
    coord2d<T> operator+(const T& a) const { return coord2d<T>(x + a, y + a); }
    coord2d<T> operator-(const coord2d<T>& c) const { return coord2d<T>(x - c.x, y - c.y); }
    coord2d<T> operator-(const T& a) const { return coord2d<T>(x - a, y - a); }
    coord2d<T> operator*(const T& a) const { return coord2d<T>(x * a, y * a); }

    coord2d<T>& operator+=(const coord2d<T>& c) { x += c.x; y += c.y; return *this; }
    coord2d<T>& operator+=(const T& a) { x += a; y += a; return *this; }
    coord2d<T>& operator-=(const coord2d<T>& c) { x -= c.x; y -= c.y; return *this; }
    coord2d<T>& operator-=(const T& a) { x -= a; y -= a; return *this; }
    coord2d<T>& operator*=(const T& a) { x *= a; y *= a; return *this; }

    coord2d<T> operator/(const T& a) const {
        refuse_zero(a == 0);
        return coord2d<T>(x / a, y / a);
    }
    template <class D>
    coord2d<T> operator/(const coord2d<D>& a) const {
        refuse_zero(a.x == 0 || a.y == 0);
        return coord2d<T>(x / a.x, y / a.y);
    }
    coord2d<T>& operator/=(const T& a) {
        refuse_zero(a == 0);
        x /= a; y /= a;
        return *this;
    }
    template <class D>
    coord2d<T>& operator/=(const coord2d<D>& a) {
        refuse_zero(a.x == 0 || a.y == 0);
        x /= a.x; y /= a.y;
        return *this;
    }

    bool operator==(const T& a) const { return x == a && y == a; }
    bool operator==(const coord2d<T>& c) const { return x == c.x && y == c.y; }
    // the reference's scalar != reads `(x != a) || (y || a)` (src/coord2d.h:138-140); kept as is
    bool operator!=(const T& a) const { return (x != a) || (y || a); }
    bool operator!=(const coord2d<T>& c) const { return x != c.x || y != c.y; }

private:
    static void refuse_zero(bool is_zero) {
        if (is_zero) throw std::runtime_error("Divide by zero exception");
    }
};

typedef coord2d<unsigned int> dim;
typedef coord2d<of2d_real> vector2d;

#endif
