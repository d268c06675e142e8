// ImathVec.h — minimal stand-in for Imath 3.1's Vec2/Vec3 (TEST INFRASTRUCTURE, not product code).
//
// The reference's libbsdl needs nothing from Imath but V2f / V3f / C3f
// (src/libbsdl/README.md).  Imath is not installed in this image, so oracle/build_ref.py
// compiles the reference's own libbsdl headers (in place, under /root/reference) against this
// shim to get a reference-built checker for the restated lobes (oracle/_ref/libref_bsdl.so).
// Semantics follow Imath 3.1: length() switches to a scaled computation below
// 2*FLT_MIN, normalized() returns the zero vector for a null vector.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <math.h>   // Imath pulls the C names (isfinite ...) into the global namespace; libbsdl relies on it
#include <string>

namespace Imath {

template<class T> struct Vec2 {
    T x, y;
    constexpr Vec2() : x(0), y(0) {}
    constexpr Vec2(T a) : x(a), y(a) {}
    constexpr Vec2(T a, T b) : x(a), y(b) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    constexpr T dot(const Vec2& v) const { return x * v.x + y * v.y; }
    constexpr Vec2 operator+(const Vec2& v) const { return Vec2(x + v.x, y + v.y); }
    constexpr Vec2 operator-(const Vec2& v) const { return Vec2(x - v.x, y - v.y); }
    constexpr Vec2 operator-() const { return Vec2(-x, -y); }
    constexpr Vec2 operator*(const Vec2& v) const { return Vec2(x * v.x, y * v.y); }
    constexpr Vec2 operator*(T a) const { return Vec2(x * a, y * a); }
    constexpr Vec2 operator/(T a) const { return Vec2(x / a, y / a); }
    Vec2& operator+=(const Vec2& v) { x += v.x; y += v.y; return *this; }
    Vec2& operator*=(T a) { x *= a; y *= a; return *this; }
    T length2() const { return dot(*this); }
    T length() const { return std::sqrt(length2()); }
};
template<class T> constexpr Vec2<T> operator*(T a, const Vec2<T>& v) { return Vec2<T>(a * v.x, a * v.y); }

template<class T> struct Vec3 {
    T x, y, z;
    constexpr Vec3() : x(0), y(0), z(0) {}
    constexpr explicit Vec3(T a) : x(a), y(a), z(a) {}
    constexpr Vec3(T a, T b, T c) : x(a), y(b), z(c) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    constexpr T dot(const Vec3& v) const { return x * v.x + y * v.y + z * v.z; }
    constexpr Vec3 cross(const Vec3& v) const
    {
        return Vec3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x);
    }
    constexpr Vec3 operator+(const Vec3& v) const { return Vec3(x + v.x, y + v.y, z + v.z); }
    constexpr Vec3 operator-(const Vec3& v) const { return Vec3(x - v.x, y - v.y, z - v.z); }
    constexpr Vec3 operator-() const { return Vec3(-x, -y, -z); }
    constexpr Vec3 operator*(const Vec3& v) const { return Vec3(x * v.x, y * v.y, z * v.z); }
    constexpr Vec3 operator*(T a) const { return Vec3(x * a, y * a, z * a); }
    constexpr Vec3 operator/(const Vec3& v) const { return Vec3(x / v.x, y / v.y, z / v.z); }
    constexpr Vec3 operator/(T a) const { return Vec3(x / a, y / a, z / a); }
    Vec3& operator+=(const Vec3& v) { x += v.x; y += v.y; z += v.z; return *this; }
    Vec3& operator-=(const Vec3& v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
    Vec3& operator*=(const Vec3& v) { x *= v.x; y *= v.y; z *= v.z; return *this; }
    Vec3& operator*=(T a) { x *= a; y *= a; z *= a; return *this; }
    Vec3& operator/=(T a) { x /= a; y /= a; z /= a; return *this; }
    constexpr bool operator==(const Vec3& v) const { return x == v.x && y == v.y && z == v.z; }
    constexpr bool operator!=(const Vec3& v) const { return !(*this == v); }
    T length2() const { return dot(*this); }
    T lengthTiny() const
    {
        T ax = std::fabs(x), ay = std::fabs(y), az = std::fabs(z);
        T m = ax;
        if (m < ay) m = ay;
        if (m < az) m = az;
        if (m == T(0)) return T(0);
        ax /= m; ay /= m; az /= m;
        return m * std::sqrt(ax * ax + ay * ay + az * az);
    }
    T length() const
    {
        T l2 = length2();
        if (l2 < T(2) * FLT_MIN) return lengthTiny();
        return std::sqrt(l2);
    }
    const Vec3& normalize()
    {
        T l = length();
        if (l != T(0)) { x /= l; y /= l; z /= l; }
        return *this;
    }
    Vec3 normalized() const
    {
        T l = length();
        if (l == T(0)) return Vec3(T(0));
        return Vec3(x / l, y / l, z / l);
    }
};
template<class T> constexpr Vec3<T> operator*(T a, const Vec3<T>& v) { return Vec3<T>(a * v.x, a * v.y, a * v.z); }

typedef Vec2<float> V2f;
typedef Vec3<float> V3f;

}  // namespace Imath
