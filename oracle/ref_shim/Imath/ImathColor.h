// ImathColor.h — minimal stand-in for Imath 3.1's Color3 (TEST INFRASTRUCTURE, see ImathVec.h).
#pragma once
#include "ImathVec.h"

namespace Imath {

template<class T> struct Color3 : public Vec3<T> {
    constexpr Color3() : Vec3<T>() {}
    constexpr explicit Color3(T a) : Vec3<T>(a) {}
    constexpr Color3(T a, T b, T c) : Vec3<T>(a, b, c) {}
    constexpr Color3(const Vec3<T>& v) : Vec3<T>(v) {}
    constexpr Color3 operator+(const Color3& v) const { return Color3(this->x + v.x, this->y + v.y, this->z + v.z); }
    constexpr Color3 operator-(const Color3& v) const { return Color3(this->x - v.x, this->y - v.y, this->z - v.z); }
    constexpr Color3 operator-() const { return Color3(-this->x, -this->y, -this->z); }
    constexpr Color3 operator*(const Color3& v) const { return Color3(this->x * v.x, this->y * v.y, this->z * v.z); }
    constexpr Color3 operator*(T a) const { return Color3(this->x * a, this->y * a, this->z * a); }
    constexpr Color3 operator/(const Color3& v) const { return Color3(this->x / v.x, this->y / v.y, this->z / v.z); }
    constexpr Color3 operator/(T a) const { return Color3(this->x / a, this->y / a, this->z / a); }
};
template<class T> constexpr Color3<T> operator*(T a, const Color3<T>& v) { return Color3<T>(a * v.x, a * v.y, a * v.z); }

typedef Color3<float> C3f;

}  // namespace Imath
