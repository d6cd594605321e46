// TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product path.
//
// Small fixed-size vector helpers that restate the floating-point evaluation ORDER of the third-party
// expressions the reference goes through (Eigen 3.4 fixed-size kernels, VTK 9.1 C loops).  Everything here must be
// compiled with -ffp-contract=off and without -ffast-math / -march=native so no FMA is formed: the reference's
// distro builds (x86-64 baseline, SSE2) cannot contract mul+add either.
//
// Eigen 3.4 reductions over a fixed-size 3-vector (dot, squaredNorm, sum) are unrolled by
// internal::redux_novec_unroller, which splits [0,3) as [0,1) + [1,3): the result is  a0 + (a1 + a2).
// Plain VTK C code (vtkPlane::EvaluateFunction, vtkTriangle::Contour) evaluates left to right: (a0 + a1) + a2.
#pragma once
#include <cmath>
#include <cstdint>

namespace orc
{
template <typename T>
struct Vec3
{
  T x, y, z;
  T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
using V3d = Vec3<double>;
using V3f = Vec3<float>;

template <typename T>
inline Vec3<T> operator+(const Vec3<T>& a, const Vec3<T>& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
template <typename T>
inline Vec3<T> operator-(const Vec3<T>& a, const Vec3<T>& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
template <typename T>
inline Vec3<T> operator*(T s, const Vec3<T>& a) { return { s * a.x, s * a.y, s * a.z }; }
template <typename T>
inline Vec3<T> operator*(const Vec3<T>& a, T s) { return { a.x * s, a.y * s, a.z * s }; }
template <typename T>
inline Vec3<T> operator/(const Vec3<T>& a, T s) { return { a.x / s, a.y / s, a.z / s }; }

/** Eigen fixed-size 3-vector sum order: a0 + (a1 + a2) */
template <typename T>
inline T eigenSum3(T a0, T a1, T a2) { return a0 + (a1 + a2); }
/** Eigen `a.dot(b)` for Vector3 */
template <typename T>
inline T eigenDot(const Vec3<T>& a, const Vec3<T>& b) { return eigenSum3(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T>
inline T eigenSquaredNorm(const Vec3<T>& a) { return eigenSum3(a.x * a.x, a.y * a.y, a.z * a.z); }
template <typename T>
inline T eigenNorm(const Vec3<T>& a) { return std::sqrt(eigenSquaredNorm(a)); }
/** Eigen 3.4 `normalized()`: divides by sqrt(squaredNorm) when squaredNorm > 0, else returns the vector unchanged */
template <typename T>
inline Vec3<T> eigenNormalized(const Vec3<T>& a)
{
  const T z = eigenSquaredNorm(a);
  if (z > T(0))
    return a / std::sqrt(z);
  return a;
}
/** Eigen `a.cross(b)` */
template <typename T>
inline Vec3<T> eigenCross(const Vec3<T>& a, const Vec3<T>& b)
{
  return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
inline V3d toDouble(const V3f& a) { return { static_cast<double>(a.x), static_cast<double>(a.y), static_cast<double>(a.z) }; }

}  // namespace orc
