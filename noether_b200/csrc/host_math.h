// 3-vector arithmetic of libnoether_b200.so, shared by the host code and the single-thread frame kernels.
//
// Results that feed discrete decisions (plane count, lattice origin, segment flips, sort keys) are evaluated with
// the operation order the reference's Eigen 3.4 expressions produce, so that the product and an x86-64 build of the
// reference take the same branches: fixed-size 3-term reductions (dot, squaredNorm) associate as a0 + (a1 + a2)
// (Eigen's redux_novec_unroller splits [0,3) into [0,1) + [1,3)).  Host and device code are both compiled without FMA
// contraction (-ffp-contract=off / -fmad=false) and use IEEE division and square root, so they agree bit for bit.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define NB2_HD __host__ __device__
#else
#define NB2_HD
#endif

namespace nb2
{
struct D3
{
  double x, y, z;
};
NB2_HD inline D3 operator+(const D3& a, const D3& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
NB2_HD inline D3 operator-(const D3& a, const D3& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
NB2_HD inline D3 operator*(double s, const D3& a) { return { s * a.x, s * a.y, s * a.z }; }
NB2_HD inline D3 operator*(const D3& a, double s) { return { a.x * s, a.y * s, a.z * s }; }
NB2_HD inline D3 operator/(const D3& a, double s) { return { a.x / s, a.y / s, a.z / s }; }
NB2_HD inline double sum3(double a, double b, double c) { return a + (b + c); }
NB2_HD inline double dot(const D3& a, const D3& b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
NB2_HD inline double squaredNorm(const D3& a) { return sum3(a.x * a.x, a.y * a.y, a.z * a.z); }
NB2_HD inline double norm(const D3& a) { return sqrt(squaredNorm(a)); }
/** Eigen 3.4 normalized(): v / sqrt(|v|^2) when |v|^2 > 0, otherwise v unchanged */
NB2_HD inline D3 normalized(const D3& a)
{
  const double z = squaredNorm(a);
  return z > 0.0 ? a / sqrt(z) : a;
}
NB2_HD inline D3 cross(const D3& a, const D3& b)
{
  return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
NB2_HD inline D3 fromFloat(const float* f) { return { static_cast<double>(f[0]), static_cast<double>(f[1]), static_cast<double>(f[2]) }; }
}  // namespace nb2
