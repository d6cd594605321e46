// NormalEstimationPCLMeshModifier (normal_estimation_pcl.cpp:15-50): the per-point arithmetic of
// pcl::NormalEstimation behind a radius search, as __host__ __device__ functions (the kernel in kernels_cluster.cu runs them
// one thread per point; the reference runs them one point after the other on one core).  Written from PCL 1.12:
//
//   computeMeanAndCovarianceMatrix (common/impl/centroid.hpp)  shifted second moments of the neighbours: K = the first
//       neighbour of FLANN's distance-sorted list = the query point itself; accu / n; cov = E[xy] - E[x] E[y], in float
//   solvePlaneParameters / pcl::eigen33 / computeRoots / computeRoots2 (features/impl/feature.hpp, common/impl/eigen.hpp)
//       the matrix scaled by its largest absolute entry, closed-form roots of the characteristic polynomial, the
//       eigenvector of the smallest root as the largest cross product of two rows of (A - lambda I); curvature = |lambda / trace|
//   flipNormalTowardsViewpoint (features/normal_3d.h)            (vp - p) . n < 0  ->  n = -n
//
// Where this differs from the reference, on purpose: the nine sums over the neighbours are formed in double from the same
// float differences (q - K) and rounded to float once — the reference adds them in float in the order of the sorted
// neighbour list, an order this path does not reproduce — and atan2f / cosf / sinf are CUDA's.  The neighbour SETS are
// FLANN's (float squared distance strictly below float(radius^2)).  Parity is therefore a tolerance on the angle between
// the normals (1e-5 rad, tests/test_normal_estimation.py), not bit equality.
#pragma once
#include "host_math.h"

#include <cfloat>
#include <cmath>

namespace nb2
{
namespace pclne
{
NB2_HD inline void computeRoots2(float b, float c, float roots[3])
{
  roots[0] = 0.f;
  float d = static_cast<float>(static_cast<double>(b * b) - 4.0 * static_cast<double>(c));
  if (d < 0.f)
    d = 0.f;
  const float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

NB2_HD inline void swapf(float& a, float& b)
{
  const float t = a;
  a = b;
  b = t;
}

/** roots of the characteristic polynomial of the symmetric matrix {m00 m01 m02; . m11 m12; . . m22}, ascending */
NB2_HD inline void computeRoots(float m00, float m01, float m02, float m11, float m12, float m22, float roots[3])
{
  const float c0 = m00 * m11 * m22 + 2.f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  const float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  const float c2 = m00 + m11 + m22;
  if (fabsf(c0) < FLT_EPSILON)
  {
    computeRoots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = static_cast<float>(1.0 / 3.0);
  const float s_sqrt3 = sqrtf(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f)
    a_over_3 = 0.f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f)
    q = 0.f;
  const float rho = sqrtf(-a_over_3);
  const float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
  const float cos_theta = cosf(theta);
  const float sin_theta = sinf(theta);
  roots[0] = c2_over_3 + 2.f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1])
    swapf(roots[0], roots[1]);
  if (roots[1] >= roots[2])
  {
    swapf(roots[1], roots[2]);
    if (roots[0] >= roots[1])
      swapf(roots[0], roots[1]);
  }
  if (roots[0] <= 0.f)
    computeRoots2(c2, c1, roots);
}

/** pcl::eigen33(mat, eigenvalue, eigenvector) for the covariance {c00 c01 c02; . c11 c12; . . c22}: the smallest eigenvalue
 *  and its unit eigenvector */
NB2_HD inline void eigen33(float c00, float c01, float c02, float c11, float c12, float c22, float& eigenvalue, float n[3])
{
  float scale = fmaxf(fmaxf(fmaxf(fabsf(c00), fabsf(c01)), fmaxf(fabsf(c02), fabsf(c11))), fmaxf(fabsf(c12), fabsf(c22)));
  if (scale <= FLT_MIN)
    scale = 1.0f;
  float s00 = c00 / scale, s01 = c01 / scale, s02 = c02 / scale, s11 = c11 / scale, s12 = c12 / scale, s22 = c22 / scale;
  float roots[3];
  computeRoots(s00, s01, s02, s11, s12, s22, roots);
  eigenvalue = roots[0] * scale;
  s00 -= roots[0];
  s11 -= roots[0];
  s22 -= roots[0];
  // rows r0 = {s00 s01 s02}, r1 = {s01 s11 s12}, r2 = {s02 s12 s22}
  const float v1[3] = { s01 * s12 - s02 * s11, s02 * s01 - s00 * s12, s00 * s11 - s01 * s01 };  // r0 x r1
  const float v2[3] = { s01 * s22 - s02 * s12, s02 * s02 - s00 * s22, s00 * s12 - s01 * s02 };  // r0 x r2
  const float v3[3] = { s11 * s22 - s12 * s12, s12 * s02 - s01 * s22, s01 * s12 - s11 * s02 };  // r1 x r2
  const float l1 = v1[0] * v1[0] + (v1[1] * v1[1] + v1[2] * v1[2]);
  const float l2 = v2[0] * v2[0] + (v2[1] * v2[1] + v2[2] * v2[2]);
  const float l3 = v3[0] * v3[0] + (v3[1] * v3[1] + v3[2] * v3[2]);
  const float* v;
  float len;
  if (l1 >= l2 && l1 >= l3)
    v = v1, len = l1;
  else if (l2 >= l1 && l2 >= l3)
    v = v2, len = l2;
  else
    v = v3, len = l3;
  const float d = sqrtf(len);
  n[0] = v[0] / d;
  n[1] = v[1] / d;
  n[2] = v[2] / d;
}

/** From the sums over a point's k neighbours (S[0..5] = xx xy xz yy yz zz, S[6..8] = x y z of the float differences q - p) to
 *  the normal, flipped towards the viewpoint, and the curvature */
NB2_HD inline void normalFromSums(const double S[9], unsigned k, const float p[3], const float vp[3], float n[3], float& curvature)
{
  const float cnt = static_cast<float>(k);
  float a[9];
  for (int i = 0; i < 9; ++i)
    a[i] = static_cast<float>(S[i]) / cnt;
  const float c00 = a[0] - a[6] * a[6], c01 = a[1] - a[6] * a[7], c02 = a[2] - a[6] * a[8], c11 = a[3] - a[7] * a[7],
              c12 = a[4] - a[7] * a[8], c22 = a[5] - a[8] * a[8];
  float ev;
  eigen33(c00, c01, c02, c11, c12, c22, ev, n);
  const float eig_sum = c00 + c11 + c22;
  curvature = eig_sum != 0.f ? fabsf(ev / eig_sum) : 0.f;
  const float wx = vp[0] - p[0], wy = vp[1] - p[1], wz = vp[2] - p[2];
  const float cos_theta = wx * n[0] + wy * n[1] + wz * n[2];
  if (cos_theta < 0.f)
  {
    n[0] *= -1.f;
    n[1] *= -1.f;
    n[2] *= -1.f;
  }
}
}  // namespace pclne
}  // namespace nb2
