// MovingAverageOrientationSmoothingModifier (moving_average_orientation_smoothing_modifier.cpp:6-84) — the arithmetic
// of one smoothing step, shared by the CUDA kernel (modifiers.cu) and the host emulation in tests/cpp.
//
// The reference's step is: Eigen::Quaterniond(pose.linear()) for every pose of the window, M = sum q q^T (Markley's
// quaternion mean), Eigen::JacobiSVD<Matrix4d>(M, ComputeFullU), mean = column 0 of U, pose.linear() =
// mean.toRotationMatrix().  Restated from Eigen 3.4: Geometry/Quaternion.h (quaternionbase_assign_impl<Other,3,3>,
// toRotationMatrix), SVD/JacobiSVD.h (compute, real_2x2_jacobi_svd), Jacobi/Jacobi.h (makeJacobi, operator*, transpose,
// apply_rotation_in_the_plane).  3-term sums associate a0 + (a1 + a2) as everywhere in this library (host_math.h).
#pragma once
#include "host_math.h"

#include <cfloat>

namespace nb2
{
namespace mod
{
/** Quaterniond(Matrix3d) — coefficients in Eigen's storage order x y z w; T is a column-major 4 x 4 pose */
NB2_HD inline void rotationToQuaternion(const double* T, double q[4])
{
  auto m = [&](int r, int c) { return T[4 * c + r]; };
  double t = sum3(m(0, 0), m(1, 1), m(2, 2));
  if (t > 0.0)
  {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  }
  else
  {
    int i = 0;
    if (m(1, 1) > m(0, 0))
      i = 1;
    if (m(2, 2) > m(i, i))
      i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
}

/** q.toRotationMatrix() written into the linear part of the column-major pose T (the quaternion is used as given) */
NB2_HD inline void quaternionToRotation(const double q[4], double* T)
{
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  T[0] = 1.0 - (tyy + tzz);
  T[4] = txy - twz;
  T[8] = txz + twy;
  T[1] = txy + twz;
  T[5] = 1.0 - (txx + tzz);
  T[9] = tyz - twx;
  T[2] = txz - twy;
  T[6] = tyz + twx;
  T[10] = 1.0 - (txx + tyy);
}

struct Jacobi2
{
  double c, s;
};

/** JacobiRotation::makeJacobi(x, y, z) for the real symmetric 2 x 2 matrix [x y; y z] */
NB2_HD inline Jacobi2 makeJacobi(double x, double y, double z)
{
  const double deno = 2.0 * fabs(y);
  if (deno < DBL_MIN)
    return { 1.0, 0.0 };
  const double tau = (x - z) / deno;
  const double w = sqrt(tau * tau + 1.0);
  const double t = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
  const double sign_t = t > 0.0 ? 1.0 : -1.0;
  const double n = 1.0 / sqrt(t * t + 1.0);
  return { n, -sign_t * (y / fabs(y)) * fabs(t) * n };
}

/** apply_rotation_in_the_plane on two strided vectors of four entries: x' = c x + s y, y' = -s x + c y */
NB2_HD inline void rotatePlane(double* x, int xs, double* y, int ys, const Jacobi2& j)
{
  if (j.c == 1.0 && j.s == 0.0)
    return;
#pragma unroll
  for (int i = 0; i < 4; ++i)
  {
    const double xi = x[i * xs], yi = y[i * ys];
    x[i * xs] = j.c * xi + j.s * yi;
    y[i * ys] = -j.s * xi + j.c * yi;
  }
}

/** Column 0 of U of Eigen::JacobiSVD<Matrix4d>(M, ComputeFullU).  M is row-major 4 x 4 and is destroyed. */
NB2_HD inline void jacobiSvdFirstLeftVector(double M[16], double u0[4])
{
  double U[16];  // row-major
  for (int i = 0; i < 16; ++i)
    U[i] = (i % 5 == 0) ? 1.0 : 0.0;
  double scale = 0.0;
  bool finite = true;
  for (int i = 0; i < 16; ++i)
  {
    finite = finite && (M[i] - M[i] == 0.0);
    scale = fmax(scale, fabs(M[i]));
  }
  if (!finite)
  {
    // Eigen stops with InvalidInput and leaves U unset; the reference's NaN check is then undefined: refused here
    u0[0] = u0[1] = u0[2] = u0[3] = M[0] - M[0] + (scale - scale) + nan("");
    return;
  }
  if (scale == 0.0)
    scale = 1.0;
  for (int i = 0; i < 16; ++i)
    M[i] = M[i] / scale;
  const double precision = 2.0 * DBL_EPSILON;
  const double considerAsZero = DBL_MIN;
  double maxDiag = 0.0;
  for (int i = 0; i < 4; ++i)
    maxDiag = fmax(maxDiag, fabs(M[5 * i]));
  bool finished = false;
  int sweeps = 0;
  while (!finished && sweeps++ < 64)  // (Eigen has no sweep limit; 64 is never reached on finite input)
  {
    finished = true;
    // (both loops unrolled: every index below is then a constant and M, U stay in registers on the device)
#pragma unroll
    for (int p = 1; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < p; ++q)
      {
        const double threshold = fmax(considerAsZero, precision * maxDiag);
        if (fabs(M[4 * p + q]) > threshold || fabs(M[4 * q + p]) > threshold)
        {
          finished = false;
          // real_2x2_jacobi_svd on [M(p,p) M(p,q); M(q,p) M(q,q)]
          double m00 = M[4 * p + p], m01 = M[4 * p + q], m10 = M[4 * q + p], m11 = M[4 * q + q];
          Jacobi2 rot1;
          const double t = m00 + m11;
          const double d = m10 - m01;
          if (fabs(d) < DBL_MIN)
            rot1 = { 1.0, 0.0 };
          else
          {
            const double u = t / d;
            const double tmp = sqrt(1.0 + u * u);
            rot1 = { u / tmp, 1.0 / tmp };
          }
          // m.applyOnTheLeft(0, 1, rot1)
          if (!(rot1.c == 1.0 && rot1.s == 0.0))
          {
            const double a0 = m00, a1 = m01, b0 = m10, b1 = m11;
            m00 = rot1.c * a0 + rot1.s * b0;
            m01 = rot1.c * a1 + rot1.s * b1;
            m10 = -rot1.s * a0 + rot1.c * b0;
            m11 = -rot1.s * a1 + rot1.c * b1;
          }
          const Jacobi2 j_right = makeJacobi(m00, m01, m11);
          // j_left = rot1 * j_right.transpose():  (c1 c2' - s1 s2', c1 s2' + s1 c2') with transpose = (c, -s)
          const Jacobi2 jt = { j_right.c, -j_right.s };
          const Jacobi2 j_left = { rot1.c * jt.c - rot1.s * jt.s, rot1.c * jt.s + rot1.s * jt.c };
          // m_workMatrix.applyOnTheLeft(p, q, j_left): rows p and q
          rotatePlane(M + 4 * p, 1, M + 4 * q, 1, j_left);
          // m_matrixU.applyOnTheRight(p, q, j_left.transpose()): columns p and q rotated by the transpose of the argument
          rotatePlane(U + p, 4, U + q, 4, j_left);
          // m_workMatrix.applyOnTheRight(p, q, j_right): columns p and q rotated by j_right.transpose()
          rotatePlane(M + p, 4, M + q, 4, jt);
          maxDiag = fmax(maxDiag, fmax(fabs(M[5 * p]), fabs(M[5 * q])));
        }
      }
  }
  // singular values positive, then sorted in descending order (columns of U follow)
  double sv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
  {
    const double a = M[5 * i];
    sv[i] = fabs(a);
    if (a < 0.0)
    {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        U[4 * r + i] = -U[4 * r + i];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    sv[i] *= scale;
  // Only column 0 of the sorted U is wanted: it is the column of the first largest singular value (the selection sort
  // moves exactly that column to the front in its first step, or stops at once when every singular value is zero).
  int pos = 0;
  double best = sv[0];
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (sv[k] > best)
    {
      best = sv[k];
      pos = k;
    }
#pragma unroll
  for (int r = 0; r < 4; ++r)
    u0[r] = pos == 0 ? U[4 * r] : (pos == 1 ? U[4 * r + 1] : (pos == 2 ? U[4 * r + 2] : U[4 * r + 3]));
}

/** computeQuaternionMean over `count` quaternions read through `get(k)` (k = 0 .. count - 1, window order).
 *  Returns false when the mean holds a NaN (the reference throws "Orientation contains NaN values"). */
template <typename Get>
NB2_HD inline bool quaternionMean(int count, Get get, double mean[4])
{
  double M[16];
  for (int i = 0; i < 16; ++i)
    M[i] = 0.0;
  for (int k = 0; k < count; ++k)
  {
    double q[4];
    get(k, q);
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c)
        M[4 * r + c] += q[r] * q[c];
  }
  jacobiSvdFirstLeftVector(M, mean);
  return !(mean[0] != mean[0] || mean[1] != mean[1] || mean[2] != mean[2] || mean[3] != mean[3]);
}
}  // namespace mod
}  // namespace nb2
