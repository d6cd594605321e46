// Frame arithmetic: everything between the device moment reduction and the lattice constants the slicing kernels
// consume.  15 scalars in, a handful out, but every quirk matters for parity.  Written once, compiled for the host
// (explicit C-ABI entry points such as nb2_make_lattice) and for the device (single-thread kernels that keep the plan
// free of host round trips, frame_dev.cuh); IEEE arithmetic without contraction on both sides.
//
//   finalizeMoments         pcl::PCA::initCompute                      (3P; call site plane_slicer_raster_planner.cpp:540-541)
//   eigenSolver3f           Eigen::SelfAdjointEigenSolver<Matrix3f>    (3P, Eigen 3.4 iterative QL path)
//   principalAxisDirection  PrincipalAxisDirectionGenerator::generate  (principal_axis_direction_generator.cpp:14-26)
//   makeLattice             planImpl :550-586 + getDistancesToMinMaxCuts :460-501
#pragma once
#include "host_math.h"
#include "nb2_internal.h"

#include <cfloat>

namespace nb2
{
namespace frame
{
// Eigen 3.4 JacobiRotation<float>::makeGivens, real case
struct Givens
{
  float c, s;
};
NB2_HD inline Givens givens(float p, float q)
{
  Givens g;
  if (q == 0.f)
  {
    g.c = p < 0.f ? -1.f : 1.f;
    g.s = 0.f;
    return g;
  }
  if (p == 0.f)
  {
    g.c = 0.f;
    g.s = q < 0.f ? 1.f : -1.f;
    return g;
  }
  if (fabsf(p) > fabsf(q))
  {
    const float t = q / p;
    float u = sqrtf(1.f + t * t);
    if (p < 0.f)
      u = -u;
    g.c = 1.f / u;
    g.s = -t * g.c;
    return g;
  }
  const float t = p / q;
  float u = sqrtf(1.f + t * t);
  if (q < 0.f)
    u = -u;
  g.s = -1.f / u;
  g.c = -t * g.s;
  return g;
}

// Eigen 3.4 numext::hypot for real scalars (internal::positive_real_hypot): p * sqrt(1 + (q/p)^2), p = max(|x|,|y|)
NB2_HD inline float eigenHypot(float x, float y)
{
  x = fabsf(x);
  y = fabsf(y);
  const float p = fmaxf(x, y);
  if (p == 0.f)
    return 0.f;
  const float qp = fminf(y, x) / p;
  return p * sqrtf(1.f + qp * qp);
}

// one implicit symmetric QR step with Wilkinson shift on the unreduced block [start, end] of a 3x3 tridiagonal,
// accumulating the rotations into Q (column-major)
NB2_HD inline void qrStep(float d[3], float e[2], int start, int end, float Q[9])
{
  const float td = (d[end - 1] - d[end]) * 0.5f;
  const float ee = e[end - 1];
  float mu = d[end];
  if (td == 0.f)
    mu -= fabsf(ee);
  else if (ee != 0.f)
  {
    const float e2 = ee * ee;
    const float h = eigenHypot(td, ee);
    if (e2 == 0.f)
      mu -= ee / ((td + (td > 0.f ? h : -h)) / ee);
    else
      mu -= e2 / (td + (td > 0.f ? h : -h));
  }
  float x = d[start] - mu;
  float z = e[start];
  for (int k = start; k < end && z != 0.f; ++k)
  {
    const Givens r = givens(x, z);
    const float sdk = r.s * d[k] + r.c * e[k];
    const float dkp1 = r.s * e[k] + r.c * d[k + 1];
    d[k] = r.c * (r.c * d[k] - r.s * e[k]) - r.s * (r.c * e[k] - r.s * d[k + 1]);
    d[k + 1] = r.s * sdk + r.c * dkp1;
    e[k] = r.c * sdk - r.s * dkp1;
    if (k > start)
      e[k - 1] = r.c * e[k - 1] - r.s * z;
    x = e[k];
    if (k < end - 1)
    {
      z = -r.s * e[k + 1];
      e[k + 1] = r.c * e[k + 1];
    }
    float* qa = Q + 3 * k;
    float* qb = Q + 3 * (k + 1);
    for (int i = 0; i < 3; ++i)
    {
      const float a = qa[i], b = qb[i];
      qa[i] = r.c * a - r.s * b;
      qb[i] = r.s * a + r.c * b;
    }
  }
}
/** Symmetric 3x3 float eigen-decomposition with the algorithm, scaling, deflation test and ordering of Eigen 3.4's
 *  SelfAdjointEigenSolver<Matrix3f>::compute (lower triangle read; eigenvalues ascending; eigenvectors in columns).
 *  The sign of each eigenvector is whatever this sequence of rotations yields — the reference inherits the same. */
NB2_HD inline void eigenSolver3f(const float A[9], float evals[3], float evecs[9])
{
  // lower triangle, scaled into [-1, 1]
  float l00 = A[0], l10 = A[1], l20 = A[2], l11 = A[4], l21 = A[5], l22 = A[8];
  float scale = fmaxf(fmaxf(fmaxf(fabsf(l00), fabsf(l10)), fmaxf(fabsf(l20), fabsf(l11))), fmaxf(fabsf(l21), fabsf(l22)));
  if (scale == 0.f)
    scale = 1.f;
  l00 /= scale;
  l10 /= scale;
  l20 /= scale;
  l11 /= scale;
  l21 /= scale;
  l22 /= scale;

  // Householder tridiagonalisation, closed form for 3x3
  float d[3], e[2];
  float Q[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
  d[0] = l00;
  const float v1norm2 = l20 * l20;
  if (v1norm2 <= FLT_MIN)
  {
    d[1] = l11;
    d[2] = l22;
    e[0] = l10;
    e[1] = l21;
  }
  else
  {
    const float beta = sqrtf(l10 * l10 + v1norm2);
    const float invBeta = 1.f / beta;
    const float m01 = l10 * invBeta;
    const float m02 = l20 * invBeta;
    const float q = 2.f * m01 * l21 + m02 * (l22 - l11);
    d[1] = l11 + m02 * q;
    d[2] = l22 - m02 * q;
    e[0] = beta;
    e[1] = l21 - m01 * q;
    // Q = [1 0 0; 0 m01 m02; 0 m02 -m01], column-major
    Q[4] = m01;
    Q[5] = m02;
    Q[7] = m02;
    Q[8] = -m01;
  }

  // QR iterations with deflation
  int end = 2, start = 0, iter = 0;
  const float tiny = FLT_MIN;
  const float precision_inv = 1.f / FLT_EPSILON;
  while (end > 0)
  {
    for (int i = start; i < end; ++i)
    {
      if (fabsf(e[i]) < tiny)
        e[i] = 0.f;
      else
      {
        const float se = precision_inv * e[i];
        if (se * se <= (fabsf(d[i]) + fabsf(d[i + 1])))
          e[i] = 0.f;
      }
    }
    while (end > 0 && e[end - 1] == 0.f)
      --end;
    if (end <= 0)
      break;
    if (++iter > 30 * 3)
      break;
    start = end - 1;
    while (start > 0 && e[start - 1] != 0.f)
      --start;
    qrStep(d, e, start, end, Q);
  }

  // ascending selection sort, columns follow
  for (int i = 0; i < 2; ++i)
  {
    int k = i;
    for (int j = i + 1; j < 3; ++j)
      if (d[j] < d[k])
        k = j;
    if (k != i)
    {
      const float td = d[i];
      d[i] = d[k];
      d[k] = td;
      for (int r = 0; r < 3; ++r)
      {
        const float tq = Q[3 * i + r];
        Q[3 * i + r] = Q[3 * k + r];
        Q[3 * k + r] = tq;
      }
    }
  }
  for (int i = 0; i < 3; ++i)
    evals[i] = d[i] * scale;
  for (int i = 0; i < 9; ++i)
    evecs[i] = Q[i];
}

/** fp64 pivot-shifted raw moments -> mean, covariance (N - 1 normalisation as pcl::PCA), float casts, eigen axes in
 *  pcl::PCA's descending order.  `scales` is filled later from the extents pass. */
NB2_HD inline void finalizeMoments(const MomentPartial& t, uint64_t V, const float pivot[3], nb2_pca& out)
{
  const double n = static_cast<double>(V);
  const double m[3] = { t.s[0] / n, t.s[1] / n, t.s[2] / n };
  for (int c = 0; c < 3; ++c)
  {
    out.mean64[c] = static_cast<double>(pivot[c]) + m[c];
    out.mean[c] = static_cast<float>(out.mean64[c]);
    out.aabb_min[c] = t.mn[c];
    out.aabb_max[c] = t.mx[c];
  }
  const double den = n - 1.0;
  const double xx = (t.q[0] - n * m[0] * m[0]) / den, xy = (t.q[1] - n * m[0] * m[1]) / den,
               xz = (t.q[2] - n * m[0] * m[2]) / den, yy = (t.q[3] - n * m[1] * m[1]) / den,
               yz = (t.q[4] - n * m[1] * m[2]) / den, zz = (t.q[5] - n * m[2] * m[2]) / den;
  const double C[9] = { xx, xy, xz, xy, yy, yz, xz, yz, zz };
  for (int i = 0; i < 9; ++i)
  {
    out.cov64[i] = C[i];
    out.cov[i] = static_cast<float>(C[i]);
  }
  float ev[3], evec[9];
  eigenSolver3f(out.cov, ev, evec);
  for (int i = 0; i < 3; ++i)
  {
    out.evals[i] = ev[2 - i];
    for (int r = 0; r < 3; ++r)
      out.evecs[3 * i + r] = evec[3 * (2 - i) + r];
  }
}

/** sin_r / cos_r = sin / cos of rotation_offset, always evaluated by the host libm (the reference's platform) */
NB2_HD inline void principalAxisDirection(const nb2_pca& pca, double sin_r, double cos_r, double out[3])
{
  const D3 e1 = normalized(fromFloat(pca.evecs + 3));
  const D3 axis = normalized(fromFloat(pca.evecs + 6));
  // Eigen::AngleAxisd::toRotationMatrix
  const D3 sin_axis = sin_r * axis;
  const double c = cos_r;
  const D3 cos1_axis = (1.0 - c) * axis;
  double R[3][3];
  double tmp = cos1_axis.x * axis.y;
  R[0][1] = tmp - sin_axis.z;
  R[1][0] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z;
  R[0][2] = tmp + sin_axis.y;
  R[2][0] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z;
  R[1][2] = tmp - sin_axis.x;
  R[2][1] = tmp + sin_axis.x;
  R[0][0] = cos1_axis.x * axis.x + c;
  R[1][1] = cos1_axis.y * axis.y + c;
  R[2][2] = cos1_axis.z * axis.z + c;
  for (int i = 0; i < 3; ++i)
    out[i] = sum3(R[i][0] * e1.x, R[i][1] * e1.y, R[i][2] * e1.z);
}

/** PCARotatedDirectionGenerator::generate (pca_rotated_direction_generator.cpp:13-25): `nominal` (as is, not normalised)
 *  rotated about the smallest principal axis */
NB2_HD inline void pcaRotatedDirection(const nb2_pca& pca, double sin_r, double cos_r, const double nominal[3], double out[3])
{
  const D3 v = { nominal[0], nominal[1], nominal[2] };
  const D3 axis = normalized(fromFloat(pca.evecs + 6));
  const D3 sin_axis = sin_r * axis;
  const double c = cos_r;
  const D3 cos1_axis = (1.0 - c) * axis;
  double R[3][3];
  double tmp = cos1_axis.x * axis.y;
  R[0][1] = tmp - sin_axis.z;
  R[1][0] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z;
  R[0][2] = tmp + sin_axis.y;
  R[2][0] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z;
  R[1][2] = tmp - sin_axis.x;
  R[2][1] = tmp + sin_axis.x;
  R[0][0] = cos1_axis.x * axis.x + c;
  R[1][1] = cos1_axis.y * axis.y + c;
  R[2][2] = cos1_axis.z * axis.z + c;
  for (int i = 0; i < 3; ++i)
    out[i] = sum3(R[i][0] * v.x, R[i][1] * v.y, R[i][2] * v.z);
}

constexpr int kLatticeOk = 0, kLatticeBadSpacing = 1, kLatticeNoNormal = 2, kLatticeTooManyPlanes = 3;

NB2_HD inline bool finite64(double x) { return x - x == 0.0; }

/** Returns one of the kLattice* codes */
NB2_HD inline int makeLattice(const nb2_pca& pca, const double dir[3], const double origin[3], double spacing,
                              int bidirectional, nb2_lattice& L)
{
  if (!(spacing > 0.0) || !finite64(spacing))
    return kLatticeBadSpacing;
  const D3 mesh_normal = normalized(fromFloat(pca.evecs + 6));
  const D3 centroid = fromFloat(pca.mean);
  const D3 cut_direction = { dir[0], dir[1], dir[2] };
  const D3 cut_origin = { origin[0], origin[1], origin[2] };
  const D3 cut_normal = normalized(cross(normalized(cut_direction), mesh_normal));
  if (!(squaredNorm(cut_normal) > 0.0) || !finite64(squaredNorm(cut_normal)))
    return kLatticeNoNormal;

  // half extents of the OBB along each principal axis, scaled in float as the reference does before the cast
  double half[3][3];  // [row][axis]
  for (int a = 0; a < 3; ++a)
    for (int r = 0; r < 3; ++r)
      half[r][a] = static_cast<double>(pca.evecs[3 * a + r] * pca.scales[a]) / 2.0;

  double d_max = -DBL_MAX;
  double d_min = DBL_MAX;
  for (int corner = 0; corner < 8; ++corner)
  {
    const double sx = (corner & 1) ? -1.0 : 1.0, sy = (corner & 2) ? -1.0 : 1.0, sz = (corner & 4) ? -1.0 : 1.0;
    D3 p;
    p.x = sum3(half[0][0] * sx, half[0][1] * sy, half[0][2] * sz) + centroid.x;
    p.y = sum3(half[1][0] * sx, half[1][1] * sy, half[1][2] * sz) + centroid.y;
    p.z = sum3(half[2][0] * sx, half[2][1] * sy, half[2][2] * sz) + centroid.z;
    double d = dot(cut_normal, cut_origin) + (-dot(cut_normal, p));
    d *= -1;
    // the reference's update: a value that raises d_max never lowers d_min
    if (d > d_max)
      d_max = d;
    else if (d < d_min)
      d_min = d;
  }

  L.no_cuts = 0;
  if (!bidirectional)
  {
    if (d_max < 0.0)
      L.no_cuts = 1;
    if (d_min < 0.0)
      d_min = 0.0;
  }
  const double cut_span = fabs(d_max - d_min);
  const double planes_f = ceil(cut_span / spacing);
  if (!(planes_f < 2147483000.0))
    return kLatticeTooManyPlanes;
  const auto num_planes = static_cast<uint64_t>(planes_f);
  const auto start_steps = static_cast<uint64_t>(floor(fabs(d_min) / spacing));
  const double start_offset = start_steps * spacing;
  const D3 start_loc = cut_origin + cut_normal * copysign(start_offset, d_min);

  const D3* src[4] = { &cut_direction, &cut_normal, &cut_origin, &start_loc };
  double* dst[4] = { L.cut_direction, L.cut_normal, L.cut_origin, L.start_loc };
  for (int i = 0; i < 4; ++i)
  {
    dst[i][0] = src[i]->x;
    dst[i][1] = src[i]->y;
    dst[i][2] = src[i]->z;
  }
  L.line_spacing = spacing;
  L.d_min = d_min;
  L.d_max = d_max;
  L.num_planes = L.no_cuts ? 0 : num_planes;
  L.pad_ = 0;
  return kLatticeOk;
}

}  // namespace frame
}  // namespace nb2
