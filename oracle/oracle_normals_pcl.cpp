// TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product path.
//
// Sequential restatement of NormalEstimationPCLMeshModifier::modify (noether_tpp/src/mesh_modifiers/normal_estimation_pcl.cpp:15-50)
// and of the third-party calls it makes.  PCL is not under /root/reference (rosdep installs it: 1.12.1 on Ubuntu jammy, the
// reference's docker/Dockerfile:1-2 and CI); what follows is written from PCL 1.12's published sources:
//
//   pcl::NormalEstimation<PointXYZ, Normal>::computeFeature (features/normal_3d.hpp): for every point, in index order,
//     searchForNeighbors(idx, radius) == 0 or !computePointNormal(...)  ->  normal = curvature = NaN
//     else flipNormalTowardsViewpoint(point, vpx, vpy, vpz, nx, ny, nz)
//   pcl::search::KdTree / pcl::KdTreeFLANN::radiusSearch (sorted results): FLANN L2_Simple distance in float,
//     ((dx^2) + dy^2) + dz^2; a neighbour when strictly smaller than float(radius * radius); the results sorted by
//     (distance, index) — FLANN's DistanceIndex::operator<; the query point is its own first neighbour
//   computePointNormal (features/normal_3d.h): fewer than 3 neighbours -> false; computeMeanAndCovarianceMatrix in float:
//     K = the first finite point of the neighbour list; nine float accumulators of the shifted coordinates, summed in the
//     order of the list; accu /= n; covariance = E[xy] - E[x]E[y] (common/impl/centroid.hpp)
//   solvePlaneParameters (features/impl/feature.hpp): pcl::eigen33 — the matrix scaled by its largest absolute entry,
//     computeRoots (closed-form roots of the characteristic polynomial: float atan2 / cos / sin of libm), the eigenvector of
//     the smallest root as the largest of the three cross products of the rows of (A - lambda I), normalised;
//     curvature = |lambda_0 / trace| (0 when the trace is 0)
//   flipNormalTowardsViewpoint (features/normal_3d.h): (vp - p) . n < 0 in float  ->  n = -n
//   pcl::toPCLPointCloud2 + pcl::concatenateFields: the pcl::Normal records (normal_x/y/z at 0/4/8, curvature at 16, 32 bytes)
//     in front of the input's fields — the layout oracle_pca.cpp's concatenation already restates for NormalsFromMeshFaces
//
// The neighbour search itself is a uniform grid (the oracle's own; it only has to find the same neighbour sets and hand them
// over in FLANN's order).
//
// PARITY STATUS: parity unpinned.  The reference's test (mesh_modifier_utest.cpp:15-19,130-183: radius 0.01 on the test
// mesh) checks only that the returned mesh has polygons, points and normal fields.  The GPU path is compared with this file
// within an angle tolerance (1e-5 rad, BASELINE.json north_star), not bit for bit: its sums are formed in double and its
// atan2 / cos / sin are CUDA's, not glibc's.
#include "oracle.h"
#include "oracle_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>

namespace
{
struct V3f
{
  float x, y, z;
};
inline V3f cross(const V3f& a, const V3f& b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
// Eigen's squaredNorm of a fixed-size 3-vector: a0 + (a1 + a2)?  For floats of Vector3f Eigen 3.4 vectorises nothing (size 3) and
// the unrolled redux splits [0,3) into [0,1) + [1,3)
inline float squaredNorm(const V3f& a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }

/** pcl::computeRoots2 */
inline void computeRoots2(float b, float c, float roots[3])
{
  roots[0] = 0.f;
  float d = static_cast<float>(b * b - 4.0 * c);  // (the literal is a double in PCL's source)
  if (d < 0.0)
    d = 0.0;
  const float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

/** pcl::computeRoots (common/impl/eigen.hpp) for a symmetric 3x3 float matrix m (row-major) */
inline void computeRoots(const float m[9], float roots[3])
{
  const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
  const float c0 = m00 * m11 * m22 + 2.f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  const float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  const float c2 = m00 + m11 + m22;
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon())
  {
    computeRoots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = static_cast<float>(1.0 / 3.0);
  const float s_sqrt3 = std::sqrt(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f)
    a_over_3 = 0.f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f)
    q = 0.f;
  const float rho = std::sqrt(-a_over_3);
  const float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  const float cos_theta = std::cos(theta);
  const float sin_theta = std::sin(theta);
  roots[0] = c2_over_3 + 2.f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1])
    std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2])
  {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1])
      std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0)
    computeRoots2(c2, c1, roots);
}

/** pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenvalue and its eigenvector */
inline void eigen33(const float mat[9], float& eigenvalue, V3f& eigenvector)
{
  float scale = 0.f;
  for (int i = 0; i < 9; ++i)
    scale = std::max(scale, std::fabs(mat[i]));
  if (scale <= std::numeric_limits<float>::min())
    scale = 1.0f;
  float s[9];
  for (int i = 0; i < 9; ++i)
    s[i] = mat[i] / scale;
  float roots[3];
  computeRoots(s, roots);
  eigenvalue = roots[0] * scale;
  s[0] -= roots[0];
  s[4] -= roots[0];
  s[8] -= roots[0];
  const V3f r0{ s[0], s[1], s[2] }, r1{ s[3], s[4], s[5] }, r2{ s[6], s[7], s[8] };
  const V3f vec1 = cross(r0, r1), vec2 = cross(r0, r2), vec3 = cross(r1, r2);
  const float len1 = squaredNorm(vec1), len2 = squaredNorm(vec2), len3 = squaredNorm(vec3);
  V3f v;
  float len;
  if (len1 >= len2 && len1 >= len3)
    v = vec1, len = len1;
  else if (len2 >= len1 && len2 >= len3)
    v = vec2, len = len2;
  else
    v = vec3, len = len3;
  const float d = std::sqrt(len);
  eigenvector = { v.x / d, v.y / d, v.z / d };
}

struct CellKey
{
  long long x, y, z;
  bool operator==(const CellKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct CellHash
{
  size_t operator()(const CellKey& k) const
  {
    return static_cast<size_t>(k.x) * 0x9e3779b97f4a7c15ull ^ (static_cast<size_t>(k.y) * 0xc2b2ae3d27d4eb4full + 0x165667b19e3779f9ull) ^
           (static_cast<size_t>(k.z) << 32 | static_cast<size_t>(k.z) >> 32) * 0xff51afd7ed558ccdull;
  }
};
}  // namespace

extern "C" int orc_normal_estimation_pcl(const float* xyz, uint64_t n_points, double radius, double vx, double vy, double vz,
                                         float* normals, float* curvature, uint32_t* neighbour_count)
{
  {
    if (!xyz || !normals)
    {
      orc::setError("NULL argument");
      return ORC_ERR_INVALID;
    }
    const float nan = std::numeric_limits<float>::quiet_NaN();
    const float r2 = static_cast<float>(radius * radius);  // pcl::KdTreeFLANN::radiusSearch hands radius * radius to FLANN
    const float vpx = static_cast<float>(vx), vpy = static_cast<float>(vy), vpz = static_cast<float>(vz);  // setViewPoint(float, ...)
    // uniform grid of cells a little wider than the radius
    std::unordered_map<CellKey, std::vector<uint32_t>, CellHash> grid;
    const double h = radius > 0.0 ? radius * (1.0 + 1.0 / 1024.0) : 1.0;
    auto cellOf = [&](const float* p) {
      return CellKey{ static_cast<long long>(std::floor(p[0] / h)), static_cast<long long>(std::floor(p[1] / h)),
                      static_cast<long long>(std::floor(p[2] / h)) };
    };
    if (radius > 0.0)
      for (uint64_t i = 0; i < n_points; ++i)
        grid[cellOf(xyz + 3 * i)].push_back(static_cast<uint32_t>(i));
    struct Hit
    {
      float d;
      uint32_t i;
    };
    std::vector<Hit> nn;
    for (uint64_t i = 0; i < n_points; ++i)
    {
      const float* p = xyz + 3 * i;
      nn.clear();
      if (radius > 0.0)
      {
        const CellKey c = cellOf(p);
        for (long long dz = -1; dz <= 1; ++dz)
          for (long long dy = -1; dy <= 1; ++dy)
            for (long long dx = -1; dx <= 1; ++dx)
            {
              auto it = grid.find(CellKey{ c.x + dx, c.y + dy, c.z + dz });
              if (it == grid.end())
                continue;
              for (uint32_t q : it->second)
              {
                const float* b = xyz + 3ull * q;
                float result = 0.f;
                for (int a = 0; a < 3; ++a)
                {
                  const float diff = p[a] - b[a];
                  result += diff * diff;
                }
                if (result < r2)
                  nn.push_back({ result, q });
              }
            }
        std::sort(nn.begin(), nn.end(), [](const Hit& a, const Hit& b) { return a.d < b.d || (a.d == b.d && a.i < b.i); });
      }
      if (neighbour_count)
        neighbour_count[i] = static_cast<uint32_t>(nn.size());
      float* n = normals + 3 * i;
      if (nn.size() < 3)
      {
        n[0] = n[1] = n[2] = nan;
        if (curvature)
          curvature[i] = nan;
        continue;
      }
      // computeMeanAndCovarianceMatrix, dense cloud
      const float* K = xyz + 3ull * nn[0].i;
      float accu[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
      for (const Hit& hit : nn)
      {
        const float* q = xyz + 3ull * hit.i;
        const float x = q[0] - K[0], y = q[1] - K[1], z = q[2] - K[2];
        accu[0] += x * x;
        accu[1] += x * y;
        accu[2] += x * z;
        accu[3] += y * y;
        accu[4] += y * z;
        accu[5] += z * z;
        accu[6] += x;
        accu[7] += y;
        accu[8] += z;
      }
      const float cnt = static_cast<float>(nn.size());
      for (float& a : accu)
        a /= cnt;
      float cov[9];
      cov[0] = accu[0] - accu[6] * accu[6];
      cov[1] = accu[1] - accu[6] * accu[7];
      cov[2] = accu[2] - accu[6] * accu[8];
      cov[4] = accu[3] - accu[7] * accu[7];
      cov[5] = accu[4] - accu[7] * accu[8];
      cov[8] = accu[5] - accu[8] * accu[8];
      cov[3] = cov[1];
      cov[6] = cov[2];
      cov[7] = cov[5];
      float ev;
      V3f nv;
      eigen33(cov, ev, nv);
      const float eig_sum = cov[0] + cov[4] + cov[8];
      const float curv = eig_sum != 0 ? std::fabs(ev / eig_sum) : 0.f;
      // flipNormalTowardsViewpoint
      const float wx = vpx - p[0], wy = vpy - p[1], wz = vpz - p[2];
      const float cos_theta = wx * nv.x + wy * nv.y + wz * nv.z;
      if (cos_theta < 0)
      {
        nv.x *= -1;
        nv.y *= -1;
        nv.z *= -1;
      }
      n[0] = nv.x, n[1] = nv.y, n[2] = nv.z;
      if (curvature)
        curvature[i] = curv;
    }
  }
  return 0;
}
