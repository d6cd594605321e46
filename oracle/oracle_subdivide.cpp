// TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product path.
//
// Sequential restatement of the reference's face subdivision mesh modifiers:
//
//   FaceMidpointSubdivisionMeshModifier::modify   face_subdivision_modifier.cpp:202-248
//     subdivideTriangleRecursive                  :157-193   (depth first; midpoints of a triangle before its children)
//     getOrCreateMidpoint                         :126-151   (std::map keyed on std::minmax(start, end): one midpoint per
//                                                             edge, created by whichever triangle asks first)
//   FaceSubdivisionMeshModifier::modify           :250-312   (std::stack: last face first, children popped in the
//                                                             reverse of their push order)
//   FaceSubdivisionByAreaMeshModifier::requiresSubdivision   :319-339
//   copyPointToEnd                                :58-72     (the new record is a byte copy of the FIRST source ...)
//   updatePointData                               :78-114    (... whose x y z / normal / rgba become the mean of the sources)
//   getPoint / getNormal / getRgba                utils.cpp:74-133
//
// Eigen 3.4 arithmetic the expressions reduce to (x86-64 baseline build, no FMA contraction):
//   mean = Zero; mean += p_k (in source order); mean / sources.size()  -> float division by 2.0f / 3.0f
//     (the size_t divisor is promoted to the expression's scalar: internal::promote_scalar_arg)
//   Vector4i mean / size_t -> int division; .cast<uint8_t>()
//   0.5f * (v1 - v0).cross(v2 - v0).norm(), norm = sqrt(x^2 + (y^2 + z^2))   (redux_novec_unroller)
//
// PARITY STATUS: parity unpinned.  The reference's own test (mesh_modifier_utest.cpp:20-23,154-183) only checks that the
// modifiers run and keep the cloud's fields; the reference cannot be compiled here.
#include "oracle.h"
#include "oracle_internal.h"
#include "oracle_math.h"

#include <array>
#include <algorithm>
#include <cstring>
#include <map>
#include <stack>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace
{
using orc::V3f;

struct SubMesh
{
  std::vector<uint8_t> cloud;
  uint64_t width = 0;
  uint32_t step = 0, off_xyz = 0;
  int32_t off_nrm = -1, off_rgba = -1;
  std::vector<std::array<uint32_t, 3>> faces;
  uint64_t max_points = 0;
};

struct TooLarge
{
};

V3f loadV3(const SubMesh& m, uint64_t idx, uint32_t off)
{
  V3f v;
  std::memcpy(&v, m.cloud.data() + idx * m.step + off, 12);
  return v;
}
void storeV3(SubMesh& m, uint64_t idx, uint32_t off, const V3f& v) { std::memcpy(m.cloud.data() + idx * m.step + off, &v, 12); }

/** copyPointToEnd (:58-72) */
uint32_t copyPointToEnd(SubMesh& m, uint32_t copy_idx)
{
  if (m.max_points && m.width >= m.max_points)
    throw TooLarge{};
  std::vector<uint8_t> rec(m.cloud.begin() + static_cast<size_t>(copy_idx) * m.step,
                           m.cloud.begin() + static_cast<size_t>(copy_idx + 1) * m.step);
  m.cloud.insert(m.cloud.end(), rec.begin(), rec.end());
  m.width += 1;
  return static_cast<uint32_t>(m.width - 1);
}

/** updatePointData (:78-114) */
void updatePointData(SubMesh& m, uint32_t target, const uint32_t* sources, int n)
{
  const float count = static_cast<float>(n);
  {
    V3f mean{ 0.f, 0.f, 0.f };
    for (int k = 0; k < n; ++k)
      mean = mean + loadV3(m, sources[k], m.off_xyz);
    storeV3(m, target, m.off_xyz, mean / count);
  }
  if (m.off_nrm >= 0)
  {
    V3f mean{ 0.f, 0.f, 0.f };
    for (int k = 0; k < n; ++k)
      mean = mean + loadV3(m, sources[k], static_cast<uint32_t>(m.off_nrm));
    storeV3(m, target, static_cast<uint32_t>(m.off_nrm), mean / count);
  }
  if (m.off_rgba >= 0)
  {
    int mean[4] = { 0, 0, 0, 0 };
    for (int k = 0; k < n; ++k)
      for (int ch = 0; ch < 4; ++ch)
        mean[ch] += m.cloud[static_cast<size_t>(sources[k]) * m.step + m.off_rgba + ch];
    for (int ch = 0; ch < 4; ++ch)
      m.cloud[static_cast<size_t>(target) * m.step + m.off_rgba + ch] = static_cast<uint8_t>(mean[ch] / n);
  }
}

using EdgeMap = std::map<std::pair<uint32_t, uint32_t>, uint32_t>;

/** getOrCreateMidpoint (:126-151) */
uint32_t getOrCreateMidpoint(uint32_t idx_start, uint32_t idx_end, SubMesh& m, EdgeMap& midpoints)
{
  const std::pair<uint32_t, uint32_t> key = std::minmax(idx_start, idx_end);
  auto it = midpoints.find(key);
  if (it != midpoints.end())
    return it->second;
  const uint32_t mid = copyPointToEnd(m, idx_start);
  const uint32_t src[2] = { idx_start, idx_end };
  updatePointData(m, mid, src, 2);
  midpoints[key] = mid;
  return mid;
}

/** subdivideTriangleRecursive (:157-193) */
void subdivideTriangleRecursive(const std::array<uint32_t, 3>& tri, SubMesh& m, EdgeMap& midpoints, unsigned recursions)
{
  const uint32_t m0 = getOrCreateMidpoint(tri[0], tri[1], m, midpoints);
  const uint32_t m1 = getOrCreateMidpoint(tri[1], tri[2], m, midpoints);
  const uint32_t m2 = getOrCreateMidpoint(tri[2], tri[0], m, midpoints);
  const std::array<uint32_t, 3> sub[4] = { { tri[0], m0, m2 }, { m0, tri[1], m1 }, { m2, m1, tri[2] }, { m0, m1, m2 } };
  if (recursions == 1)
    for (const auto& s : sub)
      m.faces.push_back(s);
  else
    for (const auto& s : sub)
      subdivideTriangleRecursive(s, m, midpoints, recursions - 1);
}

/** FaceSubdivisionByAreaMeshModifier::requiresSubdivision (:319-339) for a triangle */
bool requiresSubdivision(const SubMesh& m, const std::array<uint32_t, 3>& f, float max_area)
{
  const V3f v0 = loadV3(m, f[0], m.off_xyz), v1 = loadV3(m, f[1], m.off_xyz), v2 = loadV3(m, f[2], m.off_xyz);
  const float tri_area = 0.5f * orc::eigenNorm(orc::eigenCross(v1 - v0, v2 - v0));
  if (!std::isfinite(tri_area))
    throw std::runtime_error("Triangle area is not finite");
  float area = 0.0f;
  area += tri_area;
  return area > max_area;
}

SubMesh makeInput(const uint8_t* cloud, uint64_t n_points, uint32_t step, uint32_t off_xyz, int32_t off_nrm, int32_t off_rgba)
{
  SubMesh m;
  m.cloud.assign(cloud, cloud + n_points * step);
  m.width = n_points;
  m.step = step;
  m.off_xyz = off_xyz;
  m.off_nrm = off_nrm;
  m.off_rgba = off_rgba;
  return m;
}
}  // namespace

extern "C" {

int orc_face_midpoint_subdivision(const uint8_t* cloud, uint64_t n_points, uint32_t point_step, uint32_t off_xyz, int32_t off_normal,
                                  int32_t off_rgba, const uint32_t* triangles, uint64_t n_triangles, uint32_t n_iterations,
                                  void** out)
{
  if (!cloud || !out || (n_triangles && !triangles) || n_iterations == 0)
  {
    // n_iterations == 0 never terminates in the reference (the recursion counts down from 0u, :175-191)
    orc::setError("invalid argument (n_iterations must be at least 1)");
    return ORC_ERR_INVALID;
  }
  for (uint64_t i = 0; i < 3 * n_triangles; ++i)
    if (triangles[i] >= n_points)
    {
      orc::setError("triangle index out of range");
      return ORC_ERR_INVALID;
    }
  auto* m = new SubMesh(makeInput(cloud, n_points, point_step, off_xyz, off_normal, off_rgba));
  EdgeMap midpoints;
  for (uint64_t t = 0; t < n_triangles; ++t)
    subdivideTriangleRecursive({ triangles[3 * t], triangles[3 * t + 1], triangles[3 * t + 2] }, *m, midpoints, n_iterations);
  *out = m;
  return ORC_OK;
}

int orc_face_subdivision_by_area(const uint8_t* cloud, uint64_t n_points, uint32_t point_step, uint32_t off_xyz, int32_t off_normal,
                                 int32_t off_rgba, const uint32_t* triangles, uint64_t n_triangles, float max_area,
                                 uint64_t max_points, void** out)
{
  if (!cloud || !out || (n_triangles && !triangles))
  {
    orc::setError("invalid argument");
    return ORC_ERR_INVALID;
  }
  for (uint64_t i = 0; i < 3 * n_triangles; ++i)
    if (triangles[i] >= n_points)
    {
      orc::setError("triangle index out of range");
      return ORC_ERR_INVALID;
    }
  auto* m = new SubMesh(makeInput(cloud, n_points, point_step, off_xyz, off_normal, off_rgba));
  m->max_points = max_points;
  std::stack<std::array<uint32_t, 3>> faces;
  for (uint64_t t = 0; t < n_triangles; ++t)
    faces.push({ triangles[3 * t], triangles[3 * t + 1], triangles[3 * t + 2] });
  try
  {
    while (!faces.empty())
    {
      const std::array<uint32_t, 3> f = faces.top();
      faces.pop();
      if (!requiresSubdivision(*m, f, max_area))
      {
        m->faces.push_back(f);
        continue;
      }
      const uint32_t mid = copyPointToEnd(*m, f[0]);
      updatePointData(*m, mid, f.data(), 3);
      faces.push({ f[0], f[1], mid });
      faces.push({ mid, f[1], f[2] });
      faces.push({ f[0], mid, f[2] });
    }
  }
  catch (const TooLarge&)
  {
    delete m;
    orc::setError("subdivision exceeds the point limit given by the caller");
    return ORC_ERR_INVALID;
  }
  catch (const std::exception& e)
  {
    delete m;
    orc::setError(e.what());
    return ORC_ERR_NON_FINITE;
  }
  *out = m;
  return ORC_OK;
}

uint64_t orc_submesh_num_points(const void* h) { return static_cast<const SubMesh*>(h)->width; }
uint64_t orc_submesh_num_triangles(const void* h) { return static_cast<const SubMesh*>(h)->faces.size(); }
void orc_submesh_export(const void* h, uint8_t* cloud, uint32_t* triangles)
{
  const auto* m = static_cast<const SubMesh*>(h);
  if (cloud)
    std::memcpy(cloud, m->cloud.data(), m->cloud.size());
  if (triangles)
    for (size_t t = 0; t < m->faces.size(); ++t)
      for (int k = 0; k < 3; ++k)
        triangles[3 * t + k] = m->faces[t][k];
}
void orc_submesh_free(void* h) { delete static_cast<SubMesh*>(h); }
}
