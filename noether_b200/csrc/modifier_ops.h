// Tool-path modifiers on device-resident poses: the per-waypoint arithmetic, shared by the CUDA kernels
// (modifiers.cu) and by the host code that sizes and orders the launches (modifier_plan.cpp).
//
//   SnakeOrganizationModifier                 snake_organization_modifier.cpp:5-22
//   LinearApproachModifier                    linear_approach_modifier.cpp:33-55
//   LinearDepartureModifier                   linear_departure_modifier.cpp:5-27
//   OffsetModifier                            offset_modifier.cpp:8-22
//   FixedOrientationModifier                  fixed_orientation_modifier.cpp:5-33
//   UniformOrientationModifier                uniform_orientation_modifier.cpp:5-30
//   ConcatenateModifier                       concatenate_modifier.cpp:5-15
//   ToolDragOrientationToolPathModifier       tool_drag_orientation_modifier.cpp:11-35
//   BiasedToolDragOrientationToolPathModifier biased_tool_drag_orientation_modifier.cpp:12-35
//   DirectionOfTravelOrientationModifier      direction_of_travel_orientation_modifier.cpp:6-48
//
// A pose is Eigen::Isometry3d storage: 16 doubles, column-major, last row 0 0 0 1.  The Eigen 3.4 Transform
// arithmetic the reference's expressions reduce to:
//   T * Translation3d(v)  = translate(v):  t += L v
//   T.rotate(R)           :                L  = L R
//   T * T2 (Isometry)     :                L' = L L2,  t' = L t2 + t,  last row 0 0 0 1
// with 3-term sums associated a0 + (a1 + a2) (host_math.h).  Nothing here may be contracted into FMAs: the library is
// built with -fmad=false / -ffp-contract=off.
#pragma once
#include "host_math.h"

#include <cstdint>

namespace nb2
{
namespace mod
{
/** Where the waypoints of one output segment come from (restructuring modifiers) */
struct SegMap
{
  uint32_t out_begin;  // first output waypoint of the segment
  uint32_t src_begin;  // first source waypoint
  uint32_t src_len;
  uint32_t n_pre;      // approach points put in front of the copied waypoints
  uint32_t n_post;     // departure points appended
  uint32_t flipped;    // the copied waypoints are taken in reverse order
};

NB2_HD inline D3 column(const double* T, int c) { return { T[4 * c], T[4 * c + 1], T[4 * c + 2] }; }
NB2_HD inline void setColumn(double* T, int c, const D3& v)
{
  T[4 * c] = v.x;
  T[4 * c + 1] = v.y;
  T[4 * c + 2] = v.z;
}

/** L v */
NB2_HD inline D3 linearTimes(const double* T, const D3& v)
{
  return { sum3(T[0] * v.x, T[4] * v.y, T[8] * v.z), sum3(T[1] * v.x, T[5] * v.y, T[9] * v.z),
           sum3(T[2] * v.x, T[6] * v.y, T[10] * v.z) };
}

/** L = L R (R row-major 3 x 3) */
NB2_HD inline void linearTimesMatrix(double* T, const double* R)
{
  double out[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      out[r][c] = sum3(T[r] * R[c], T[4 + r] * R[3 + c], T[8 + r] * R[6 + c]);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      T[4 * c + r] = out[r][c];
}

/** t += L v */
NB2_HD inline void translate(double* T, const D3& v)
{
  const D3 d = linearTimes(T, v);
  T[12] += d.x;
  T[13] += d.y;
  T[14] += d.z;
}

/** OffsetModifier: w = w * offset (offset column-major 4 x 4) */
NB2_HD inline void opOffset(double* T, const double* O)
{
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      R[3 * r + c] = O[4 * c + r];
  const D3 d = linearTimes(T, D3{ O[12], O[13], O[14] });
  linearTimesMatrix(T, R);
  T[12] = d.x + T[12];
  T[13] = d.y + T[13];
  T[14] = d.z + T[14];
  T[3] = T[7] = T[11] = 0.0;
  T[15] = 1.0;
}

/** FixedOrientationModifier: y = z x ref, x = y x z (columns are not re-normalised) */
NB2_HD inline void opFixedOrientation(double* T, const D3& ref)
{
  const D3 z = column(T, 2);
  const D3 y = cross(z, ref);
  const D3 x = cross(y, z);
  setColumn(T, 0, x);
  setColumn(T, 1, y);
}

/** UniformOrientationModifier: half a turn about z when x opposes the reference direction */
NB2_HD inline void opUniformOrientation(double* T, const D3& ref, const double* R_pi_z)
{
  if (dot(ref, column(T, 0)) < 0.0)
    linearTimesMatrix(T, R_pi_z);
}

/** ToolDragOrientationToolPathModifier: translate(0, 0, z_offset) then rotate(R) */
NB2_HD inline void opToolDrag(double* T, const double* R, double z_offset)
{
  translate(T, D3{ 0.0, 0.0, z_offset });
  linearTimesMatrix(T, R);
}

/** BiasedToolDragOrientationToolPathModifier: rotate(R) then translate(sign * tool_radius, 0, 0) */
NB2_HD inline void opBiasedToolDrag(double* T, const double* R, double x_offset)
{
  linearTimesMatrix(T, R);
  translate(T, D3{ x_offset, 0.0, 0.0 });
}

/** DirectionOfTravelOrientationModifier: keep z, y = z^ x dir^, x = y x z^ (dir = forward difference, backward at the end) */
NB2_HD inline void opDirectionOfTravel(double* T, const D3& dir)
{
  const D3 ref_x = normalized(dir);
  const D3 z = normalized(column(T, 2));
  const D3 y = normalized(cross(z, ref_x));
  const D3 x = normalized(cross(y, z));
  setColumn(T, 0, x);
  setColumn(T, 1, y);
}

/** createTransform (plane_slicer_raster_planner.cpp:380-394) from the float cut point, its interpolated normal and the
 *  step to the neighbouring point; with dot_modifier the DirectionOfTravelOrientationModifier of RasterPlanner::plan on
 *  top (what nb2_result_poses yields for a result whose raster post-ops were applied) */
NB2_HD inline void createTransform(const D3& p, const D3& n, const D3& dir, bool dot_modifier, double* T)
{
  D3 y = cross(n, dir);
  D3 x = normalized(cross(y, n));
  y = normalized(y);
  const D3 z = normalized(n);
  if (dot_modifier)
  {
    const D3 ref_x = normalized(dir);
    const D3 zz = normalized(z);
    y = normalized(cross(zz, ref_x));
    x = normalized(cross(y, zz));
  }
  T[0] = x.x, T[1] = x.y, T[2] = x.z, T[3] = 0.0;
  T[4] = y.x, T[5] = y.y, T[6] = y.z, T[7] = 0.0;
  T[8] = z.x, T[9] = z.y, T[10] = z.z, T[11] = 0.0;
  T[12] = p.x, T[13] = p.y, T[14] = p.z, T[15] = 1.0;
}

/** Sign of a tool path for the tool-drag modifiers: is the first waypoint's y axis aligned with z x (path direction)? */
NB2_HD inline double toolDragSign(const double* first, const D3& path_dir)
{
  return dot(cross(column(first, 2), path_dir), column(first, 1)) > 0.0 ? 1.0 : -1.0;
}

/** Index of the segment holding output waypoint w: the last s with seg_off[s] <= w (empty segments are skipped) */
NB2_HD inline uint32_t segmentOf(const uint32_t* seg_off, uint32_t n_segments, uint32_t w)
{
  uint32_t lo = 0, hi = n_segments;  // invariant: seg_off[lo] <= w < seg_off[hi]
  while (hi - lo > 1)
  {
    const uint32_t mid = (lo + hi) >> 1;
    if (seg_off[mid] <= w)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

/** Waypoint j of an output segment of a restructuring modifier (snake / approach / departure).
 *  Returns the source waypoint to copy; for approach / departure points also the step index i (the point is the anchor
 *  translated by offset * (i + 1) / n_points) through *lead, otherwise *lead = -1. */
NB2_HD inline uint32_t restructuredSource(const SegMap& m, uint32_t j, int* lead)
{
  *lead = -1;
  if (j < m.n_pre)
  {
    *lead = static_cast<int>(m.n_pre - 1 - j);  // inserted in reverse: the farthest point comes first
    return m.flipped ? m.src_begin + m.src_len - 1 : m.src_begin;
  }
  j -= m.n_pre;
  if (j < m.src_len)
    return m.flipped ? m.src_begin + m.src_len - 1 - j : m.src_begin + j;
  *lead = static_cast<int>(j - m.src_len);
  return m.flipped ? m.src_begin : m.src_begin + m.src_len - 1;
}

/** CircularLeadIn / CircularLeadOut point: (anchor * Translation(0, 0, r)) * AngleAxis(angle_i, Y) * Translation(0, 0, -r)
 *  with the anchor's orientation put back; T holds the anchor on entry, R = the rotation of step i (row-major) */
NB2_HD inline void circularLeadPoint(double* T, const double* R, double radius)
{
  double L[9];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      L[3 * c + r] = T[4 * c + r];
  translate(T, D3{ 0.0, 0.0, radius });
  linearTimesMatrix(T, R);
  translate(T, D3{ 0.0, 0.0, -radius });
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      T[4 * c + r] = L[3 * c + r];
}

NB2_HD inline void leadPoint(double* T, const double* offset, int i, int n_points)
{
  const double f = static_cast<double>(i + 1) / static_cast<double>(n_points);
  translate(T, D3{ offset[0] * f, offset[1] * f, offset[2] * f });
}
}  // namespace mod
}  // namespace nb2
