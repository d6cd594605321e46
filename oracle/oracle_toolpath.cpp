// TEST INFRASTRUCTURE ONLY — see oracle.h.
//
// Restatement of the tool-path post-ops on the hot path:
//   RasterOrganizationModifier::modify            raster_organization_modifier.cpp:7-46
//   estimateToolPathDirection / estimateRasterDirection   utils.cpp:7-47
//   DirectionOfTravelOrientationModifier::modify  direction_of_travel_orientation_modifier.cpp:6-48
//   JoinCloseSegmentsToolPathModifier::modify     join_close_segments_modifier.cpp:11-39
//   MinimumSegmentLengthToolPathModifier::modify  minimum_segment_length_modifier.cpp:12-37, utils.cpp:213-234
//   UniformSpacingModifier (degree 1, endpoints)  uniform_spacing_modifier.cpp:18-70, splines.cpp:38-58,149-229
// and of two reference test fixtures (tool_path_planner_utest.cpp:32-46, utils.cpp:249-291).
#include "oracle_internal.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <random>

namespace orc
{
namespace
{
/** estimateToolPathDirection (utils.cpp:7-29).  Returns false when it would throw. */
bool estimateToolPathDirection(const Path& tool_path, V3d& out)
{
  V3d avg_dir = { 0, 0, 0 };
  int n = 0;
  for (const Segment& seg : tool_path.segments)
  {
    if (seg.size() > 1)
    {
      for (size_t i = 0; i + 1 < seg.size(); ++i)
      {
        avg_dir = avg_dir + eigenNormalized(translation(seg[i + 1]) - translation(seg[i]));
        ++n;
      }
    }
  }
  if (n < 1)
  {
    setError("Insufficient number of points to calculate average tool path direction");
    return false;
  }
  avg_dir = avg_dir / static_cast<double>(n);
  out = eigenNormalized(avg_dir);
  return true;
}

/** estimateRasterDirection (utils.cpp:31-47) */
V3d estimateRasterDirection(const std::vector<Path>& tool_paths, const V3d& reference_tool_path_dir)
{
  const V3d first_wp = translation(tool_paths.front().segments.at(0).at(0));
  const V3d first_wp_last_path = translation(tool_paths.back().segments.at(0).at(0));
  const V3d norm_ref = eigenNormalized(reference_tool_path_dir);
  V3d raster_dir = first_wp_last_path - first_wp;
  raster_dir = raster_dir - eigenDot(raster_dir, norm_ref) * norm_ref;
  return eigenNormalized(raster_dir);
}
}  // namespace

int rasterOrganization(orc_paths& p)
{
  auto& tool_paths = p.paths;
  if (tool_paths.empty())
  {
    setError("vector::_M_range_check (tool_paths.at(0))");
    return ORC_ERR_INVALID;
  }
  V3d ref;
  if (!estimateToolPathDirection(tool_paths.at(0), ref))
    return ORC_ERR_TOO_FEW_POINTS;

  for (Path& tool_path : tool_paths)
  {
    for (Segment& segment : tool_path.segments)
    {
      const V3d diff = translation(segment.back()) - translation(segment.front());
      if (eigenDot(diff, ref) < 0.0)
        std::reverse(segment.begin(), segment.end());
    }
    // The reference's comparator re-reads tool_path.at(0).at(0) while sorting; that term is a common offset of both
    // sides so the order is the one of first_waypoint . ref.  std::sort is used as in the reference.
    const V3d base = translation(tool_path.segments.at(0).at(0));
    std::sort(tool_path.segments.begin(), tool_path.segments.end(), [&](const Segment& a, const Segment& b) {
      const V3d da = translation(a.at(0)) - base;
      const V3d db = translation(b.at(0)) - base;
      return eigenDot(da, ref) < eigenDot(db, ref);
    });
  }

  const V3d raster_dir = estimateRasterDirection(tool_paths, ref);
  const V3d first_wp = translation(tool_paths.at(0).segments.at(0).at(0));
  std::sort(tool_paths.begin(), tool_paths.end(), [&](const Path& a, const Path& b) {
    const V3d da = translation(a.segments.at(0).at(0)) - first_wp;
    const V3d db = translation(b.segments.at(0).at(0)) - first_wp;
    return eigenDot(da, raster_dir) < eigenDot(db, raster_dir);
  });
  return ORC_OK;
}

int directionOfTravel(orc_paths& p)
{
  for (Path& tool_path : p.paths)
  {
    for (Segment& segment : tool_path.segments)
    {
      if (segment.size() < 2)
      {
        setError("segment with a single waypoint (undefined behaviour in the reference)");
        return ORC_ERR_TOO_FEW_POINTS;
      }
      for (size_t i = 0; i < segment.size(); ++i)
      {
        V3d ref_x;
        if (i < segment.size() - 1)
          ref_x = eigenNormalized(translation(segment[i + 1]) - translation(segment[i]));
        else
          ref_x = eigenNormalized(translation(segment[i]) - translation(segment[i - 1]));
        Waypoint& w = segment[i];
        const V3d z_axis = eigenNormalized(V3d{ w.T[8], w.T[9], w.T[10] });
        const V3d y_axis = eigenNormalized(eigenCross(z_axis, ref_x));
        const V3d x_axis = eigenNormalized(eigenCross(y_axis, z_axis));
        w.T[0] = x_axis.x;
        w.T[1] = x_axis.y;
        w.T[2] = x_axis.z;
        w.T[4] = y_axis.x;
        w.T[5] = y_axis.y;
        w.T[6] = y_axis.z;
      }
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Tool-path modifiers applied after a raster plan.  Matrix products follow this oracle's Eigen
// convention for fixed-size 3-term sums, a0 + (a1 + a2) (oracle_math.h); Transform arithmetic restated from Eigen 3.4
// Geometry/Transform.h: operator*(Translation) = translate(): t += L v;  operator*(RotationBase) = rotate(): L = L R;
// Transform * Transform (Isometry): L = La Lb, t = La tb + ta, last row 0 0 0 1.
// ---------------------------------------------------------------------------------------------------------------
namespace
{
inline double& at(Waypoint& w, int r, int c) { return w.T[4 * c + r]; }
inline double at(const Waypoint& w, int r, int c) { return w.T[4 * c + r]; }

/** L(w) * v */
V3d linearTimes(const Waypoint& w, const V3d& v)
{
  V3d r;
  for (int i = 0; i < 3; ++i)
    r[i] = eigenSum3(at(w, i, 0) * v.x, at(w, i, 1) * v.y, at(w, i, 2) * v.z);
  return r;
}
/** L(w) = L(w) * R   (R row-major 3x3) */
void linearTimesMatrix(Waypoint& w, const double R[9])
{
  double out[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      out[r][c] = eigenSum3(at(w, r, 0) * R[0 * 3 + c], at(w, r, 1) * R[1 * 3 + c], at(w, r, 2) * R[2 * 3 + c]);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      at(w, r, c) = out[r][c];
}
/** Eigen::AngleAxisd(angle, axis).toRotationMatrix(), row-major (Eigen 3.4 Geometry/AngleAxis.h) */
void angleAxisMatrix(double angle, const V3d& axis, double R[9])
{
  const double s = std::sin(angle), c = std::cos(angle);
  const V3d sin_axis = s * axis;
  const V3d cos1_axis = (1.0 - c) * axis;
  double tmp;
  tmp = cos1_axis.x * axis.y;
  R[0 * 3 + 1] = tmp - sin_axis.z;
  R[1 * 3 + 0] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z;
  R[0 * 3 + 2] = tmp + sin_axis.y;
  R[2 * 3 + 0] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z;
  R[1 * 3 + 2] = tmp - sin_axis.x;
  R[2 * 3 + 1] = tmp + sin_axis.x;
  R[0 * 3 + 0] = cos1_axis.x * axis.x + c;
  R[1 * 3 + 1] = cos1_axis.y * axis.y + c;
  R[2 * 3 + 2] = cos1_axis.z * axis.z + c;
}
/** w.translate(v) */
void translate(Waypoint& w, const V3d& v)
{
  const V3d d = linearTimes(w, v);
  w.T[12] += d.x;
  w.T[13] += d.y;
  w.T[14] += d.z;
}
}  // namespace

/** SnakeOrganizationModifier::modify (snake_organization_modifier.cpp:5-22) */
int snakeOrganization(orc_paths& p)
{
  for (size_t i = 1; i < p.paths.size(); i += 2)
  {
    for (Segment& segment : p.paths[i].segments)
      std::reverse(segment.begin(), segment.end());
    std::reverse(p.paths[i].segments.begin(), p.paths[i].segments.end());
  }
  return ORC_OK;
}

/** LinearApproachModifier::modify (linear_approach_modifier.cpp:33-55) / LinearDepartureModifier::modify
 *  (linear_departure_modifier.cpp:5-27) */
int linearApproachDeparture(orc_paths& p, const double offset[3], int n_points, bool departure)
{
  const V3d off = { offset[0], offset[1], offset[2] };
  for (Path& tool_path : p.paths)
    for (Segment& segment : tool_path.segments)
    {
      if (segment.empty())
      {
        setError("empty segment (front() / back() of an empty vector in the reference)");
        return ORC_ERR_INVALID;
      }
      const Waypoint anchor = departure ? segment.back() : segment.front();
      Segment added;
      for (int i = 0; i < n_points; ++i)
      {
        Waypoint pt = anchor;
        translate(pt, off * (static_cast<double>(i + 1) / static_cast<double>(n_points)));
        added.push_back(pt);  // pt.linear() = anchor.linear(): unchanged by translate
      }
      if (departure)
        segment.insert(segment.end(), added.begin(), added.end());
      else
        segment.insert(segment.begin(), added.rbegin(), added.rend());
    }
  return ORC_OK;
}

/** OffsetModifier::modify (offset_modifier.cpp:8-22): w = w * offset */
int offsetModifier(orc_paths& p, const double O[16])
{
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      R[r * 3 + c] = O[4 * c + r];
  const V3d to = { O[12], O[13], O[14] };
  for (Path& tool_path : p.paths)
    for (Segment& segment : tool_path.segments)
      for (Waypoint& w : segment)
      {
        const V3d d = linearTimes(w, to);  // lhs.linear() * rhs.translation() is evaluated before the sum
        linearTimesMatrix(w, R);
        w.T[12] = d.x + w.T[12];
        w.T[13] = d.y + w.T[13];
        w.T[14] = d.z + w.T[14];
        w.T[3] = w.T[7] = w.T[11] = 0.0;  // makeAffine()
        w.T[15] = 1.0;
      }
  return ORC_OK;
}

/** FixedOrientationModifier::modify (fixed_orientation_modifier.cpp:10-33); the constructor normalises the direction */
int fixedOrientation(orc_paths& p, const double ref_x[3])
{
  const V3d ref = eigenNormalized(V3d{ ref_x[0], ref_x[1], ref_x[2] });
  for (Path& tool_path : p.paths)
    for (Segment& segment : tool_path.segments)
      for (Waypoint& w : segment)
      {
        const V3d z = { w.T[8], w.T[9], w.T[10] };
        const V3d y = eigenCross(z, ref);
        const V3d x = eigenCross(y, z);
        w.T[0] = x.x, w.T[1] = x.y, w.T[2] = x.z;
        w.T[4] = y.x, w.T[5] = y.y, w.T[6] = y.z;
      }
  return ORC_OK;
}

/** UniformOrientationModifier::modify (uniform_orientation_modifier.cpp:5-30) */
int uniformOrientation(orc_paths& p)
{
  if (p.paths.empty() || p.paths[0].segments.empty() || p.paths[0].segments[0].size() < 2)
  {
    setError("vector::_M_range_check (tool_paths.at(0).at(0).at(1))");
    return ORC_ERR_INVALID;
  }
  const V3d ref = eigenNormalized(translation(p.paths[0].segments[0][1]) - translation(p.paths[0].segments[0][0]));
  double R[9];
  angleAxisMatrix(M_PI, V3d{ 0.0, 0.0, 1.0 }, R);
  for (Path& tool_path : p.paths)
    for (Segment& segment : tool_path.segments)
      for (Waypoint& w : segment)
      {
        const double dp = eigenDot(ref, V3d{ w.T[0], w.T[1], w.T[2] });
        if (dp < 0.0)
          linearTimesMatrix(w, R);
      }
  return ORC_OK;
}

/** ConcatenateModifier::modify (concatenate_modifier.cpp:5-15) */
int concatenate(orc_paths& p)
{
  Path combined;
  combined.plane = p.paths.empty() ? -1 : p.paths.front().plane;
  for (const Path& path : p.paths)
    for (const Segment& segment : path.segments)
      combined.segments.push_back(segment);
  p.paths.clear();
  p.paths.push_back(std::move(combined));
  return ORC_OK;
}

/** ToolDragOrientationToolPathModifier::modify (tool_drag_orientation_modifier.cpp:11-35) and
 *  BiasedToolDragOrientationToolPathModifier::modify (biased_tool_drag_orientation_modifier.cpp:12-35) */
int toolDrag(orc_paths& p, double angle_offset, double tool_radius, bool biased)
{
  for (Path& tool_path : p.paths)
  {
    V3d dir;
    if (!estimateToolPathDirection(tool_path, dir))
      return ORC_ERR_TOO_FEW_POINTS;
    const Waypoint& first = tool_path.segments.at(0).at(0);
    const V3d y = { first.T[4], first.T[5], first.T[6] };
    const V3d z = { first.T[8], first.T[9], first.T[10] };
    const double sign = eigenDot(eigenCross(z, dir), y) > 0.0 ? 1.0 : -1.0;
    double R[9];
    angleAxisMatrix(sign * -angle_offset, V3d{ 0.0, 1.0, 0.0 }, R);
    for (Segment& segment : tool_path.segments)
      for (Waypoint& w : segment)
      {
        if (!biased)
        {
          const double z_offset = tool_radius * std::sin(angle_offset);
          translate(w, V3d{ 0.0, 0.0, z_offset });
          linearTimesMatrix(w, R);
        }
        else
        {
          linearTimesMatrix(w, R);
          translate(w, V3d{ sign * tool_radius, 0.0, 0.0 });
        }
      }
  }
  return ORC_OK;
}

/** CircularLeadInModifier::modify (circular_lead_in_modifier.cpp:12-44) / CircularLeadOutModifier::modify
 *  (circular_lead_out_modifier.cpp:12-44).  pt = radius_center * AngleAxisd(angle, UnitY) * Translation3d(0, 0, -r) is
 *  Transform::rotate (L = L R) followed by Transform::translate (t += L v); the orientation is then reset to the anchor's. */
int circularLead(orc_paths& p, double arc_angle, double arc_radius, int n_points, bool lead_out)
{
  const double n_points_d = static_cast<double>(n_points);  // the member is a double
  const double delta_theta = arc_angle / (n_points_d - 1);
  for (Path& tool_path : p.paths)
  {
    V3d dir;
    if (!estimateToolPathDirection(tool_path, dir))
      return ORC_ERR_TOO_FEW_POINTS;
    if (tool_path.segments.at(0).empty())
    {
      setError("first segment of a tool path is empty (tool_path.at(0).at(0))");
      return ORC_ERR_INVALID;
    }
    const Waypoint& first = tool_path.segments.at(0).at(0);
    const V3d y = { first.T[4], first.T[5], first.T[6] };
    const V3d z = { first.T[8], first.T[9], first.T[10] };
    const double sign = eigenDot(eigenCross(z, dir), y) > 0.0 ? 1.0 : -1.0;
    for (Segment& segment : tool_path.segments)
    {
      if (segment.empty())
      {
        setError("empty segment (front() / back() of an empty vector in the reference)");
        return ORC_ERR_INVALID;
      }
      const Waypoint anchor = lead_out ? segment.back() : segment.front();
      Waypoint radius_center = anchor;
      translate(radius_center, V3d{ 0.0, 0.0, arc_radius });
      Segment added;
      for (int i = 0; i < n_points; ++i)
      {
        const double angle = lead_out ? sign * -delta_theta * (i + 1) : sign * delta_theta * (i + 1);
        double R[9];
        angleAxisMatrix(angle, V3d{ 0.0, 1.0, 0.0 }, R);
        Waypoint pt = radius_center;
        linearTimesMatrix(pt, R);
        translate(pt, V3d{ 0.0, 0.0, -arc_radius });
        for (int c = 0; c < 3; ++c)
          for (int r = 0; r < 3; ++r)
            at(pt, r, c) = at(anchor, r, c);
        added.push_back(pt);
      }
      if (lead_out)
        segment.insert(segment.end(), added.begin(), added.end());  // (inserted at the front, then appended reversed)
      else
        segment.insert(segment.begin(), added.rbegin(), added.rend());
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// MovingAverageOrientationSmoothingModifier (moving_average_orientation_smoothing_modifier.cpp:6-84).  Third-party
// pieces restated from Eigen 3.4: Quaternion(Matrix3) (Geometry/Quaternion.h quaternionbase_assign_impl<Other,3,3>),
// QuaternionBase::toRotationMatrix, JacobiSVD<Matrix4d>::compute with ComputeFullU (SVD/JacobiSVD.h: no
// preconditioner for a square matrix, scaling by the largest |entry|, sweeps over p > q with the 2 * epsilon * maxDiag
// threshold, real_2x2_jacobi_svd, sign fix, descending sort) and JacobiRotation (Jacobi/Jacobi.h).
// ---------------------------------------------------------------------------------------------------------------
namespace
{
struct Quat
{
  double x, y, z, w;
  double& coeff(int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
  double coeff(int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

Quat quaternionFromLinear(const Waypoint& p)
{
  Quat q{};
  double t = eigenSum3(at(p, 0, 0), at(p, 1, 1), at(p, 2, 2));  // mat.trace()
  if (t > 0.0)
  {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (at(p, 2, 1) - at(p, 1, 2)) * t;
    q.y = (at(p, 0, 2) - at(p, 2, 0)) * t;
    q.z = (at(p, 1, 0) - at(p, 0, 1)) * t;
    return q;
  }
  int i = 0;
  if (at(p, 1, 1) > at(p, 0, 0))
    i = 1;
  if (at(p, 2, 2) > at(p, i, i))
    i = 2;
  const int j = (i + 1) % 3;
  const int k = (j + 1) % 3;
  t = std::sqrt(at(p, i, i) - at(p, j, j) - at(p, k, k) + 1.0);
  q.coeff(i) = 0.5 * t;
  t = 0.5 / t;
  q.w = (at(p, k, j) - at(p, j, k)) * t;
  q.coeff(j) = (at(p, j, i) + at(p, i, j)) * t;
  q.coeff(k) = (at(p, k, i) + at(p, i, k)) * t;
  return q;
}

void setLinearFromQuaternion(Waypoint& p, const Quat& q)
{
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  at(p, 0, 0) = 1.0 - (tyy + tzz);
  at(p, 0, 1) = txy - twz;
  at(p, 0, 2) = txz + twy;
  at(p, 1, 0) = txy + twz;
  at(p, 1, 1) = 1.0 - (txx + tzz);
  at(p, 1, 2) = tyz - twx;
  at(p, 2, 0) = txz - twy;
  at(p, 2, 1) = tyz + twx;
  at(p, 2, 2) = 1.0 - (txx + tyy);
}

/** Eigen::JacobiRotation<double> */
struct Rotation
{
  double c = 1.0, s = 0.0;
  Rotation transpose() const { return { c, -s }; }
  Rotation operator*(const Rotation& o) const { return { c * o.c - s * o.s, c * o.s + s * o.c }; }
  /** makeJacobi(x, y, z) for [x y; y z] */
  void makeJacobi(double x, double y, double z)
  {
    const double deno = 2.0 * std::abs(y);
    if (deno < std::numeric_limits<double>::min())
    {
      c = 1.0;
      s = 0.0;
      return;
    }
    const double tau = (x - z) / deno;
    const double w = std::sqrt(tau * tau + 1.0);
    const double t = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
    const double sign_t = t > 0.0 ? 1.0 : -1.0;
    const double n = 1.0 / std::sqrt(t * t + 1.0);
    s = -sign_t * (y / std::abs(y)) * std::abs(t) * n;
    c = n;
  }
};

/** Square matrix of doubles with Eigen's in-plane rotations */
template <int N>
struct Square
{
  double v[N][N];
  double& operator()(int r, int c) { return v[r][c]; }
  double operator()(int r, int c) const { return v[r][c]; }
  /** applyOnTheLeft(p, q, j): rows p and q */
  void rotateRows(int p, int q, const Rotation& j)
  {
    if (j.c == 1.0 && j.s == 0.0)
      return;
    for (int i = 0; i < N; ++i)
    {
      const double xi = v[p][i], yi = v[q][i];
      v[p][i] = j.c * xi + j.s * yi;
      v[q][i] = -j.s * xi + j.c * yi;
    }
  }
  /** applyOnTheRight(p, q, j): columns p and q, rotated by j.transpose() */
  void rotateColumns(int p, int q, const Rotation& j)
  {
    const Rotation jt = j.transpose();
    if (jt.c == 1.0 && jt.s == 0.0)
      return;
    for (int i = 0; i < N; ++i)
    {
      const double xi = v[i][p], yi = v[i][q];
      v[i][p] = jt.c * xi + jt.s * yi;
      v[i][q] = -jt.s * xi + jt.c * yi;
    }
  }
};

/** JacobiSVD<Matrix4d>(M, ComputeFullU).matrixU().col(0); false when M is not finite (Eigen: InvalidInput) */
bool firstLeftSingularVector(Square<4> W, double u0[4])
{
  Square<4> U{};
  for (int i = 0; i < 4; ++i)
    U(i, i) = 1.0;
  double scale = 0.0;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c)
    {
      if (!std::isfinite(W(r, c)))
        return false;
      scale = std::max(scale, std::abs(W(r, c)));
    }
  if (scale == 0.0)
    scale = 1.0;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c)
      W(r, c) = W(r, c) / scale;
  const double precision = 2.0 * std::numeric_limits<double>::epsilon();
  const double consider_as_zero = std::numeric_limits<double>::min();
  double max_diag = 0.0;
  for (int i = 0; i < 4; ++i)
    max_diag = std::max(max_diag, std::abs(W(i, i)));
  for (bool finished = false; !finished;)
  {
    finished = true;
    for (int p = 1; p < 4; ++p)
      for (int q = 0; q < p; ++q)
      {
        const double threshold = std::max(consider_as_zero, precision * max_diag);
        if (!(std::abs(W(p, q)) > threshold || std::abs(W(q, p)) > threshold))
          continue;
        finished = false;
        // real_2x2_jacobi_svd
        Square<2> m{};
        m(0, 0) = W(p, p), m(0, 1) = W(p, q), m(1, 0) = W(q, p), m(1, 1) = W(q, q);
        Rotation rot1;
        const double t = m(0, 0) + m(1, 1);
        const double d = m(1, 0) - m(0, 1);
        if (std::abs(d) < std::numeric_limits<double>::min())
        {
          rot1.s = 0.0;
          rot1.c = 1.0;
        }
        else
        {
          const double u = t / d;
          const double tmp = std::sqrt(1.0 + u * u);
          rot1.s = 1.0 / tmp;
          rot1.c = u / tmp;
        }
        m.rotateRows(0, 1, rot1);
        Rotation j_right;
        j_right.makeJacobi(m(0, 0), m(0, 1), m(1, 1));
        const Rotation j_left = rot1 * j_right.transpose();
        W.rotateRows(p, q, j_left);
        U.rotateColumns(p, q, j_left.transpose());
        W.rotateColumns(p, q, j_right);
        max_diag = std::max(max_diag, std::max(std::abs(W(p, p)), std::abs(W(q, q))));
      }
  }
  double sv[4];
  for (int i = 0; i < 4; ++i)
  {
    const double a = W(i, i);
    sv[i] = std::abs(a);
    if (a < 0.0)
      for (int r = 0; r < 4; ++r)
        U(r, i) = -U(r, i);
  }
  for (double& x : sv)
    x *= scale;
  for (int i = 0; i < 4; ++i)
  {
    int pos = 0;
    double best = sv[i];
    for (int k = 1; k < 4 - i; ++k)
      if (sv[i + k] > best)
      {
        best = sv[i + k];
        pos = k;
      }
    if (best == 0.0)
      break;
    if (pos)
    {
      pos += i;
      std::swap(sv[i], sv[pos]);
      for (int r = 0; r < 4; ++r)
        std::swap(U(r, i), U(r, pos));
    }
  }
  for (int r = 0; r < 4; ++r)
    u0[r] = U(r, 0);
  return true;
}
}  // namespace

/** MovingAverageOrientationSmoothingModifier::modify (:70-84): every pose is replaced in place, so the window of a
 *  pose holds the already smoothed poses before it */
int movingAverageOrientationSmoothing(orc_paths& p, int window_size)
{
  for (Path& tp : p.paths)
    for (Segment& tps : tp.segments)
    {
      const int n = static_cast<int>(tps.size());
      for (int i = 0; i < n; ++i)
      {
        Quat out = quaternionFromLinear(tps[i]);
        if (i > 0 && i < n - 1)
        {
          const int n_each_side = window_size / 2;
          const int start_idx = std::max(i - n_each_side, 0);
          const int stop_idx = std::min(i + n_each_side, n - 1);
          Square<4> M{};
          for (int k = start_idx; k < stop_idx + 1; ++k)
          {
            const Quat q = quaternionFromLinear(tps[k]);
            for (int r = 0; r < 4; ++r)
              for (int c = 0; c < 4; ++c)
                M(r, c) += q.coeff(r) * q.coeff(c);
          }
          double u0[4];
          const bool ok = firstLeftSingularVector(M, u0);
          if (!ok || std::isnan(u0[0]) || std::isnan(u0[1]) || std::isnan(u0[2]) || std::isnan(u0[3]))
          {
            setError("Orientation contains NaN values");
            return ORC_ERR_INVALID;
          }
          out = { u0[0], u0[1], u0[2], u0[3] };
        }
        setLinearFromQuaternion(tps[i], out);
      }
    }
  return ORC_OK;
}

int joinCloseSegments(orc_paths& p, double distance)
{
  for (Path& tool_path : p.paths)
  {
    auto& segs = tool_path.segments;
    if (segs.size() < 2)
      continue;
    size_t seg = 0, next = 1;
    while (next < segs.size())
    {
      const double d = eigenNorm(translation(segs[seg].back()) - translation(segs[next].front()));
      if (d < distance)
      {
        segs[seg].insert(segs[seg].end(), segs[next].begin(), segs[next].end());
        segs.erase(segs.begin() + next);
      }
      else
      {
        ++seg;
        ++next;
      }
    }
  }
  return ORC_OK;
}

/** computeLength(segment) (utils.cpp:213-234) */
static double segmentLength(const Segment& s, std::vector<double>* dists = nullptr)
{
  double length = 0.0;
  if (dists)
  {
    dists->clear();
    dists->push_back(0.0);
  }
  for (size_t i = 0; i + 1 < s.size(); ++i)
  {
    length += eigenNorm(translation(s[i + 1]) - translation(s[i]));
    if (dists)
      dists->push_back(length);
  }
  return length;
}

int minimumSegmentLength(orc_paths& p, double min_length)
{
  auto& paths = p.paths;
  for (auto tp = paths.begin(); tp != paths.end();)
  {
    auto& segs = tp->segments;
    for (auto s = segs.begin(); s != segs.end();)
    {
      if (segmentLength(*s) < min_length)
        s = segs.erase(s);
      else
        ++s;
    }
    if (segs.empty())
      tp = paths.erase(tp);
    else
      ++tp;
  }
  return ORC_OK;
}

/** UniformSpacingModifier::resample for spline_degree 1, include_endpoints true.
 *
 *  A degree-1 chord-length interpolating B-spline (Eigen::SplineFitting::Interpolate, 3P) is the polyline itself:
 *  knots {0,0,u_1..u_{n-2},1,1} with u_i = cumulative chord length / total, control points = data points, so
 *  computeSpanLengths (Simpson 3/8 on a constant derivative, splines.cpp:15-36,60-77) returns the chord lengths and
 *  computeSplineParameterAtDistance (splines.cpp:149-176) converges after one step to the linear position inside the
 *  span.  This restatement evaluates that closed form; it is not a bit-level restatement of Eigen's basis-function
 *  evaluation (tolerance-level parity). */
static int resample(const Segment& segment, double spacing, Segment& output)
{
  const size_t n = segment.size();
  if (n < 3)  // segment.size() < spline_degree + 2 (uniform_spacing_modifier.cpp:20-21)
  {
    setError("Spline cannot be fit to tool path segment with less than degree+2 waypoints");
    return ORC_ERR_SEGMENT_TOO_SHORT;
  }
  std::vector<V3d> P(n);
  for (size_t i = 0; i < n; ++i)
    P[i] = translation(segment[i]);
  // span lengths / knot distances
  std::vector<double> kd(n);  // cumulative distance at data point i
  kd[0] = 0.0;
  for (size_t i = 1; i < n; ++i)
    kd[i] = kd[i - 1] + eigenNorm(P[i] - P[i - 1]);
  const double l = kd[n - 1];

  // computeEquidistantKnots (splines.cpp:190-229): distances at which to sample
  std::vector<double> dist;
  double s;
  const double remainder = std::fmod(l, spacing);
  if (remainder < (spacing / 2.0))
    s = spacing / 2.0 + remainder / 2.0;
  else
    s = remainder / 2.0;
  while (s < l)
  {
    dist.push_back(s);
    s += spacing;
  }

  struct Sample
  {
    V3d pos;
    V3d tangent;
  };
  auto spanTangent = [&](size_t i) { return P[i + 1] - P[i]; };  // derivative direction on span [i, i+1]
  std::vector<Sample> samples;
  samples.push_back({ P[0], spanTangent(0) });  // u = 0
  for (double d : dist)
  {
    // estimateSplineParameterAtDistance (splines.cpp:107-144): exact knot hit, else lower_bound span
    size_t hit = n;
    for (size_t i = 0; i < n; ++i)
    {
      if (std::abs(kd[i] - d) < std::numeric_limits<double>::epsilon())
      {
        hit = i;
        break;
      }
    }
    if (hit != n)
    {
      // derivative at a knot is taken from the span that starts there (last span at the final knot)
      const size_t span = std::min(hit, n - 2);
      samples.push_back({ P[hit], spanTangent(span) });
      continue;
    }
    const size_t idx_end = std::lower_bound(kd.begin(), kd.end(), d) - kd.begin();
    const size_t i0 = idx_end - 1;
    const double frac = (d - kd[i0]) / (kd[idx_end] - kd[i0]);
    samples.push_back({ P[i0] + frac * (P[idx_end] - P[i0]), spanTangent(i0) });
  }
  samples.push_back({ P[n - 1], spanTangent(n - 2) });  // u = 1

  // sample() (splines.cpp:38-58) with the propagated reference z (uniform_spacing_modifier.cpp:39-48)
  V3d ref_z = { segment.front().T[8], segment.front().T[9], segment.front().T[10] };
  output.clear();
  for (const Sample& sm : samples)
  {
    const V3d x = sm.tangent;
    const V3d y = eigenCross(ref_z, x);
    const V3d z = eigenCross(x, y);
    const V3d xn = eigenNormalized(x), yn = eigenNormalized(y), zn = eigenNormalized(z);
    Waypoint w{};
    const double M[16] = { xn.x, xn.y, xn.z, 0.0, yn.x, yn.y, yn.z, 0.0, zn.x, zn.y, zn.z, 0.0, sm.pos.x, sm.pos.y,
                           sm.pos.z, 1.0 };
    std::copy(M, M + 16, w.T);
    for (int c = 0; c < 3; ++c)
    {
      w.pos[c] = static_cast<float>(sm.pos[c]);
      w.nrm[c] = static_cast<float>(zn[c]);
    }
    w.edge_lo = w.edge_hi = 0xFFFFFFFFu;
    output.push_back(w);
    ref_z = zn;
  }
  return ORC_OK;
}

int uniformSpacing(orc_paths& p, double point_spacing)
{
  if (!(point_spacing > 0.0))
  {
    setError("point_spacing must be positive");
    return ORC_ERR_INVALID;
  }
  for (Path& tool_path : p.paths)
  {
    for (Segment& segment : tool_path.segments)
    {
      Segment out;
      const int rc = resample(segment, point_spacing, out);
      if (rc != ORC_OK)
        return rc;
      segment.swap(out);
    }
  }
  return ORC_OK;
}

}  // namespace orc

// =================================================================================================================
// C API
// =================================================================================================================
using namespace orc;

extern "C" {

const char* orc_last_error(void) { return g_last_error.c_str(); }

int orc_pca(const float* xyz, uint64_t n, int mode, orc_pca_result* out) { return pca(xyz, n, mode, out); }
int orc_centroid(const float* xyz, uint64_t n, int mode, double out[3]) { return centroid(xyz, n, mode, out); }
int orc_principal_axis_direction(const orc_pca_result* pca_, double rot, double out[3])
{
  return principalAxisDirection(pca_, rot, out);
}
int orc_pca_rotated_direction(const orc_pca_result* pca_, double angle, const double nominal[3], double out[3])
{
  return pcaRotatedDirection(pca_, angle, nominal, out);
}
int orc_eigen_solver_3f(const float cov[9], float evals[3], float evecs[9])
{
  eigenSolver3f(cov, evals, evecs);
  return ORC_OK;
}
int orc_make_lattice(const orc_pca_result* pca_,
                     const double dir[3],
                     const double origin[3],
                     double spacing,
                     int bidirectional,
                     orc_lattice* out)
{
  return makeLattice(pca_, dir, origin, spacing, bidirectional, out);
}
int orc_normals_from_mesh_faces(const float* xyz, uint64_t V, const uint32_t* tri, uint64_t F, float* out)
{
  return normalsFromMeshFaces(xyz, V, tri, F, out);
}

int orc_slice(const float* xyz,
              const float* normals,
              uint64_t V,
              const uint32_t* tri,
              uint64_t F,
              const orc_lattice* L,
              int slice_mode,
              int scan_mode,
              uint64_t plane_begin,
              uint64_t plane_end,
              orc_paths** out)
{
  auto* p = new orc_paths();
  const int rc = slice(xyz, normals, V, tri, F, *L, slice_mode, scan_mode, plane_begin, plane_end, *p);
  if (rc != ORC_OK)
  {
    delete p;
    *out = nullptr;
    return rc;
  }
  *out = p;
  return ORC_OK;
}

orc_paths* orc_paths_create(void) { return new orc_paths(); }
void orc_paths_free(orc_paths* p) { delete p; }
orc_paths* orc_paths_clone(const orc_paths* p) { return new orc_paths(*p); }
void orc_paths_add_path(orc_paths* p) { p->paths.emplace_back(); }
void orc_paths_add_segment(orc_paths* p) { p->paths.back().segments.emplace_back(); }
void orc_paths_add_waypoint(orc_paths* p, const double pose16[16])
{
  Waypoint w{};
  std::copy(pose16, pose16 + 16, w.T);
  for (int c = 0; c < 3; ++c)
  {
    w.pos[c] = static_cast<float>(pose16[12 + c]);
    w.nrm[c] = static_cast<float>(pose16[8 + c]);
  }
  w.edge_lo = w.edge_hi = 0xFFFFFFFFu;
  p->paths.back().segments.back().push_back(w);
}

uint64_t orc_paths_num_paths(const orc_paths* p) { return p->paths.size(); }
uint64_t orc_paths_num_segments(const orc_paths* p)
{
  uint64_t n = 0;
  for (const auto& path : p->paths)
    n += path.segments.size();
  return n;
}
uint64_t orc_paths_num_waypoints(const orc_paths* p)
{
  uint64_t n = 0;
  for (const auto& path : p->paths)
    for (const auto& s : path.segments)
      n += s.size();
  return n;
}

void orc_paths_export(const orc_paths* p,
                      uint64_t* path_off,
                      uint64_t* seg_off,
                      int64_t* path_plane,
                      float* pos,
                      float* nrm,
                      uint32_t* edge_lo,
                      uint32_t* edge_hi,
                      double* pose16)
{
  uint64_t si = 0, wi = 0, pi = 0;
  for (const auto& path : p->paths)
  {
    if (path_off)
      path_off[pi] = si;
    if (path_plane)
      path_plane[pi] = path.plane;
    ++pi;
    for (const auto& s : path.segments)
    {
      if (seg_off)
        seg_off[si] = wi;
      ++si;
      for (const auto& w : s)
      {
        if (pos)
          std::copy(w.pos, w.pos + 3, pos + 3 * wi);
        if (nrm)
          std::copy(w.nrm, w.nrm + 3, nrm + 3 * wi);
        if (edge_lo)
          edge_lo[wi] = w.edge_lo;
        if (edge_hi)
          edge_hi[wi] = w.edge_hi;
        if (pose16)
          std::copy(w.T, w.T + 16, pose16 + 16 * wi);
        ++wi;
      }
    }
  }
  if (path_off)
    path_off[pi] = si;
  if (seg_off)
    seg_off[si] = wi;
}

int orc_raster_organization(orc_paths* p) { return rasterOrganization(*p); }
int orc_direction_of_travel(orc_paths* p) { return directionOfTravel(*p); }
int orc_snake_organization(orc_paths* p) { return snakeOrganization(*p); }
int orc_linear_approach(orc_paths* p, const double offset[3], int n_points) { return linearApproachDeparture(*p, offset, n_points, false); }
int orc_linear_departure(orc_paths* p, const double offset[3], int n_points) { return linearApproachDeparture(*p, offset, n_points, true); }
int orc_offset(orc_paths* p, const double offset16[16]) { return offsetModifier(*p, offset16); }
int orc_fixed_orientation(orc_paths* p, const double ref_x[3]) { return fixedOrientation(*p, ref_x); }
int orc_uniform_orientation(orc_paths* p) { return uniformOrientation(*p); }
int orc_concatenate(orc_paths* p) { return concatenate(*p); }
int orc_tool_drag_orientation(orc_paths* p, double angle_offset, double tool_radius) { return toolDrag(*p, angle_offset, tool_radius, false); }
int orc_biased_tool_drag_orientation(orc_paths* p, double angle_offset, double tool_radius) { return toolDrag(*p, angle_offset, tool_radius, true); }
int orc_circular_lead_in(orc_paths* p, double arc_angle, double arc_radius, int n_points) { return circularLead(*p, arc_angle, arc_radius, n_points, false); }
int orc_circular_lead_out(orc_paths* p, double arc_angle, double arc_radius, int n_points) { return circularLead(*p, arc_angle, arc_radius, n_points, true); }
int orc_moving_average_orientation_smoothing(orc_paths* p, int window_size) { return movingAverageOrientationSmoothing(*p, window_size); }
int orc_join_close_segments(orc_paths* p, double d) { return joinCloseSegments(*p, d); }
int orc_minimum_segment_length(orc_paths* p, double l) { return minimumSegmentLength(*p, l); }
int orc_uniform_spacing(orc_paths* p, double s) { return uniformSpacing(*p, s); }

void orc_fixture_random_transform(double translation_limit, double rotation_limit, double out16[16])
{
  // tool_path_planner_utest.cpp:32-46.  libstdc++'s std::mt19937 / uniform_real_distribution are the ones the
  // reference's CI links, so they are used directly rather than restated.
  std::mt19937 rand_gen(0);
  std::uniform_real_distribution<double> pos_dist(-std::abs(translation_limit), std::abs(translation_limit));
  double pos[3];
  for (double& v : pos)
    v = pos_dist(rand_gen);
  std::uniform_real_distribution<double> rot_dist(-std::abs(rotation_limit), std::abs(rotation_limit));
  double ea[3];
  for (double& v : ea)
    v = rot_dist(rand_gen);

  auto matmul = [](const double A[9], const double B[9], double C[9]) {  // row-major 3x3
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        C[r * 3 + c] = eigenSum3(A[r * 3 + 0] * B[0 * 3 + c], A[r * 3 + 1] * B[1 * 3 + c], A[r * 3 + 2] * B[2 * 3 + c]);
  };
  const double cx = std::cos(ea[0]), sx = std::sin(ea[0]);
  const double cy = std::cos(ea[1]), sy = std::sin(ea[1]);
  const double cz = std::cos(ea[2]), sz = std::sin(ea[2]);
  // AngleAxisd(angle, UnitX/Y/Z).toRotationMatrix(): (1-c)*axis_i*axis_j + c on the diagonal, -+ s off-diagonal
  const double Rx[9] = { (1.0 - cx) * 1.0 * 1.0 + cx, 0.0 - 0.0, 0.0 + 0.0, 0.0 + 0.0, 0.0 + cx, 0.0 - sx, 0.0 - 0.0, 0.0 + sx,
                         0.0 + cx };
  const double Ry[9] = { 0.0 + cy, 0.0 - 0.0, 0.0 + sy, 0.0 + 0.0, (1.0 - cy) * 1.0 * 1.0 + cy, 0.0 - 0.0, 0.0 - sy, 0.0 + 0.0,
                         0.0 + cy };
  const double Rz[9] = { 0.0 + cz, 0.0 - sz, 0.0 + 0.0, 0.0 + sz, 0.0 + cz, 0.0 - 0.0, 0.0 - 0.0, 0.0 + 0.0,
                         (1.0 - cz) * 1.0 * 1.0 + cz };
  double Rxy[9], R[9];
  matmul(Rx, Ry, Rxy);
  matmul(Rxy, Rz, R);
  const double M[16] = { R[0], R[3], R[6], 0.0, R[1], R[4], R[7], 0.0, R[2], R[5], R[8], 0.0, pos[0], pos[1], pos[2], 1.0 };
  std::copy(M, M + 16, out16);
}

void orc_fixture_plane_mesh(float lx, float ly, const double T[16], float xyz[12], float nrm[12], uint32_t tri[6])
{
  // utils.cpp:266-285.  pcl::transformPointCloud<PointNormal, double> moves xyz only (normals keep (0,0,1));
  // pcl::detail::Transformer<double>::se3: tf(r,0)*x + tf(r,1)*y + tf(r,2)*z + tf(r,3), double, cast to float
  const float src[4][3] = { { -lx / 2, ly / 2, 0 }, { -lx / 2, -ly / 2, 0 }, { lx / 2, -ly / 2, 0 }, { lx / 2, ly / 2, 0 } };
  for (int i = 0; i < 4; ++i)
  {
    const double p[3] = { src[i][0], src[i][1], src[i][2] };
    for (int r = 0; r < 3; ++r)
      xyz[3 * i + r] = static_cast<float>(T[0 * 4 + r] * p[0] + T[1 * 4 + r] * p[1] + T[2 * 4 + r] * p[2] + T[3 * 4 + r]);
    nrm[3 * i + 0] = 0.f;
    nrm[3 * i + 1] = 0.f;
    nrm[3 * i + 2] = 1.f;
  }
  const uint32_t t[6] = { 0, 1, 3, 1, 2, 3 };
  std::copy(t, t + 6, tri);
}

}  // extern "C"
