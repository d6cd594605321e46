// Host-side planning of a tool-path modifier chain (see modifier_plan.h): structure bookkeeping only — offsets, launch
// tables, rotation constants.  The waypoints themselves never come to the host while the chain runs.
#include "modifier_plan.h"
#include "nb2_internal.h"

#include <algorithm>
#include <cmath>
#include <string>

namespace nb2
{
namespace mod
{
void angleAxisMatrix(double angle, const double axis[3], double R[9])
{
  // Eigen 3.4 AngleAxis::toRotationMatrix(): sin_axis = sin * axis, cos1_axis = (1 - cos) * axis, off-diagonal terms
  // from tmp = cos1_axis[i] * axis[j] -/+ sin_axis[k], diagonal cos1_axis[i] * axis[i] + cos
  const double s = std::sin(angle), c = std::cos(angle);
  const double sa[3] = { s * axis[0], s * axis[1], s * axis[2] };
  const double ca[3] = { (1.0 - c) * axis[0], (1.0 - c) * axis[1], (1.0 - c) * axis[2] };
  double tmp = ca[0] * axis[1];
  R[1] = tmp - sa[2];
  R[3] = tmp + sa[2];
  tmp = ca[0] * axis[2];
  R[2] = tmp + sa[1];
  R[6] = tmp - sa[1];
  tmp = ca[1] * axis[2];
  R[5] = tmp - sa[0];
  R[7] = tmp + sa[0];
  R[0] = ca[0] * axis[0] + c;
  R[4] = ca[1] * axis[1] + c;
  R[8] = ca[2] * axis[2] + c;
}

namespace
{
const double kUnitY[3] = { 0.0, 1.0, 0.0 };
const double kUnitZ[3] = { 0.0, 0.0, 1.0 };

void checkStructureImpl(const Structure& s)
{
  if (s.path_off.empty() || s.seg_off.empty() || s.path_off.front() != 0 || s.seg_off.front() != 0 ||
      s.path_off.back() != s.segments() || s.path_plane.size() != s.paths())
    fail(NB2_ERR_INVALID_ARGUMENT, "tool path offsets are inconsistent");
  for (size_t i = 0; i + 1 < s.path_off.size(); ++i)
    if (s.path_off[i] > s.path_off[i + 1])
      fail(NB2_ERR_INVALID_ARGUMENT, "path offsets must not decrease");
  for (size_t i = 0; i + 1 < s.seg_off.size(); ++i)
    if (s.seg_off[i] > s.seg_off[i + 1])
      fail(NB2_ERR_INVALID_ARGUMENT, "segment offsets must not decrease");
}

uint32_t segLen(const Structure& s, uint32_t seg) { return s.seg_off[seg + 1] - s.seg_off[seg]; }

/** segments of `cur` re-listed as `order` (source segment, flipped), each with n_pre / n_post extra points */
void restructure(const Structure& cur, const std::vector<std::pair<uint32_t, bool>>& order, uint32_t n_pre, uint32_t n_post,
                 Step& st)
{
  st.restructure = true;
  st.out.path_off = cur.path_off;
  st.out.path_plane = cur.path_plane;
  st.out.seg_off.assign(1, 0u);
  st.map.clear();
  uint64_t w = 0;
  for (const auto& o : order)
  {
    SegMap m{};
    m.out_begin = static_cast<uint32_t>(w);
    m.src_begin = cur.seg_off[o.first];
    m.src_len = segLen(cur, o.first);
    m.n_pre = n_pre;
    m.n_post = n_post;
    m.flipped = o.second ? 1u : 0u;
    st.map.push_back(m);
    w += static_cast<uint64_t>(m.src_len) + n_pre + n_post;
    if (w > 0xfffffff0ull)
      fail(NB2_ERR_INVALID_ARGUMENT, "modified tool paths exceed 2^32 waypoints");
    st.out.seg_off.push_back(static_cast<uint32_t>(w));
  }
}

/** RasterOrganizationModifier::modify (raster_organization_modifier.cpp:7-46) decided on the first / last translation
 *  of every segment: segments whose (last - first) opposes the direction of tool path 0 are reversed, the segments of a
 *  tool path are sorted by their first waypoint along that direction, the tool paths by their first waypoint along the
 *  raster direction (estimateRasterDirection, utils.cpp:31-47).  Same arithmetic as the planner's own post-op
 *  (toolpath.cpp::rasterOrganization). */
void organizeRasters(const Structure& cur, const std::vector<double>& ends, const D3& ref, Step& st)
{
  const uint64_t P = cur.paths();
  struct SegRef
  {
    uint32_t seg;
    bool flipped;
  };
  auto first = [&](uint32_t s) { return D3{ ends[6 * s], ends[6 * s + 1], ends[6 * s + 2] }; };
  auto last = [&](uint32_t s) { return D3{ ends[6 * s + 3], ends[6 * s + 4], ends[6 * s + 5] }; };
  auto firstOf = [&](const SegRef& r) { return r.flipped ? last(r.seg) : first(r.seg); };
  std::vector<std::vector<SegRef>> segs(P);
  bool identity = true;
  for (uint64_t p = 0; p < P; ++p)
  {
    auto& v = segs[p];
    for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
    {
      const bool flip = dot(last(s) - first(s), ref) < 0.0;
      identity &= !flip;
      v.push_back({ s, flip });
    }
    if (v.size() > 1)
    {
      const D3 base = firstOf(v[0]);
      std::sort(v.begin(), v.end(), [&](const SegRef& a, const SegRef& b) {
        return dot(firstOf(a) - base, ref) < dot(firstOf(b) - base, ref);
      });
      for (size_t i = 0; i < v.size(); ++i)
        identity &= (v[i].seg == cur.path_off[p] + i);
    }
  }
  std::vector<uint32_t> order(P);
  for (uint64_t p = 0; p < P; ++p)
    order[p] = static_cast<uint32_t>(p);
  {
    const D3 first_wp = firstOf(segs.front().front());
    const D3 first_wp_last = firstOf(segs.back().front());
    const D3 nref = normalized(ref);
    D3 raster_dir = first_wp_last - first_wp;
    raster_dir = raster_dir - dot(raster_dir, nref) * nref;
    raster_dir = normalized(raster_dir);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      return dot(firstOf(segs[a].front()) - first_wp, raster_dir) < dot(firstOf(segs[b].front()) - first_wp, raster_dir);
    });
    for (uint32_t i = 0; i < order.size(); ++i)
      identity &= (order[i] == i);
  }
  if (identity)
  {
    st.device = false;
    return;
  }
  std::vector<std::pair<uint32_t, bool>> seq;
  seq.reserve(cur.segments());
  Structure reordered;
  reordered.path_off.assign(1, 0u);
  for (uint64_t p = 0; p < P; ++p)
  {
    for (const SegRef& r : segs[order[p]])
      seq.emplace_back(r.seg, r.flipped);
    reordered.path_off.push_back(static_cast<uint32_t>(seq.size()));
    reordered.path_plane.push_back(cur.path_plane[order[p]]);
  }
  restructure(cur, seq, 0, 0, st);
  st.out.path_off = reordered.path_off;
  st.out.path_plane = reordered.path_plane;
}
}  // namespace

void checkStructure(const Structure& s) { checkStructureImpl(s); }

Step buildStep(const Structure& cur, const nb2_modifier& m, Probe* probe)
{
  {
    Step st;
    st.kind = m.kind;
    st.n_points = m.n_points;
    for (int i = 0; i < 16; ++i)
      st.v[i] = m.v[i];
    st.out = cur;
    st.device = cur.waypoints() > 0;
    const uint64_t P = cur.paths(), S = cur.segments();
    switch (m.kind)
    {
      case NB2_MOD_SNAKE_ORGANIZATION:
      {
        std::vector<std::pair<uint32_t, bool>> order;
        order.reserve(S);
        bool any = false;
        for (uint64_t p = 0; p < P; ++p)
        {
          const uint32_t b = cur.path_off[p], e = cur.path_off[p + 1];
          if (p & 1)
          {
            for (uint32_t s = e; s > b; --s)
              order.emplace_back(s - 1, true);
            any |= e > b;
          }
          else
            for (uint32_t s = b; s < e; ++s)
              order.emplace_back(s, false);
        }
        if (any)
          restructure(cur, order, 0, 0, st);
        else
          st.device = false;
        break;
      }
      case NB2_MOD_LINEAR_APPROACH:
      case NB2_MOD_LINEAR_DEPARTURE:
      {
        if (m.n_points < 0)
          fail(NB2_ERR_INVALID_ARGUMENT, "n_points must not be negative");
        for (uint64_t s = 0; s < S; ++s)
          if (segLen(cur, static_cast<uint32_t>(s)) == 0)
            fail(NB2_ERR_INVALID_ARGUMENT, "empty tool path segment (front() / back() of an empty vector in the reference)");
        if (m.n_points == 0)
        {
          st.device = false;
          break;
        }
        std::vector<std::pair<uint32_t, bool>> order;
        order.reserve(S);
        for (uint64_t s = 0; s < S; ++s)
          order.emplace_back(static_cast<uint32_t>(s), false);
        const bool approach = m.kind == NB2_MOD_LINEAR_APPROACH;
        restructure(cur, order, approach ? m.n_points : 0, approach ? 0 : m.n_points, st);
        break;
      }
      case NB2_MOD_OFFSET:
        break;
      case NB2_MOD_FIXED_ORIENTATION:
      {
        // the constructor stores reference_x_direction.normalized() (fixed_orientation_modifier.cpp:7); the YAML decoder
        // assigns the member directly (:52)
        const D3 raw = { m.v[0], m.v[1], m.v[2] };
        const D3 r = m.v[3] != 0.0 ? raw : normalized(raw);
        st.v[0] = r.x, st.v[1] = r.y, st.v[2] = r.z;
        break;
      }
      case NB2_MOD_UNIFORM_ORIENTATION:
        // tool_paths.at(0).at(0).at(1) (uniform_orientation_modifier.cpp:8-9)
        if (P == 0 || cur.path_off[1] == 0 || segLen(cur, 0) < 2)
          fail(NB2_ERR_INVALID_ARGUMENT, "UniformOrientation needs two waypoints in the first segment of the first tool path "
                                         "(vector::_M_range_check in the reference)");
        angleAxisMatrix(M_PI, kUnitZ, st.R[0]);
        break;
      case NB2_MOD_CONCATENATE:
        st.device = false;
        st.out.path_off = { 0u, static_cast<uint32_t>(S) };
        st.out.path_plane = { P ? cur.path_plane.front() : int64_t(-1) };
        break;
      case NB2_MOD_TOOL_DRAG_ORIENTATION:
      case NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION:
      {
        for (uint64_t p = 0; p < P; ++p)
        {
          uint64_t steps = 0;
          for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
            steps += segLen(cur, s) > 1 ? segLen(cur, s) - 1 : 0;
          if (steps < 1)
            fail(NB2_ERR_TOO_FEW_POINTS, "Insufficient number of points to calculate average tool path direction");
          if (segLen(cur, cur.path_off[p]) == 0)
            fail(NB2_ERR_INVALID_ARGUMENT, "first segment of a tool path is empty (tool_path.at(0).at(0) in the reference)");
        }
        const double angle = m.v[0], radius = m.v[1];
        angleAxisMatrix(1.0 * -angle, kUnitY, st.R[0]);
        angleAxisMatrix(-1.0 * -angle, kUnitY, st.R[1]);
        if (m.kind == NB2_MOD_TOOL_DRAG_ORIENTATION)
          st.v[2] = radius * std::sin(angle);  // z_offset (tool_drag_orientation_modifier.cpp:28)
        break;
      }
      case NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION:
        for (uint64_t s = 0; s < S; ++s)
          if (segLen(cur, static_cast<uint32_t>(s)) < 2)
            fail(NB2_ERR_TOO_FEW_POINTS, "segment with a single waypoint (segment[i - 1] out of bounds in the reference)");
        break;
      case NB2_MOD_CIRCULAR_LEAD_IN:
      case NB2_MOD_CIRCULAR_LEAD_OUT:
      {
        if (m.n_points < 0)
          fail(NB2_ERR_INVALID_ARGUMENT, "n_points must not be negative");
        for (uint64_t p = 0; p < P; ++p)
        {
          uint64_t steps = 0;
          for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
            steps += segLen(cur, s) > 1 ? segLen(cur, s) - 1 : 0;
          if (steps < 1)
            fail(NB2_ERR_TOO_FEW_POINTS, "Insufficient number of points to calculate average tool path direction");
        }
        for (uint64_t s = 0; s < S; ++s)
          if (segLen(cur, static_cast<uint32_t>(s)) == 0)
            fail(NB2_ERR_INVALID_ARGUMENT, "empty tool path segment (front() / back() of an empty vector in the reference)");
        if (m.n_points == 0)
        {
          st.device = false;
          break;
        }
        const bool lead_in = m.kind == NB2_MOD_CIRCULAR_LEAD_IN;
        // delta_theta = arc_angle_ / (n_points_ - 1) with a double n_points_ (circular_lead_in_modifier.cpp:14)
        const double delta_theta = m.v[0] / (static_cast<double>(m.n_points) - 1);
        st.rot.resize(18 * static_cast<size_t>(m.n_points));
        for (int sgn = 0; sgn < 2; ++sgn)
          for (int i = 0; i < m.n_points; ++i)
          {
            const double sign = sgn == 0 ? 1.0 : -1.0;
            const double angle = lead_in ? sign * delta_theta * (i + 1) : sign * -delta_theta * (i + 1);
            angleAxisMatrix(angle, kUnitY, &st.rot[9 * (static_cast<size_t>(sgn) * m.n_points + i)]);
          }
        std::vector<std::pair<uint32_t, bool>> order;
        order.reserve(S);
        for (uint64_t s = 0; s < S; ++s)
          order.emplace_back(static_cast<uint32_t>(s), false);
        restructure(cur, order, lead_in ? m.n_points : 0, lead_in ? 0 : m.n_points, st);
        break;
      }
      case NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING:
        if (m.n_points < 0)
          fail(NB2_ERR_INVALID_ARGUMENT, "window_size must not be negative");
        break;
      case NB2_MOD_JOIN_CLOSE_SEGMENTS:
      {
        st.device = false;  // the segments of a tool path are adjacent in memory: joining two of them drops an offset
        bool any = false;
        for (uint64_t p = 0; p < P; ++p)
          any |= cur.path_off[p + 1] - cur.path_off[p] >= 2;
        if (!any || cur.waypoints() == 0)
          break;
        for (uint64_t p = 0; p < P; ++p)
          if (cur.path_off[p + 1] - cur.path_off[p] >= 2)
            for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
              if (segLen(cur, s) == 0)
                fail(NB2_ERR_INVALID_ARGUMENT, "empty tool path segment (front() / back() of an empty vector in the reference)");
        if (!probe)
          fail(NB2_ERR_INTERNAL, "JoinCloseSegments needs the poses");
        std::vector<double> ends;
        double dir0[3];
        probe->rasterInputs(cur, ends, dir0);
        Structure out;
        out.path_off.assign(1, 0u);
        out.seg_off.assign(1, 0u);
        out.path_plane = cur.path_plane;
        for (uint64_t p = 0; p < P; ++p)
        {
          for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
          {
            bool joined = false;
            if (s > cur.path_off[p])
            {
              // segment->back() is the last waypoint of the previous source segment, whether or not that one was merged
              const D3 back = { ends[6 * (s - 1) + 3], ends[6 * (s - 1) + 4], ends[6 * (s - 1) + 5] };
              const D3 front = { ends[6 * s], ends[6 * s + 1], ends[6 * s + 2] };
              joined = norm(back - front) < m.v[0];
            }
            if (joined)
              out.seg_off.back() = cur.seg_off[s + 1];
            else
              out.seg_off.push_back(cur.seg_off[s + 1]);
          }
          out.path_off.push_back(static_cast<uint32_t>(out.seg_off.size() - 1));
        }
        st.out = out;
        break;
      }
      case NB2_MOD_MINIMUM_SEGMENT_LENGTH:
      {
        if (S == 0)
        {
          // (tool paths without segments are erased)
          st.device = false;
          st.out.path_off.assign(1, 0u);
          st.out.path_plane.clear();
          break;
        }
        for (uint64_t s = 0; s < S; ++s)
          if (segLen(cur, static_cast<uint32_t>(s)) == 0)
            fail(NB2_ERR_INVALID_ARGUMENT, "empty tool path segment (segment.size() - 1 wraps around in computeLength)");
        if (!probe)
          fail(NB2_ERR_INTERNAL, "MinimumSegmentLength needs the poses");
        std::vector<double> lengths;
        probe->segmentLengths(cur, lengths);
        std::vector<std::pair<uint32_t, bool>> keep;
        Structure kept;
        kept.path_off.assign(1, 0u);
        for (uint64_t p = 0; p < P; ++p)
        {
          const size_t before = keep.size();
          for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
            if (!(lengths[s] < m.v[0]))
              keep.emplace_back(s, false);
          if (keep.size() > before)
          {
            kept.path_off.push_back(static_cast<uint32_t>(keep.size()));
            kept.path_plane.push_back(cur.path_plane[p]);
          }
        }
        if (keep.size() == S)
        {
          // nothing is dropped; only tool paths that had no segment to begin with disappear
          st.device = false;
          Structure out = cur;
          out.path_off = kept.path_off;
          out.path_plane = kept.path_plane;
          st.out = out;
          break;
        }
        restructure(cur, keep, 0, 0, st);
        st.out.path_off = kept.path_off;
        st.out.path_plane = kept.path_plane;
        st.device = true;  // (even when nothing is left: the executor handles an empty destination)
        break;
      }
      case NB2_MOD_RASTER_ORGANIZATION:
      {
        if (P == 0)
          fail(NB2_ERR_INVALID_ARGUMENT, "RasterOrganization on empty tool paths (vector::_M_range_check: tool_paths.at(0) in the "
                                         "reference; RasterPlanner::plan returns before it gets there)");
        uint64_t steps = 0;
        for (uint32_t s = cur.path_off[0]; s < cur.path_off[1]; ++s)
          steps += segLen(cur, s) > 1 ? segLen(cur, s) - 1 : 0;
        if (steps < 1)
          fail(NB2_ERR_TOO_FEW_POINTS, "Insufficient number of points to calculate average tool path direction");
        for (uint64_t p = 0; p < P; ++p)
        {
          if (cur.path_off[p] == cur.path_off[p + 1])
            fail(NB2_ERR_INVALID_ARGUMENT, "tool path without segments (tool_path.at(0) in the reference)");
          for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
            if (segLen(cur, s) == 0)
              fail(NB2_ERR_INVALID_ARGUMENT, "empty tool path segment (front() / back() of an empty vector in the reference)");
        }
        if (!probe)
          fail(NB2_ERR_INTERNAL, "RasterOrganization needs the poses");
        std::vector<double> ends;
        double dir0[3];
        probe->rasterInputs(cur, ends, dir0);
        organizeRasters(cur, ends, D3{ dir0[0], dir0[1], dir0[2] }, st);
        break;
      }
      default:
        fail(NB2_ERR_INVALID_ARGUMENT, "unknown modifier kind " + std::to_string(m.kind));
    }
    return st;
  }
}
}  // namespace mod
}  // namespace nb2
