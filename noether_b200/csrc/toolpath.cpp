// Host-side tool-path bookkeeping on the compact SoA result: the raster post-ops RasterPlanner::plan hard-wires
// (raster_planner.cpp:16-35), the expansion to Eigen::Isometry3d storage, and the legacy planner's modifiers.
//
//   rasterOrganization   RasterOrganizationModifier::modify            raster_organization_modifier.cpp:7-46
//                        estimateToolPathDirection / estimateRasterDirection   utils.cpp:7-47
//   expandPoses          createTransform (plane_slicer_raster_planner.cpp:380-394) then
//                        DirectionOfTravelOrientationModifier::modify  direction_of_travel_orientation_modifier.cpp:6-48
//   legacyModifiers      JoinCloseSegments / MinimumSegmentLength / UniformSpacing(degree 1, endpoints)
//                        join_close_segments_modifier.cpp:11-39, minimum_segment_length_modifier.cpp:12-37,
//                        uniform_spacing_modifier.cpp:18-70, splines.cpp:38-58,149-229
//
// These are O(segments) decisions plus one streaming pass over the waypoints; the waypoints themselves were produced
// on the device.  The streaming passes are spread over the host cores.
#include "host_math.h"
#include "nb2_internal.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <numeric>
#include <thread>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace nb2
{
namespace
{
/** fn(0), fn(1), ... fn(n - 1), each once, started in that order by as many threads as the machine has */
void parallelInOrder(uint64_t n, const std::function<void(uint64_t)>& fn)
{
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0)
    hw = 1;
  const uint64_t workers = std::min<uint64_t>(hw, n);
  std::atomic<uint64_t> next{ 0 };
  auto run = [&]() {
    for (uint64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1))
      fn(i);
  };
  std::vector<std::thread> th;
  for (uint64_t w = 1; w < workers; ++w)
    th.emplace_back(run);
  run();
  for (auto& t : th)
    t.join();
}

void parallelFor(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)>& fn)
{
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0)
    hw = 1;
  uint64_t chunks = std::min<uint64_t>(hw, (n + grain - 1) / std::max<uint64_t>(grain, 1));
  if (chunks <= 1)
  {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  const uint64_t per = (n + chunks - 1) / chunks;
  for (uint64_t c = 0; c < chunks; ++c)
  {
    const uint64_t b = c * per, e = std::min(n, b + per);
    if (b >= e)
      break;
    th.emplace_back(fn, b, e);
  }
  for (auto& t : th)
    t.join();
}

inline D3 trans(const nb2_result& r, uint64_t i)
{
  if (r.poses)
    return { r.poses[16 * i + 12], r.poses[16 * i + 13], r.poses[16 * i + 14] };
  return fromFloat(r.positions + 3 * i);
}

struct SegRef
{
  uint32_t seg;   // index into the source result's segments
  bool flipped;
};

/** Copy src into a fresh result with segments taken in the given order (per path) and direction. */
nb2_result* permuted(const nb2_result& src, const std::vector<uint32_t>& path_order,
                     const std::vector<std::vector<SegRef>>& segs_of_path)
{
  nb2_result* dst = allocResult(src.owner, src.n_paths, src.n_segments, src.n_waypoints, src.edge_lo != nullptr,
                                src.poses != nullptr);
  dst->post_ops_applied = src.post_ops_applied;
  dst->legacy = src.legacy;
  dst->params = src.params;
  dst->lattice = src.lattice;
  uint32_t s_out = 0, w_out = 0;
  for (uint64_t p = 0; p < src.n_paths; ++p)
  {
    const uint32_t sp = path_order[p];
    dst->path_offsets[p] = s_out;
    dst->path_plane[p] = src.path_plane[sp];
    for (const SegRef& sr : segs_of_path[sp])
    {
      const uint32_t b = src.segment_offsets[sr.seg], e = src.segment_offsets[sr.seg + 1];
      dst->segment_offsets[s_out++] = w_out;
      for (uint32_t i = 0; i < e - b; ++i)
      {
        const uint32_t from = sr.flipped ? (e - 1 - i) : (b + i);
        std::memcpy(dst->positions + 3 * w_out, src.positions + 3 * from, 12);
        std::memcpy(dst->normals + 3 * w_out, src.normals + 3 * from, 12);
        if (src.edge_lo)
        {
          dst->edge_lo[w_out] = src.edge_lo[from];
          dst->edge_hi[w_out] = src.edge_hi[from];
        }
        if (src.poses)
          std::memcpy(dst->poses + 16 * w_out, src.poses + 16 * from, 128);
        ++w_out;
      }
    }
  }
  dst->path_offsets[src.n_paths] = s_out;
  dst->segment_offsets[s_out] = w_out;
  return dst;
}
}  // namespace

void rasterOrganization(nb2_result*& rp)
{
  nb2_result& r = *rp;
  if (r.n_paths == 0)
    return;

  // reference direction: mean of the normalised steps of tool path 0 (utils.cpp:7-29)
  D3 ref{ 0, 0, 0 };
  {
    int n = 0;
    for (uint32_t s = r.path_offsets[0]; s < r.path_offsets[1]; ++s)
    {
      const uint32_t b = r.segment_offsets[s], e = r.segment_offsets[s + 1];
      for (uint32_t i = b; i + 1 < e; ++i)
      {
        ref = ref + normalized(trans(r, i + 1) - trans(r, i));
        ++n;
      }
    }
    if (n < 1)
      fail(NB2_ERR_TOO_FEW_POINTS, "Insufficient number of points to calculate average tool path direction");
    ref = normalized(ref / static_cast<double>(n));
  }

  bool identity = true;
  std::vector<std::vector<SegRef>> segs(r.n_paths);
  auto firstOf = [&](const SegRef& sr) {
    const uint32_t b = r.segment_offsets[sr.seg], e = r.segment_offsets[sr.seg + 1];
    return trans(r, sr.flipped ? e - 1 : b);
  };
  for (uint64_t p = 0; p < r.n_paths; ++p)
  {
    auto& v = segs[p];
    for (uint32_t s = r.path_offsets[p]; s < r.path_offsets[p + 1]; ++s)
    {
      const uint32_t b = r.segment_offsets[s], e = r.segment_offsets[s + 1];
      const D3 diff = trans(r, e - 1) - trans(r, b);
      const bool flip = dot(diff, ref) < 0.0;
      identity &= !flip;
      v.push_back({ s, flip });
    }
    if (v.size() > 1)
    {
      // The reference's comparator subtracts tool_path.at(0).at(0) on both sides (a common offset);
      // the offset is taken once, before sorting.
      const D3 base = firstOf(v[0]);
      std::sort(v.begin(), v.end(), [&](const SegRef& a, const SegRef& b) {
        return dot(firstOf(a) - base, ref) < dot(firstOf(b) - base, ref);
      });
      for (size_t i = 0; i < v.size(); ++i)
        identity &= (v[i].seg == r.path_offsets[p] + i);
    }
  }

  // order of the tool paths along the raster direction (utils.cpp:31-47)
  std::vector<uint32_t> order(r.n_paths);
  std::iota(order.begin(), order.end(), 0u);
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
    return;
  nb2_result* out = permuted(r, order, segs);
  freeResult(rp);
  rp = out;
}

namespace
{
/** dst(s) = where the poses of segment s go */
#if defined(__x86_64__)
// createTransform (+ DirectionOfTravelOrientationModifier) for four waypoints at a time.  The expansion is bound by the
// latency of double-precision square roots and divisions (three to five normalisations per waypoint); the vector forms
// are IEEE operations like the scalar ones and nothing is contracted into FMA, so every lane equals the scalar code above
// bit for bit (tests/test_toolpath_host.py compares the two).
struct V3
{
  __m256d x, y, z;
};
__attribute__((target("avx2"))) inline V3 vcross(const V3& a, const V3& b)
{
  return { _mm256_sub_pd(_mm256_mul_pd(a.y, b.z), _mm256_mul_pd(a.z, b.y)), _mm256_sub_pd(_mm256_mul_pd(a.z, b.x), _mm256_mul_pd(a.x, b.z)),
           _mm256_sub_pd(_mm256_mul_pd(a.x, b.y), _mm256_mul_pd(a.y, b.x)) };
}
__attribute__((target("avx2"))) inline V3 vnormalized(const V3& a)
{
  // Eigen 3.4 normalized(): v / sqrt(|v|^2) when |v|^2 > 0, otherwise v unchanged; |v|^2 = x x + (y y + z z)
  const __m256d z = _mm256_add_pd(_mm256_mul_pd(a.x, a.x), _mm256_add_pd(_mm256_mul_pd(a.y, a.y), _mm256_mul_pd(a.z, a.z)));
  const __m256d pos = _mm256_cmp_pd(z, _mm256_setzero_pd(), _CMP_GT_OQ);
  const __m256d d = _mm256_blendv_pd(_mm256_set1_pd(1.0), _mm256_sqrt_pd(z), pos);  // (x / 1.0 == x exactly)
  return { _mm256_div_pd(a.x, d), _mm256_div_pd(a.y, d), _mm256_div_pd(a.z, d) };
}
/** waypoints [i, i + 4) of a segment [b, e): poses to out + 16 i ... (i + 3 < e) */
__attribute__((target("avx2"))) void expandFourAvx2(const float* positions, const float* normals, uint32_t b, uint32_t e, uint32_t i,
                                                    bool dot_modifier, double* out)
{
  alignas(32) double px[4], py[4], pz[4], nx[4], ny[4], nz[4], dx[4], dy[4], dz[4];
  for (int l = 0; l < 4; ++l)
  {
    const uint32_t w = i + l;
    const float* p = positions + 3ull * w;
    const float* n = normals + 3ull * w;
    px[l] = p[0], py[l] = p[1], pz[l] = p[2];
    nx[l] = n[0], ny[l] = n[1], nz[l] = n[2];
    if (w + 1 < e)
      dx[l] = static_cast<double>(p[3]) - px[l], dy[l] = static_cast<double>(p[4]) - py[l], dz[l] = static_cast<double>(p[5]) - pz[l];
    else if (w > b)
      dx[l] = px[l] - static_cast<double>(p[-3]), dy[l] = py[l] - static_cast<double>(p[-2]), dz[l] = pz[l] - static_cast<double>(p[-1]);
    else
      dx[l] = dy[l] = dz[l] = 0.0;
  }
  const V3 n{ _mm256_load_pd(nx), _mm256_load_pd(ny), _mm256_load_pd(nz) };
  const V3 dir{ _mm256_load_pd(dx), _mm256_load_pd(dy), _mm256_load_pd(dz) };
  V3 y = vcross(n, dir);
  V3 x = vnormalized(vcross(y, n));
  y = vnormalized(y);
  const V3 z = vnormalized(n);
  if (dot_modifier)
  {
    const V3 ref_x = vnormalized(dir);
    const V3 zz = vnormalized(z);
    y = vnormalized(vcross(zz, ref_x));
    x = vnormalized(vcross(y, zz));
  }
  alignas(32) double c[9][4];
  _mm256_store_pd(c[0], x.x), _mm256_store_pd(c[1], x.y), _mm256_store_pd(c[2], x.z);
  _mm256_store_pd(c[3], y.x), _mm256_store_pd(c[4], y.y), _mm256_store_pd(c[5], y.z);
  _mm256_store_pd(c[6], z.x), _mm256_store_pd(c[7], z.y), _mm256_store_pd(c[8], z.z);
  for (int l = 0; l < 4; ++l)
  {
    double* T = out + 16ull * (i + l);
    if ((reinterpret_cast<uintptr_t>(T) & 15u) == 0)
    {
      _mm_stream_pd(T + 0, _mm_set_pd(c[1][l], c[0][l]));
      _mm_stream_pd(T + 2, _mm_set_pd(0.0, c[2][l]));
      _mm_stream_pd(T + 4, _mm_set_pd(c[4][l], c[3][l]));
      _mm_stream_pd(T + 6, _mm_set_pd(0.0, c[5][l]));
      _mm_stream_pd(T + 8, _mm_set_pd(c[7][l], c[6][l]));
      _mm_stream_pd(T + 10, _mm_set_pd(0.0, c[8][l]));
      _mm_stream_pd(T + 12, _mm_set_pd(py[l], px[l]));
      _mm_stream_pd(T + 14, _mm_set_pd(1.0, pz[l]));
    }
    else
    {
      T[0] = c[0][l], T[1] = c[1][l], T[2] = c[2][l], T[3] = 0.0;
      T[4] = c[3][l], T[5] = c[4][l], T[6] = c[5][l], T[7] = 0.0;
      T[8] = c[6][l], T[9] = c[7][l], T[10] = c[8][l], T[11] = 0.0;
      T[12] = px[l], T[13] = py[l], T[14] = pz[l], T[15] = 1.0;
    }
  }
}
bool haveAvx2()
{
  static const bool yes = __builtin_cpu_supports("avx2") && !std::getenv("NB2_NO_AVX2");
  return yes;
}
#endif

template <typename Dst>
void expandPosesTo(const nb2_result& r, Dst dst)
{
  if (r.poses)
  {
    parallelFor(r.n_segments, 64, [&](uint64_t sb, uint64_t se) {
      for (uint64_t s = sb; s < se; ++s)
      {
        const uint32_t b = r.segment_offsets[s], e = r.segment_offsets[s + 1];
        std::memcpy(dst(s), r.poses + 16ull * b, sizeof(double) * 16 * (e - b));
      }
    });
    return;
  }
  const bool dot_modifier = r.post_ops_applied;
  // a piece of work = a few thousand waypoints, whatever the segment sizes; the pieces are handed out in order, so that a
  // result whose positions and normals are still arriving from the device (download = 2) is expanded front to back, every
  // worker on the part that has landed
  const uint64_t grain = std::max<uint64_t>(1, 4096 / std::max<uint64_t>(1, r.n_waypoints / std::max<uint64_t>(1, r.n_segments)));
  parallelInOrder((r.n_segments + grain - 1) / grain, [&](uint64_t piece) {
    const uint64_t sb = piece * grain, se = std::min<uint64_t>(r.n_segments, sb + grain);
    waitWaypoints(r, r.segment_offsets[se]);
    for (uint64_t s = sb; s < se; ++s)
    {
      const uint32_t b = r.segment_offsets[s], e = r.segment_offsets[s + 1];
      double* const out = dst(s) - 16ull * b;  // so that waypoint i lands at out + 16 i
      uint32_t i = b;
#if defined(__x86_64__)
      if (haveAvx2())
        for (; i + 4 <= e; i += 4)
          expandFourAvx2(r.positions, r.normals, b, e, i, dot_modifier, out);
#endif
      for (; i < e; ++i)
      {
        const D3 p = fromFloat(r.positions + 3 * i);
        const D3 n = fromFloat(r.normals + 3 * i);
        D3 dir;
        if (i + 1 < e)
          dir = fromFloat(r.positions + 3 * (i + 1)) - p;
        else if (i > b)
          dir = p - fromFloat(r.positions + 3 * (i - 1));
        else
          dir = { 0, 0, 0 };
        // createTransform: y = n x dir, x = y x n, every column normalised
        D3 y = cross(n, dir);
        D3 x = normalized(cross(y, n));
        y = normalized(y);
        D3 z = normalized(n);
        if (dot_modifier)
        {
          // DirectionOfTravelOrientationModifier: keep z, rebuild y and x from the normalised travel direction
          // (column 2 itself is left as createTransform wrote it; only its re-normalised copy enters x and y)
          const D3 ref_x = normalized(dir);
          const D3 zz = normalized(z);
          y = normalized(cross(zz, ref_x));
          x = normalized(cross(y, zz));
        }
        double* T = out + 16ull * i;
#if defined(__x86_64__)
        // 128 bytes per waypoint that nobody reads back soon: streaming stores skip the read-for-ownership of every
        // cache line (the expansion is bound by host memory traffic: 512 MB for cfg3's four million waypoints)
        if ((reinterpret_cast<uintptr_t>(T) & 15u) == 0)
        {
          _mm_stream_pd(T + 0, _mm_set_pd(x.y, x.x));
          _mm_stream_pd(T + 2, _mm_set_pd(0.0, x.z));
          _mm_stream_pd(T + 4, _mm_set_pd(y.y, y.x));
          _mm_stream_pd(T + 6, _mm_set_pd(0.0, y.z));
          _mm_stream_pd(T + 8, _mm_set_pd(z.y, z.x));
          _mm_stream_pd(T + 10, _mm_set_pd(0.0, z.z));
          _mm_stream_pd(T + 12, _mm_set_pd(p.y, p.x));
          _mm_stream_pd(T + 14, _mm_set_pd(1.0, p.z));
          continue;
        }
#endif
        T[0] = x.x, T[1] = x.y, T[2] = x.z, T[3] = 0.0;
        T[4] = y.x, T[5] = y.y, T[6] = y.z, T[7] = 0.0;
        T[8] = z.x, T[9] = z.y, T[10] = z.z, T[11] = 0.0;
        T[12] = p.x, T[13] = p.y, T[14] = p.z, T[15] = 1.0;
      }
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
  });
}
}  // namespace

void expandPoses(const nb2_result& r, double* out)
{
  expandPosesTo(r, [&](uint64_t s) { return out + 16ull * r.segment_offsets[s]; });
}

void expandPosesSegments(const nb2_result& r, double* const* segment_out)
{
  expandPosesTo(r, [&](uint64_t s) { return segment_out[s]; });
}

// ---------------------------------------------------------------------------------------------------------------
// legacy modifiers
// ---------------------------------------------------------------------------------------------------------------
namespace
{
struct Pose
{
  double T[16];
  float nrm[3];
  uint32_t lo, hi;
};
using Seg = std::vector<Pose>;
struct PathN
{
  int64_t plane;
  std::vector<Seg> segs;
};
inline D3 tr(const Pose& p) { return { p.T[12], p.T[13], p.T[14] }; }

/** UniformSpacingModifier::resample, spline degree 1, endpoints included.  A degree-1 chord-length interpolating
 *  B-spline is the polyline itself, so the arc-length lookup is a prefix sum over chord lengths plus a linear
 *  interpolation inside the span; the z axis is propagated pose to pose as the reference does. */
Seg resample(const Seg& in, double spacing)
{
  const size_t n = in.size();
  if (n < 3)
    fail(NB2_ERR_SEGMENT_TOO_SHORT, "Spline cannot be fit to tool path segment with less than degree+2 waypoints");
  std::vector<D3> P(n);
  std::vector<double> kd(n, 0.0);
  for (size_t i = 0; i < n; ++i)
    P[i] = tr(in[i]);
  for (size_t i = 1; i < n; ++i)
    kd[i] = kd[i - 1] + norm(P[i] - P[i - 1]);
  const double l = kd[n - 1];

  struct Sample
  {
    D3 pos, tangent;
  };
  std::vector<Sample> samples;
  samples.push_back({ P[0], P[1] - P[0] });
  const double rem = std::fmod(l, spacing);
  double s = rem < spacing / 2.0 ? spacing / 2.0 + rem / 2.0 : rem / 2.0;
  for (; s < l; s += spacing)
  {
    // first knot within machine epsilon of s (the reference scans all knots; kd is non-decreasing, so the first
    // candidate is the first knot above s - eps)
    size_t hit = n;
    {
      const double eps = std::numeric_limits<double>::epsilon();
      for (size_t i = std::lower_bound(kd.begin(), kd.end(), s - eps) - kd.begin(); i < n && kd[i] <= s + eps; ++i)
        if (std::abs(kd[i] - s) < eps)
        {
          hit = i;
          break;
        }
    }
    if (hit != n)
    {
      const size_t span = std::min(hit, n - 2);
      samples.push_back({ P[hit], P[span + 1] - P[span] });
      continue;
    }
    const size_t hi = std::lower_bound(kd.begin(), kd.end(), s) - kd.begin();
    const size_t lo = hi - 1;
    const double f = (s - kd[lo]) / (kd[hi] - kd[lo]);
    samples.push_back({ P[lo] + f * (P[hi] - P[lo]), P[hi] - P[lo] });
  }
  samples.push_back({ P[n - 1], P[n - 1] - P[n - 2] });

  Seg out;
  out.reserve(samples.size());
  D3 ref_z = { in.front().T[8], in.front().T[9], in.front().T[10] };
  for (const Sample& sm : samples)
  {
    const D3 x = sm.tangent;
    const D3 y = cross(ref_z, x);
    const D3 z = cross(x, y);
    const D3 xn = normalized(x), yn = normalized(y), zn = normalized(z);
    Pose p{};
    const double M[16] = { xn.x, xn.y, xn.z, 0, yn.x, yn.y, yn.z, 0, zn.x, zn.y, zn.z, 0, sm.pos.x, sm.pos.y, sm.pos.z, 1 };
    std::memcpy(p.T, M, sizeof(M));
    p.nrm[0] = static_cast<float>(zn.x);
    p.nrm[1] = static_cast<float>(zn.y);
    p.nrm[2] = static_cast<float>(zn.z);
    p.lo = p.hi = 0xffffffffu;
    out.push_back(p);
    ref_z = zn;
  }
  return out;
}
}  // namespace

void legacyModifiers(nb2_result*& rp, double point_spacing, double min_segment_size, double min_hole_size)
{
  if (!(point_spacing > 0.0))
    fail(NB2_ERR_INVALID_ARGUMENT, "point_spacing must be positive");
  nb2_result& r = *rp;
  std::vector<double> poses(16 * r.n_waypoints);
  expandPoses(r, poses.data());  // createTransform poses (post-ops not yet applied)

  std::vector<PathN> paths(r.n_paths);
  for (uint64_t p = 0; p < r.n_paths; ++p)
  {
    paths[p].plane = r.path_plane[p];
    for (uint32_t s = r.path_offsets[p]; s < r.path_offsets[p + 1]; ++s)
    {
      Seg seg;
      for (uint32_t i = r.segment_offsets[s]; i < r.segment_offsets[s + 1]; ++i)
      {
        Pose q{};
        std::memcpy(q.T, poses.data() + 16ull * i, 128);
        std::memcpy(q.nrm, r.normals + 3 * i, 12);
        q.lo = r.edge_lo ? r.edge_lo[i] : 0xffffffffu;
        q.hi = r.edge_hi ? r.edge_hi[i] : 0xffffffffu;
        seg.push_back(q);
      }
      paths[p].segs.push_back(std::move(seg));
    }
  }

  // JoinCloseSegments: consecutive segments, in planner output order, closer than min_hole_size
  for (PathN& path : paths)
  {
    auto& v = path.segs;
    size_t a = 0, b = 1;
    while (b < v.size())
    {
      if (norm(tr(v[a].back()) - tr(v[b].front())) < min_hole_size)
      {
        v[a].insert(v[a].end(), v[b].begin(), v[b].end());
        v.erase(v.begin() + b);
      }
      else
      {
        ++a;
        ++b;
      }
    }
  }
  // MinimumSegmentLength
  for (auto pit = paths.begin(); pit != paths.end();)
  {
    auto& v = pit->segs;
    for (auto sit = v.begin(); sit != v.end();)
    {
      double len = 0.0;
      for (size_t i = 0; i + 1 < sit->size(); ++i)
        len += norm(tr((*sit)[i + 1]) - tr((*sit)[i]));
      if (len < min_segment_size)
        sit = v.erase(sit);
      else
        ++sit;
    }
    if (v.empty())
      pit = paths.erase(pit);
    else
      ++pit;
  }
  // UniformSpacing
  uint64_t S = 0, Wn = 0;
  for (PathN& path : paths)
    for (Seg& seg : path.segs)
    {
      seg = resample(seg, point_spacing);
      ++S;
      Wn += seg.size();
    }

  nb2_result* out = allocResult(r.owner, paths.size(), S, Wn, true, true);
  out->legacy = true;
  out->post_ops_applied = false;
  out->params = r.params;
  out->lattice = r.lattice;
  uint32_t s_out = 0, w_out = 0;
  for (size_t p = 0; p < paths.size(); ++p)
  {
    out->path_offsets[p] = s_out;
    out->path_plane[p] = paths[p].plane;
    for (const Seg& seg : paths[p].segs)
    {
      out->segment_offsets[s_out++] = w_out;
      for (const Pose& q : seg)
      {
        std::memcpy(out->poses + 16ull * w_out, q.T, 128);
        for (int c = 0; c < 3; ++c)
        {
          out->positions[3 * w_out + c] = static_cast<float>(q.T[12 + c]);
          out->normals[3 * w_out + c] = q.nrm[c];
        }
        out->edge_lo[w_out] = q.lo;
        out->edge_hi[w_out] = q.hi;
        ++w_out;
      }
    }
  }
  out->path_offsets[paths.size()] = s_out;
  out->segment_offsets[s_out] = w_out;
  freeResult(rp);
  rp = out;
}

/** DirectionOfTravelOrientationModifier on explicit poses (legacy results) */
void directionOfTravelInPlace(nb2_result& r)
{
  if (!r.poses)
    return;
  for (uint64_t s = 0; s < r.n_segments; ++s)
  {
    const uint32_t b = r.segment_offsets[s], e = r.segment_offsets[s + 1];
    for (uint32_t i = b; i < e; ++i)
    {
      D3 ref_x;
      if (i + 1 < e)
        ref_x = normalized(trans(r, i + 1) - trans(r, i));
      else
        ref_x = normalized(trans(r, i) - trans(r, i - 1));
      double* T = r.poses + 16ull * i;
      const D3 z = normalized(D3{ T[8], T[9], T[10] });
      const D3 y = normalized(cross(z, ref_x));
      const D3 x = normalized(cross(y, z));
      T[0] = x.x, T[1] = x.y, T[2] = x.z;
      T[4] = y.x, T[5] = y.y, T[6] = y.z;
    }
  }
}

}  // namespace nb2
