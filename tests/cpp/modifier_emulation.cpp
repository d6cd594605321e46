// TEST INFRASTRUCTURE ONLY — not part of libnoether_b200.so and never loaded by the product.
//
// Host-side emulation of the launch sequence of nb2_result_modify (noether_b200/csrc/modifiers.cu): the same program
// builder (modifier_plan.cpp) and the same per-waypoint functions (modifier_ops.h, __host__ __device__), with every
// kernel replaced by a loop over its threads.  tests/test_modifiers.py compares it with the oracle on the CPU, so that
// the arithmetic and the bookkeeping of the device path are checked before they reach a GPU; the kernels themselves are
// compared with the oracle by the `-m gpu` tests.
#include "../../noether_b200/csrc/modifier_plan.h"
#include "../../noether_b200/csrc/modifier_smooth.h"
#include "../../noether_b200/csrc/nb2_internal.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace nb2
{
static thread_local std::string g_err;
void setLastError(const std::string& msg) { g_err = msg; }
[[noreturn]] void fail(int status, const std::string& msg) { throw Error{ status, msg }; }
}  // namespace nb2

using namespace nb2;
using namespace nb2::mod;

namespace
{
struct Emu
{
  Structure s;
  std::vector<double> poses;
};
Emu g_last;

D3 tr(const std::vector<double>& P, uint64_t w) { return { P[16 * w + 12], P[16 * w + 13], P[16 * w + 14] }; }

void runMap(const Step& st, const Structure& cur, const std::vector<double>& src, std::vector<double>& dst)
{
  const uint32_t W = static_cast<uint32_t>(cur.waypoints()), S = static_cast<uint32_t>(cur.segments());
  std::vector<int> path_sign;
  std::vector<uint32_t> seg_path(S);
  for (uint64_t p = 0; p < cur.paths(); ++p)
    for (uint32_t k = cur.path_off[p]; k < cur.path_off[p + 1]; ++k)
      seg_path[k] = static_cast<uint32_t>(p);
  if (st.kind == NB2_MOD_TOOL_DRAG_ORIENTATION || st.kind == NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION)
  {
    // k_mod_path_sign
    path_sign.resize(cur.paths());
    for (uint64_t p = 0; p < cur.paths(); ++p)
    {
      D3 acc{ 0, 0, 0 };
      unsigned count = 0;
      for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
      {
        const uint32_t b = cur.seg_off[s], e = cur.seg_off[s + 1];
        if (e - b < 2)
          continue;
        for (uint32_t i = b; i + 1 < e; ++i)
        {
          acc = acc + normalized(tr(src, i + 1) - tr(src, i));
          ++count;
        }
      }
      const D3 dir = normalized(acc / static_cast<double>(static_cast<int>(count)));
      path_sign[p] = toolDragSign(&src[16ull * cur.seg_off[cur.path_off[p]]], dir) > 0.0 ? 1 : -1;
    }
  }
  dst.resize(src.size());
  if (st.kind == NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING)
  {
    // k_mod_quaternions
    std::vector<double> quat(4ull * W);
    for (uint32_t w = 0; w < W; ++w)
      rotationToQuaternion(&src[16ull * w], &quat[4ull * w]);
    // k_mod_smooth: one thread per segment
    for (uint32_t s = 0; s < S; ++s)
    {
      const uint32_t b = cur.seg_off[s], e = cur.seg_off[s + 1];
      const int n = static_cast<int>(e - b), n_each_side = st.n_points / 2;
      for (int i = 0; i < n; ++i)
      {
        double q[4];
        std::memcpy(q, &quat[4ull * (b + i)], sizeof(q));
        if (i > 0 && i < n - 1)
        {
          const int start = std::max(i - n_each_side, 0), stop = std::min(i + n_each_side, n - 1);
          const bool ok = quaternionMean(stop - start + 1,
                                         [&](int k, double* out) { std::memcpy(out, &quat[4ull * (b + start + k)], 32); }, q);
          if (!ok)
            fail(NB2_ERR_NON_FINITE, "Orientation contains NaN values");
        }
        double T[16];
        std::memcpy(T, &src[16ull * (b + i)], sizeof(T));
        quaternionToRotation(q, T);
        std::memcpy(&dst[16ull * (b + i)], T, sizeof(T));
        rotationToQuaternion(T, &quat[4ull * (b + i)]);
      }
    }
    return;
  }
  for (uint32_t w = 0; w < W; ++w)
  {
    double T[16];
    std::memcpy(T, &src[16ull * w], sizeof(T));
    switch (st.kind)
    {
      case NB2_MOD_OFFSET:
        opOffset(T, st.v);
        break;
      case NB2_MOD_FIXED_ORIENTATION:
        opFixedOrientation(T, D3{ st.v[0], st.v[1], st.v[2] });
        break;
      case NB2_MOD_UNIFORM_ORIENTATION:
        opUniformOrientation(T, normalized(tr(src, 1) - tr(src, 0)), st.R[0]);
        break;
      case NB2_MOD_TOOL_DRAG_ORIENTATION:
      case NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION:
      {
        const uint32_t s = segmentOf(cur.seg_off.data(), S, w);
        const int sign = path_sign[seg_path[s]];
        const double* R = sign > 0 ? st.R[0] : st.R[1];
        if (st.kind == NB2_MOD_TOOL_DRAG_ORIENTATION)
          opToolDrag(T, R, st.v[2]);
        else
          opBiasedToolDrag(T, R, static_cast<double>(sign) * st.v[1]);
        break;
      }
      case NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION:
      {
        const uint32_t s = segmentOf(cur.seg_off.data(), S, w);
        const uint32_t e = cur.seg_off[s + 1];
        const D3 here = { T[12], T[13], T[14] };
        const D3 dir = (w + 1 < e) ? tr(src, w + 1) - here : here - tr(src, w - 1);
        opDirectionOfTravel(T, dir);
        break;
      }
      default:
        fail(NB2_ERR_INTERNAL, "modifier kind without a kernel");
    }
    std::memcpy(&dst[16ull * w], T, sizeof(T));
  }
}

std::vector<int> pathSigns(const Structure& cur, const std::vector<double>& src)
{
  std::vector<int> path_sign(cur.paths());
  for (uint64_t p = 0; p < cur.paths(); ++p)
  {
    D3 acc{ 0, 0, 0 };
    unsigned count = 0;
    for (uint32_t s = cur.path_off[p]; s < cur.path_off[p + 1]; ++s)
    {
      const uint32_t b = cur.seg_off[s], e = cur.seg_off[s + 1];
      for (uint32_t i = b; i + 1 < e; ++i)
      {
        acc = acc + normalized(tr(src, i + 1) - tr(src, i));
        ++count;
      }
    }
    const D3 dir = normalized(acc / static_cast<double>(static_cast<int>(count)));
    path_sign[p] = toolDragSign(&src[16ull * cur.seg_off[cur.path_off[p]]], dir) > 0.0 ? 1 : -1;
  }
  return path_sign;
}

void runRestructure(const Step& st, const Structure& cur, const std::vector<double>& src, std::vector<double>& dst)
{
  const bool circular = st.kind == NB2_MOD_CIRCULAR_LEAD_IN || st.kind == NB2_MOD_CIRCULAR_LEAD_OUT;
  std::vector<int> path_sign;
  std::vector<uint32_t> seg_path(cur.segments());
  if (circular)
  {
    path_sign = pathSigns(cur, src);
    for (uint64_t p = 0; p < cur.paths(); ++p)
      for (uint32_t k = cur.path_off[p]; k < cur.path_off[p + 1]; ++k)
        seg_path[k] = static_cast<uint32_t>(p);
  }
  const uint32_t W_out = static_cast<uint32_t>(st.out.waypoints()), S = static_cast<uint32_t>(st.out.segments());
  dst.assign(16ull * W_out, 0.0);
  for (uint32_t w = 0; w < W_out; ++w)
  {
    const uint32_t s = segmentOf(st.out.seg_off.data(), S, w);
    const SegMap m = st.map[s];
    int lead;
    const uint32_t from = restructuredSource(m, w - m.out_begin, &lead);
    double T[16];
    std::memcpy(T, &src[16ull * from], sizeof(T));
    if (lead >= 0 && circular)
    {
      const int neg = path_sign[seg_path[s]] > 0 ? 0 : 1;
      circularLeadPoint(T, &st.rot[9ull * (static_cast<size_t>(neg) * st.n_points + lead)], st.v[1]);
    }
    else if (lead >= 0)
      leadPoint(T, st.v, lead, st.n_points);
    std::memcpy(&dst[16ull * w], T, sizeof(T));
  }
}
}  // namespace

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

/** poses16 == NULL: expand from (positions, normals) as k_mod_expand does */
int emu_modify(uint64_t n_paths, const uint32_t* path_off, uint64_t n_segments, const uint32_t* seg_off,
               const double* poses16, const float* positions, const float* normals, int post_ops_applied,
               const nb2_modifier* chain, int n_chain)
{
  try
  {
    Structure cur;
    cur.path_off.assign(path_off, path_off + n_paths + 1);
    cur.seg_off.assign(seg_off, seg_off + n_segments + 1);
    cur.path_plane.assign(n_paths, -1);
    checkStructure(cur);
    if (n_chain < 0 || (n_chain > 0 && !chain))
      fail(NB2_ERR_INVALID_ARGUMENT, "modifier chain is NULL");
    const uint64_t W0 = cur.waypoints();
    std::vector<double> a(16 * W0), b;
    if (poses16)
      std::memcpy(a.data(), poses16, sizeof(double) * 16 * W0);
    else
      for (uint32_t w = 0; w < W0; ++w)
      {
        const uint32_t s = segmentOf(cur.seg_off.data(), static_cast<uint32_t>(n_segments), w);
        const uint32_t sb = cur.seg_off[s], se = cur.seg_off[s + 1];
        const D3 p = fromFloat(positions + 3ull * w), n = fromFloat(normals + 3ull * w);
        D3 dir{ 0, 0, 0 };
        if (w + 1 < se)
          dir = fromFloat(positions + 3ull * (w + 1)) - p;
        else if (w > sb)
          dir = p - fromFloat(positions + 3ull * (w - 1));
        createTransform(p, n, dir, post_ops_applied != 0, &a[16ull * w]);
      }
    // k_mod_segment_ends + k_mod_path_sign(path 0) as the device probe runs them
    struct HostProbe : Probe
    {
      const std::vector<double>* poses;
      const Structure* dev;  // the probe kernels read the device copy of the segment offsets
      void rasterInputs(const Structure& asked, std::vector<double>& ends, double dir0[3]) override
      {
        Structure cur = asked;
        cur.seg_off = dev->seg_off;
        const std::vector<double>& P = *poses;
        ends.assign(6 * cur.segments(), 0.0);
        for (uint32_t s = 0; s < cur.segments(); ++s)
          if (cur.seg_off[s + 1] > cur.seg_off[s])
          {
            const D3 f = tr(P, cur.seg_off[s]), l = tr(P, cur.seg_off[s + 1] - 1);
            ends[6 * s] = f.x, ends[6 * s + 1] = f.y, ends[6 * s + 2] = f.z;
            ends[6 * s + 3] = l.x, ends[6 * s + 4] = l.y, ends[6 * s + 5] = l.z;
          }
        D3 acc{ 0, 0, 0 };
        unsigned count = 0;
        for (uint32_t s = cur.path_off[0]; s < cur.path_off[1]; ++s)
          for (uint32_t i = cur.seg_off[s]; i + 1 < cur.seg_off[s + 1]; ++i)
          {
            acc = acc + normalized(tr(P, i + 1) - tr(P, i));
            ++count;
          }
        const D3 dir = normalized(acc / static_cast<double>(static_cast<int>(count)));
        dir0[0] = dir.x, dir0[1] = dir.y, dir0[2] = dir.z;
      }
      void segmentLengths(const Structure& asked, std::vector<double>& lengths) override
      {
        Structure cur = asked;
        cur.seg_off = dev->seg_off;
        const std::vector<double>& P = *poses;
        lengths.assign(cur.segments(), 0.0);
        for (uint32_t s = 0; s < cur.segments(); ++s)
        {
          double length = 0.0;
          for (uint32_t i = cur.seg_off[s]; i + 1 < cur.seg_off[s + 1]; ++i)
            length += norm(tr(P, i + 1) - tr(P, i));
          lengths[s] = length;
        }
      }
    } probe;
    // `dev` stands for what the device holds: its segment offsets change only where the executor uploads them
    Structure dev = cur;
    for (int k = 0; k < n_chain; ++k)
    {
      probe.poses = &a;
      probe.dev = &dev;
      dev.path_off = cur.path_off;  // (uploaded by the executor whenever a kernel needs them)
      dev.path_plane = cur.path_plane;
      const Step st = buildStep(cur, chain[k], &probe);
      if (st.device)
      {
        if (st.restructure)
        {
          runRestructure(st, dev, a, b);
          dev.seg_off = st.out.seg_off;
        }
        else
          runMap(st, dev, a, b);
        a.swap(b);
      }
      else if (st.out.seg_off != cur.seg_off)
        dev.seg_off = st.out.seg_off;
      cur = st.out;
    }
    g_last.s = cur;
    g_last.poses = a;
    return NB2_OK;
  }
  catch (const Error& e)
  {
    setLastError(e.msg);
    return e.status;
  }
}

uint64_t emu_num_paths() { return g_last.s.paths(); }
uint64_t emu_num_segments() { return g_last.s.segments(); }
uint64_t emu_num_waypoints() { return g_last.s.waypoints(); }
void emu_export(uint32_t* path_off, uint32_t* seg_off, double* poses16)
{
  std::memcpy(path_off, g_last.s.path_off.data(), sizeof(uint32_t) * g_last.s.path_off.size());
  std::memcpy(seg_off, g_last.s.seg_off.data(), sizeof(uint32_t) * g_last.s.seg_off.size());
  if (!g_last.poses.empty())
    std::memcpy(poses16, g_last.poses.data(), sizeof(double) * g_last.poses.size());
}
}
