// Tool-path modifiers on device-resident poses (reference files listed in modifier_ops.h).
//
// Layout: two pose buffers of 16 doubles per waypoint (Eigen::Isometry3d storage, column-major) that the modifiers of a
// chain ping-pong between, the segment offsets of the current structure and small per-segment / per-path tables.  Every
// modifier is one streaming pass: 128 waypoints per CTA, poses staged through shared memory so that both the global
// loads and the global stores are contiguous 16 KB runs (a thread owns one 128-byte pose; row stride 17 doubles keeps
// the per-thread reads off a single bank), the per-waypoint arithmetic in registers (modifier_ops.h — the same
// functions the host-side emulation in tests/cpp checks against the oracle).  Algorithmic bytes per modifier: 128 W in
// + 128 W' out; bound by HBM bandwidth.
//
//   k_mod_expand        compact float result (cut point, interpolated normal) -> poses: createTransform
//                       (plane_slicer_raster_planner.cpp:380-394) [+ DirectionOfTravelOrientationModifier]
//   k_mod_map           Offset / FixedOrientation / UniformOrientation / ToolDrag / BiasedToolDrag / DirectionOfTravel
//   k_mod_restructure   SnakeOrganization / LinearApproach / LinearDeparture / CircularLeadIn / CircularLeadOut
//                       (segment table -> gather; lead points from the segment's first / last pose)
//   k_mod_path_sign     estimateToolPathDirection (utils.cpp:7-29) per tool path, one CTA each: normalisations on all
//                       warps, the reference's additions in the reference's order; then the tool-drag sign of the path
//   k_mod_quaternions / k_mod_smooth   MovingAverageOrientationSmoothing: quaternions of all poses in parallel, then
//                       the reference's in-place recurrence along each segment, one thread per segment (modifier_smooth.h)
//   k_mod_finish        float translation / z axis of the final poses (what nb2_result_view::positions / normals hold)
#include "modifier_plan.h"
#include "modifier_smooth.h"
#include "nb2_internal.h"

#include <algorithm>
#include <cstring>

namespace nb2
{
namespace
{
using namespace mod;

constexpr int kTile = 128;    // waypoints per CTA
constexpr int kRow = 17;      // doubles per staged pose (16 + 1 pad)

struct MapConsts
{
  int kind;
  double v[16];
  double R[2][9];
};

__device__ __forceinline__ void tileLoad(const double* __restrict__ src, unsigned long long base, int n, double* tile)
{
  const double* g = src + 16ull * base;
  for (int i = threadIdx.x; i < 16 * n; i += kTile)
    tile[(i >> 4) * kRow + (i & 15)] = __ldg(g + i);
}
__device__ __forceinline__ void tileStore(double* __restrict__ dst, unsigned long long base, int n, const double* tile)
{
  double* g = dst + 16ull * base;
  for (int i = threadIdx.x; i < 16 * n; i += kTile)
    g[i] = tile[(i >> 4) * kRow + (i & 15)];
}

__device__ __forceinline__ D3 translationOf(const double* __restrict__ poses, unsigned long long w)
{
  const double* p = poses + 16ull * w + 12;
  return { __ldg(p), __ldg(p + 1), __ldg(p + 2) };
}

/** compact result -> poses */
__global__ void __launch_bounds__(kTile) k_mod_expand(const float* __restrict__ pos, const float* __restrict__ nrm,
                                                      const unsigned* __restrict__ seg_off, unsigned n_segments,
                                                      unsigned W, int dot_modifier, double* __restrict__ dst)
{
  __shared__ double tile[kTile * kRow];
  const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kTile;
  const int n = static_cast<int>(min(static_cast<unsigned long long>(kTile), W - base));
  if (static_cast<int>(threadIdx.x) < n)
  {
    const unsigned w = static_cast<unsigned>(base) + threadIdx.x;
    const unsigned s = segmentOf(seg_off, n_segments, w);
    const unsigned b = seg_off[s], e = seg_off[s + 1];
    const D3 p = fromFloat(pos + 3ull * w);
    const D3 nn = fromFloat(nrm + 3ull * w);
    D3 dir;
    if (w + 1 < e)
      dir = fromFloat(pos + 3ull * (w + 1)) - p;
    else if (w > b)
      dir = p - fromFloat(pos + 3ull * (w - 1));
    else
      dir = { 0.0, 0.0, 0.0 };
    double T[16];
    createTransform(p, nn, dir, dot_modifier != 0, T);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      tile[threadIdx.x * kRow + i] = T[i];
  }
  __syncthreads();
  tileStore(dst, base, n, tile);
}

/** per-waypoint modifiers: dst[w] = op(src[w]); two phases so that neighbour reads of the staged tile never race with the
 *  write-back of a row */
template <int KIND>
__global__ void __launch_bounds__(kTile) k_mod_map_t(const double* __restrict__ src, double* __restrict__ dst, unsigned W,
                                                     const MapConsts c, const unsigned* __restrict__ seg_off,
                                                     unsigned n_segments, const unsigned* __restrict__ seg_path,
                                                     const int* __restrict__ path_sign)
{
  __shared__ double tile[kTile * kRow];
  const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kTile;
  const int n = static_cast<int>(min(static_cast<unsigned long long>(kTile), W - base));
  tileLoad(src, base, n, tile);
  __syncthreads();
  const bool active = static_cast<int>(threadIdx.x) < n;
  double T[16];
  if (active)
  {
    const unsigned w = static_cast<unsigned>(base) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      T[i] = tile[threadIdx.x * kRow + i];
    if (KIND == NB2_MOD_OFFSET)
      opOffset(T, c.v);
    else if (KIND == NB2_MOD_FIXED_ORIENTATION)
      opFixedOrientation(T, D3{ c.v[0], c.v[1], c.v[2] });
    else if (KIND == NB2_MOD_UNIFORM_ORIENTATION)
    {
      // reference direction from the first two waypoints of the first segment of the first tool path
      const D3 ref = normalized(translationOf(src, 1) - translationOf(src, 0));
      opUniformOrientation(T, ref, c.R[0]);
    }
    else if (KIND == NB2_MOD_TOOL_DRAG_ORIENTATION || KIND == NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION)
    {
      const unsigned s = segmentOf(seg_off, n_segments, w);
      const int sign = path_sign[seg_path[s]];
      const double* R = sign > 0 ? c.R[0] : c.R[1];
      if (KIND == NB2_MOD_TOOL_DRAG_ORIENTATION)
        opToolDrag(T, R, c.v[2]);
      else
        opBiasedToolDrag(T, R, static_cast<double>(sign) * c.v[1]);
    }
    else if (KIND == NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION)
    {
      const unsigned s = segmentOf(seg_off, n_segments, w);
      const unsigned e = seg_off[s + 1];
      const D3 here = { T[12], T[13], T[14] };
      // neighbours inside the tile come from shared memory, the others from the source buffer
      D3 dir;
      if (w + 1 < e)
      {
        const int j = threadIdx.x + 1;
        const D3 next = j < n ? D3{ tile[j * kRow + 12], tile[j * kRow + 13], tile[j * kRow + 14] } : translationOf(src, w + 1);
        dir = next - here;
      }
      else
      {
        const int j = static_cast<int>(threadIdx.x) - 1;
        const D3 prev = j >= 0 ? D3{ tile[j * kRow + 12], tile[j * kRow + 13], tile[j * kRow + 14] } : translationOf(src, w - 1);
        dir = here - prev;
      }
      opDirectionOfTravel(T, dir);
    }
  }
  __syncthreads();  // all neighbour reads of the staged tile are done
  if (active)
  {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      tile[threadIdx.x * kRow + i] = T[i];
  }
  __syncthreads();
  tileStore(dst, base, n, tile);
}

/** restructuring modifiers: dst[w] = src[map(w)] (+ approach / departure points) */
__global__ void __launch_bounds__(kTile) k_mod_restructure(const double* __restrict__ src, double* __restrict__ dst,
                                                           unsigned W_out, const SegMap* __restrict__ map,
                                                           const unsigned* __restrict__ seg_off_out, unsigned n_segments,
                                                           double ox, double oy, double oz, int n_points,
                                                           const double* __restrict__ rot, double radius,
                                                           const unsigned* __restrict__ seg_path,
                                                           const int* __restrict__ path_sign)
{
  __shared__ double tile[kTile * kRow];
  const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kTile;
  const int n = static_cast<int>(min(static_cast<unsigned long long>(kTile), W_out - base));
  if (static_cast<int>(threadIdx.x) < n)
  {
    const unsigned w = static_cast<unsigned>(base) + threadIdx.x;
    const unsigned s = segmentOf(seg_off_out, n_segments, w);
    const SegMap m = map[s];
    int lead;
    const unsigned from = restructuredSource(m, w - m.out_begin, &lead);
    double T[16];
    const double2* g = reinterpret_cast<const double2*>(src + 16ull * from);
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
      const double2 v = __ldg(g + i);
      T[2 * i] = v.x;
      T[2 * i + 1] = v.y;
    }
    if (lead >= 0)
    {
      if (rot)  // circular lead in / out: the rotation of step `lead` for the sign of the segment's tool path
      {
        const int neg = path_sign[seg_path[s]] > 0 ? 0 : 1;
        const double* Rg = rot + 9ull * (static_cast<unsigned>(neg) * static_cast<unsigned>(n_points) + static_cast<unsigned>(lead));
        double R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i)
          R[i] = __ldg(Rg + i);
        circularLeadPoint(T, R, radius);
      }
      else
      {
        const double off[3] = { ox, oy, oz };
        leadPoint(T, off, lead, n_points);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
      tile[threadIdx.x * kRow + i] = T[i];
  }
  __syncthreads();
  tileStore(dst, base, n, tile);
}

/** one CTA per tool path: mean of the normalised steps in the reference's order, then the tool-drag sign.
 *  The waypoints of a tool path are contiguous, so the CTA walks them 256 at a time; a thread's step i -> i + 1 counts
 *  when both ends lie in one segment.  The normalisations (an fp64 square root and three divisions each: long dependent
 *  instruction sequences) run on all eight warps; the sum itself has to be the reference's sequential one, so warp 0
 *  then adds the 256 staged steps in order.  Threads without a step stage +0, and acc + (+0) == acc bit for bit (acc
 *  starts at +0 and can never become -0), so the additions run as one unrolled chain. */
constexpr int kSignThreads = 256;
__global__ void __launch_bounds__(kSignThreads) k_mod_path_sign(const double* __restrict__ poses,
                                                                const unsigned* __restrict__ seg_off,
                                                                const unsigned* __restrict__ path_off, unsigned n_paths,
                                                                int* __restrict__ path_sign, double* __restrict__ path_dir)
{
  __shared__ double su[3][kSignThreads];
  const unsigned p = blockIdx.x;
  const unsigned tid = threadIdx.x;
  const unsigned s_begin = path_off[p], s_end = path_off[p + 1];
  const unsigned w_begin = seg_off[s_begin], w_end = seg_off[s_end];
  D3 acc = { 0.0, 0.0, 0.0 };
  unsigned count = 0;
  unsigned s = s_begin;  // segment of this thread's waypoint (advanced monotonically)
  for (unsigned i0 = w_begin; i0 < w_end; i0 += kSignThreads)
  {
    const unsigned i = i0 + tid;
    bool valid = false;
    D3 u = { 0.0, 0.0, 0.0 };
    if (i + 1 < w_end)
    {
      while (i >= seg_off[s + 1])
        ++s;
      valid = i + 1 < seg_off[s + 1];
      if (valid)
        u = normalized(translationOf(poses, i + 1) - translationOf(poses, i));
    }
    su[0][tid] = u.x;
    su[1][tid] = u.y;
    su[2][tid] = u.z;
    count += __syncthreads_count(valid);
    if (tid < 32)
    {
#pragma unroll 16
      for (int k = 0; k < kSignThreads; ++k)
        acc = acc + D3{ su[0][k], su[1][k], su[2][k] };
    }
    __syncthreads();
  }
  if (tid == 0)
  {
    // the host has checked count >= 1 and that the first segment is not empty
    const D3 dir = normalized(acc / static_cast<double>(static_cast<int>(count)));
    const double* first = poses + 16ull * w_begin;
    double F[16];
    for (int i = 0; i < 16; ++i)
      F[i] = first[i];
    path_sign[p] = toolDragSign(F, dir) > 0.0 ? 1 : -1;
    if (path_dir)
    {
      path_dir[3 * p] = dir.x;
      path_dir[3 * p + 1] = dir.y;
      path_dir[3 * p + 2] = dir.z;
    }
  }
}

/** computeLength (utils.cpp:213-234) of every segment, one CTA each: the chord lengths of 256 steps on all warps, then
 *  `length += d` in waypoint order by warp 0 (threads past the end stage +0: adding it is exact) */
__global__ void __launch_bounds__(kSignThreads) k_mod_segment_lengths(const double* __restrict__ poses,
                                                                      const unsigned* __restrict__ seg_off,
                                                                      double* __restrict__ lengths)
{
  __shared__ double sd[kSignThreads];
  const unsigned s = blockIdx.x, tid = threadIdx.x;
  const unsigned b = seg_off[s], e = seg_off[s + 1];
  double length = 0.0;
  for (unsigned i0 = b; i0 + 1 < e; i0 += kSignThreads)
  {
    const unsigned i = i0 + tid;
    sd[tid] = i + 1 < e ? norm(translationOf(poses, i + 1) - translationOf(poses, i)) : 0.0;
    __syncthreads();
    if (tid < 32)
    {
#pragma unroll 16
      for (int k = 0; k < kSignThreads; ++k)
        length += sd[k];
    }
    __syncthreads();
  }
  if (tid == 0)
    lengths[s] = length;
}

/** first and last translation of every segment (RasterOrganization decides on them, on the host) */
__global__ void __launch_bounds__(256) k_mod_segment_ends(const double* __restrict__ poses, const unsigned* __restrict__ seg_off,
                                                          unsigned n_segments, double* __restrict__ ends)
{
  const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segments)
    return;
  const unsigned b = seg_off[s], e = seg_off[s + 1];
  D3 a = { 0.0, 0.0, 0.0 }, z = { 0.0, 0.0, 0.0 };
  if (e > b)
  {
    a = translationOf(poses, b);
    z = translationOf(poses, e - 1);
  }
  ends[6 * s] = a.x, ends[6 * s + 1] = a.y, ends[6 * s + 2] = a.z;
  ends[6 * s + 3] = z.x, ends[6 * s + 4] = z.y, ends[6 * s + 5] = z.z;
}

/** MovingAverageOrientationSmoothing, pass 1: Quaterniond(pose.linear()) of every input pose (fully parallel) */
__global__ void __launch_bounds__(256) k_mod_quaternions(const double* __restrict__ poses, unsigned W, double4* __restrict__ quat)
{
  const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W)
    return;
  double T[16];
  const double2* g = reinterpret_cast<const double2*>(poses + 16ull * w);
#pragma unroll
  for (int i = 0; i < 6; ++i)  // the linear part lives in the first three columns
  {
    const double2 v = __ldg(g + i);
    T[2 * i] = v.x;
    T[2 * i + 1] = v.y;
  }
  double q[4];
  rotationToQuaternion(T, q);
  quat[w] = make_double4(q[0], q[1], q[2], q[3]);
}

/** pass 2: one thread per segment walks it in order.  quat[k] holds the quaternion of the input pose k until pose k has
 *  been smoothed, then the quaternion of the smoothed pose (re-extracted from its rotation matrix, as the reference does
 *  when a later window reads the modified pose).  dst receives the poses (translation copied, linear part replaced). */
__global__ void __launch_bounds__(64) k_mod_smooth(const double* __restrict__ src, double* __restrict__ dst,
                                                   const unsigned* __restrict__ seg_off, unsigned n_segments, int window,
                                                   double4* __restrict__ quat, unsigned* __restrict__ nan_flag)
{
  const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segments)
    return;
  const unsigned b = seg_off[s], e = seg_off[s + 1];
  const int n = static_cast<int>(e - b);
  const int n_each_side = window / 2;
  for (int i = 0; i < n; ++i)
  {
    double q[4];
    {
      const double4 own = quat[b + i];
      q[0] = own.x, q[1] = own.y, q[2] = own.z, q[3] = own.w;
    }
    if (i > 0 && i < n - 1)
    {
      const int start = max(i - n_each_side, 0), stop = min(i + n_each_side, n - 1);
      const bool ok = quaternionMean(stop - start + 1,
                                     [&](int k, double* out) {
                                       const double4 v = quat[b + start + k];
                                       out[0] = v.x, out[1] = v.y, out[2] = v.z, out[3] = v.w;
                                     },
                                     q);
      if (!ok)
      {
        *nan_flag = 1u;
        return;
      }
    }
    double T[16];
    const double* in = src + 16ull * (b + i);
#pragma unroll
    for (int k = 0; k < 16; ++k)
      T[k] = in[k];
    quaternionToRotation(q, T);
    double* out = dst + 16ull * (b + i);
#pragma unroll
    for (int k = 0; k < 16; ++k)
      out[k] = T[k];
    double q2[4];
    rotationToQuaternion(T, q2);
    quat[b + i] = make_double4(q2[0], q2[1], q2[2], q2[3]);
  }
}

__global__ void __launch_bounds__(256) k_mod_finish(const double* __restrict__ poses, unsigned W, float* __restrict__ pos,
                                                    float* __restrict__ nrm)
{
  // thread i handles one float of the output: component c of waypoint w of array a
  const unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 3ull * W)
    return;
  const unsigned long long w = i / 3;
  const unsigned c = static_cast<unsigned>(i - 3 * w);
  pos[i] = static_cast<float>(__ldg(poses + 16ull * w + 12 + c));
  nrm[i] = static_cast<float>(__ldg(poses + 16ull * w + 8 + c));
}

struct DeviceGuard
{
  int prev = -1;
  explicit DeviceGuard(int dev)
  {
    cudaGetDevice(&prev);
    if (prev != dev)
      NB2_CUDA(cudaSetDevice(dev));
    else
      prev = -1;
  }
  ~DeviceGuard()
  {
    if (prev >= 0)
      cudaSetDevice(prev);
  }
};

template <typename T>
void uploadVector(nb2_context* c, DevBuf& buf, const std::vector<T>& v)
{
  buf.ensure(sizeof(T) * std::max<size_t>(v.size(), 1));
  if (!v.empty())
    NB2_CUDA(cudaMemcpyAsync(buf.ptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, c->stream));
}

unsigned tiles(uint64_t W) { return static_cast<unsigned>((W + kTile - 1) / kTile); }

void launchMap(nb2_context* c, const Step& st, const double* src, double* dst, unsigned W, unsigned S)
{
  MapConsts mc;
  mc.kind = st.kind;
  std::memcpy(mc.v, st.v, sizeof(mc.v));
  std::memcpy(mc.R, st.R, sizeof(mc.R));
  const unsigned* so = c->modSegOff.as<unsigned>();
  const unsigned* sp = c->modSegPath.as<unsigned>();
  const int* ps = c->modPathSign.as<int>();
  const unsigned g = tiles(W);
  switch (st.kind)
  {
#define NB2_MAP_CASE(K)                                                                                                \
  case K:                                                                                                              \
    k_mod_map_t<K><<<g, kTile, 0, c->stream>>>(src, dst, W, mc, so, S, sp, ps);                                        \
    break;
    NB2_MAP_CASE(NB2_MOD_OFFSET)
    NB2_MAP_CASE(NB2_MOD_FIXED_ORIENTATION)
    NB2_MAP_CASE(NB2_MOD_UNIFORM_ORIENTATION)
    NB2_MAP_CASE(NB2_MOD_TOOL_DRAG_ORIENTATION)
    NB2_MAP_CASE(NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION)
    NB2_MAP_CASE(NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION)
#undef NB2_MAP_CASE
    default:
      fail(NB2_ERR_INTERNAL, "modifier kind without a kernel");
  }
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

std::vector<uint32_t> segmentPaths(const Structure& s)
{
  std::vector<uint32_t> sp(s.segments());
  for (uint64_t p = 0; p < s.paths(); ++p)
    for (uint32_t k = s.path_off[p]; k < s.path_off[p + 1]; ++k)
      sp[k] = static_cast<uint32_t>(p);
  return sp;
}

const char* stageName(int kind)
{
  switch (kind)
  {
    case NB2_MOD_SNAKE_ORGANIZATION: return "snake organization";
    case NB2_MOD_LINEAR_APPROACH: return "linear approach";
    case NB2_MOD_LINEAR_DEPARTURE: return "linear departure";
    case NB2_MOD_OFFSET: return "offset";
    case NB2_MOD_FIXED_ORIENTATION: return "fixed orientation";
    case NB2_MOD_UNIFORM_ORIENTATION: return "uniform orientation";
    case NB2_MOD_CONCATENATE: return "concatenate";
    case NB2_MOD_TOOL_DRAG_ORIENTATION: return "tool drag orientation";
    case NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION: return "biased tool drag orientation";
    case NB2_MOD_DIRECTION_OF_TRAVEL_ORIENTATION: return "direction of travel orientation";
    case NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING: return "moving average orientation smoothing";
    case NB2_MOD_CIRCULAR_LEAD_IN: return "circular lead in";
    case NB2_MOD_CIRCULAR_LEAD_OUT: return "circular lead out";
    case NB2_MOD_RASTER_ORGANIZATION: return "raster organization";
    case NB2_MOD_JOIN_CLOSE_SEGMENTS: return "join close segments";
    case NB2_MOD_MINIMUM_SEGMENT_LENGTH: return "minimum segment length";
  }
  return "modifier";
}
}  // namespace
}  // namespace nb2

using namespace nb2;
using namespace nb2::mod;

extern "C" {

static nb2_result* resultFromPoses(nb2_context* c, uint64_t n_paths, const uint32_t* path_offsets, uint64_t n_segments,
                                   const uint32_t* segment_offsets, const double* flat, const double* const* per_segment)
{
  if (!c || !path_offsets || !segment_offsets)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  std::lock_guard<std::mutex> lock(c->mu);
  nb2::DeviceGuard g(c->device);
  if (path_offsets[0] != 0 || segment_offsets[0] != 0 || path_offsets[n_paths] != n_segments)
    fail(NB2_ERR_INVALID_ARGUMENT, "tool path offsets are inconsistent");
  for (uint64_t s = 0; s < n_segments; ++s)
    if (segment_offsets[s] > segment_offsets[s + 1])
      fail(NB2_ERR_INVALID_ARGUMENT, "segment offsets must not decrease");
  const uint64_t W = segment_offsets[n_segments];
  if (W && !flat && !per_segment)
    fail(NB2_ERR_INVALID_ARGUMENT, "poses are NULL");
  nb2_result* r = allocResult(c, n_paths, n_segments, W, false, true);
  std::memcpy(r->path_offsets, path_offsets, sizeof(uint32_t) * (n_paths + 1));
  std::memcpy(r->segment_offsets, segment_offsets, sizeof(uint32_t) * (n_segments + 1));
  for (uint64_t p = 0; p < n_paths; ++p)
    r->path_plane[p] = -1;
  if (flat && W)
    std::memcpy(r->poses, flat, sizeof(double) * 16 * W);
  else
    for (uint64_t s = 0; s < n_segments; ++s)
    {
      const uint64_t b = segment_offsets[s], n = segment_offsets[s + 1] - b;
      if (!n)
        continue;
      if (!per_segment[s])
      {
        freeResult(r);
        fail(NB2_ERR_INVALID_ARGUMENT, "segment storage is NULL");
      }
      std::memcpy(r->poses + 16 * b, per_segment[s], sizeof(double) * 16 * n);
    }
  for (uint64_t w = 0; w < W; ++w)
    for (int k = 0; k < 3; ++k)
    {
      r->positions[3 * w + k] = static_cast<float>(r->poses[16 * w + 12 + k]);
      r->normals[3 * w + k] = static_cast<float>(r->poses[16 * w + 8 + k]);
    }
  return r;
}

int nb2_result_from_poses(nb2_context* c, uint64_t n_paths, const uint32_t* path_offsets, uint64_t n_segments,
                          const uint32_t* segment_offsets, const double* poses16, nb2_result** out)
{
  NB2_API_BEGIN
  if (!out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = resultFromPoses(c, n_paths, path_offsets, n_segments, segment_offsets, poses16, nullptr);
  NB2_API_END
}

int nb2_result_from_pose_segments(nb2_context* c, uint64_t n_paths, const uint32_t* path_offsets, uint64_t n_segments,
                                  const uint32_t* segment_offsets, const double* const* segment_poses, nb2_result** out)
{
  NB2_API_BEGIN
  if (!out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = resultFromPoses(c, n_paths, path_offsets, n_segments, segment_offsets, nullptr, segment_poses);
  NB2_API_END
}

int nb2_result_modify(nb2_context* c, const nb2_result* in, const nb2_modifier* chain, int n_chain, nb2_result** out)
{
  NB2_API_BEGIN
  if (in)
    nb2::waitWaypoints(*in, in->n_waypoints);
  if (!c || !in || !out)
    fail(NB2_ERR_INVALID_ARGUMENT, "NULL argument");
  if (in->device_resident)
    fail(NB2_ERR_INVALID_ARGUMENT, "the slab of a sharded plan must be gathered (nb2_result_gather) before it is modified");
  if (in->n_waypoints && !in->poses && (!in->positions || !in->normals))
    fail(NB2_ERR_INVALID_ARGUMENT, "the result carries no tool paths on the host (planned with download = 0)");
  std::lock_guard<std::mutex> lock(c->mu);
  nb2::DeviceGuard g(c->device);

  Structure cur;
  cur.path_off.assign(in->path_offsets, in->path_offsets + in->n_paths + 1);
  cur.seg_off.assign(in->segment_offsets, in->segment_offsets + in->n_segments + 1);
  cur.path_plane.assign(in->path_plane, in->path_plane + in->n_paths);
  checkStructure(cur);
  if (n_chain < 0 || (n_chain > 0 && !chain))
    fail(NB2_ERR_INVALID_ARGUMENT, "modifier chain is NULL");
  for (int k = 0; k < n_chain; ++k)
    if (chain[k].kind < NB2_MOD_SNAKE_ORGANIZATION || chain[k].kind > NB2_MOD_MINIMUM_SEGMENT_LENGTH)
      fail(NB2_ERR_INVALID_ARGUMENT, "unknown modifier kind " + std::to_string(chain[k].kind));

  c->timer.begin(c->stream);
  double* buf[2] = { nullptr, nullptr };
  int at = 0;
  const uint64_t W0 = cur.waypoints();
  const uint64_t w_max = W0;
  // a pose buffer is (re)sized when it becomes the destination of a pass: its old contents are dead by then
  auto ensurePoses = [&](int which, uint64_t waypoints) {
    c->modPose[which].ensure(sizeof(double) * 16 * std::max<uint64_t>(waypoints, 1));
    buf[which] = c->modPose[which].as<double>();
  };
  ensurePoses(0, W0);
  uploadVector(c, c->modSegOff, cur.seg_off);
  if (W0)
  {
    if (in->poses)
    {
      NB2_CUDA(cudaMemcpyAsync(buf[0], in->poses, sizeof(double) * 16 * W0, cudaMemcpyHostToDevice, c->stream));
      c->timer.mark("h2d poses", c->stream);
    }
    else
    {
      c->modPos.ensure(sizeof(float) * 3 * w_max);
      c->modNrm.ensure(sizeof(float) * 3 * w_max);
      NB2_CUDA(cudaMemcpyAsync(c->modPos.ptr, in->positions, sizeof(float) * 3 * W0, cudaMemcpyHostToDevice, c->stream));
      NB2_CUDA(cudaMemcpyAsync(c->modNrm.ptr, in->normals, sizeof(float) * 3 * W0, cudaMemcpyHostToDevice, c->stream));
      c->timer.mark("h2d compact", c->stream);
      k_mod_expand<<<tiles(W0), kTile, 0, c->stream>>>(c->modPos.as<float>(), c->modNrm.as<float>(),
                                                         c->modSegOff.as<unsigned>(), static_cast<unsigned>(cur.segments()),
                                                         static_cast<unsigned>(W0), in->post_ops_applied ? 1 : 0, buf[0]);
      NB2_CUDA(cudaGetLastError());
      ++c->launches;
      c->timer.mark("expand poses", c->stream);
    }
  }

  // RasterOrganization looks at the poses: segment ends and the direction of tool path 0 come back from the device
  struct DeviceProbe : Probe
  {
    nb2_context* c;
    double* const* buf;
    const int* at;
    void rasterInputs(const Structure& cur, std::vector<double>& seg_ends, double dir0[3]) override
    {
      const unsigned S = static_cast<unsigned>(cur.segments());
      c->modEnds.ensure(sizeof(double) * (6ull * S + 8));
      c->modPathSign.ensure(sizeof(int) * 4);
      const std::vector<uint32_t> path0 = { cur.path_off[0], cur.path_off[1] };
      uploadVector(c, c->modPathOff, path0);
      double* d_dir = c->modEnds.as<double>() + 6ull * S;
      k_mod_path_sign<<<1, kSignThreads, 0, c->stream>>>(buf[*at], c->modSegOff.as<unsigned>(), c->modPathOff.as<unsigned>(), 1u,
                                                          c->modPathSign.as<int>(), d_dir);
      NB2_CUDA(cudaGetLastError());
      k_mod_segment_ends<<<(S + 255) / 256, 256, 0, c->stream>>>(buf[*at], c->modSegOff.as<unsigned>(), S, c->modEnds.as<double>());
      NB2_CUDA(cudaGetLastError());
      c->launches += 2;
      seg_ends.resize(6ull * S + 3);
      NB2_CUDA(cudaMemcpyAsync(seg_ends.data(), c->modEnds.ptr, sizeof(double) * (6ull * S + 3), cudaMemcpyDeviceToHost, c->stream));
      NB2_CUDA(cudaStreamSynchronize(c->stream));
      for (int k = 0; k < 3; ++k)
        dir0[k] = seg_ends[6ull * S + k];
      seg_ends.resize(6ull * S);
    }
    void segmentLengths(const Structure& cur, std::vector<double>& lengths) override
    {
      const unsigned S = static_cast<unsigned>(cur.segments());
      c->modEnds.ensure(sizeof(double) * (6ull * S + 8));
      k_mod_segment_lengths<<<S, kSignThreads, 0, c->stream>>>(buf[*at], c->modSegOff.as<unsigned>(), c->modEnds.as<double>());
      NB2_CUDA(cudaGetLastError());
      ++c->launches;
      lengths.resize(S);
      NB2_CUDA(cudaMemcpyAsync(lengths.data(), c->modEnds.ptr, sizeof(double) * S, cudaMemcpyDeviceToHost, c->stream));
      NB2_CUDA(cudaStreamSynchronize(c->stream));
    }
  } probe;
  probe.c = c;
  probe.buf = buf;
  probe.at = &at;

  bool smoothing = false;
  for (int k = 0; k < n_chain; ++k)
  {
    const Step st = buildStep(cur, chain[k], &probe);
    if (st.device)
    {
      ensurePoses(at ^ 1, st.out.waypoints());
      const unsigned W_in = static_cast<unsigned>(cur.waypoints());
      const unsigned S_in = static_cast<unsigned>(cur.segments());
      if (st.restructure)
      {
        const bool circular = st.kind == NB2_MOD_CIRCULAR_LEAD_IN || st.kind == NB2_MOD_CIRCULAR_LEAD_OUT;
        if (circular)
        {
          // sign of every tool path from the poses as they are now (before the new offsets replace the old ones)
          const unsigned P = static_cast<unsigned>(cur.paths());
          uploadVector(c, c->modSegPath, segmentPaths(cur));  // segments keep their tool path and their index
          uploadVector(c, c->modPathOff, cur.path_off);
          uploadVector(c, c->modRot, st.rot);
          c->modPathSign.ensure(sizeof(int) * std::max(P, 1u));
          k_mod_path_sign<<<P, kSignThreads, 0, c->stream>>>(buf[at], c->modSegOff.as<unsigned>(),
                                                               c->modPathOff.as<unsigned>(), P, c->modPathSign.as<int>(), nullptr);
          NB2_CUDA(cudaGetLastError());
          ++c->launches;
        }
        uploadVector(c, c->modSegMap, st.map);
        uploadVector(c, c->modSegOff, st.out.seg_off);  // stream-ordered after the kernels that read the old offsets
        const unsigned W_out = static_cast<unsigned>(st.out.waypoints());
        if (W_out)  // (MinimumSegmentLength may leave nothing)
          k_mod_restructure<<<tiles(W_out), kTile, 0, c->stream>>>(buf[at], buf[at ^ 1], W_out, c->modSegMap.as<SegMap>(),
                                                                   c->modSegOff.as<unsigned>(),
                                                                   static_cast<unsigned>(st.out.segments()), st.v[0], st.v[1],
                                                                   st.v[2], st.n_points,
                                                                   circular ? c->modRot.as<double>() : nullptr, st.v[1],
                                                                   c->modSegPath.as<unsigned>(), c->modPathSign.as<int>());
        NB2_CUDA(cudaGetLastError());
        ++c->launches;
      }
      else if (st.kind == NB2_MOD_MOVING_AVERAGE_ORIENTATION_SMOOTHING)
      {
        c->modQuat.ensure(sizeof(double4) * std::max(W_in, 1u));
        unsigned* flag = c->counters.as<unsigned>() + 17;
        if (!smoothing)
          NB2_CUDA(cudaMemsetAsync(flag, 0, sizeof(unsigned), c->stream));
        k_mod_quaternions<<<(W_in + 255) / 256, 256, 0, c->stream>>>(buf[at], W_in, c->modQuat.as<double4>());
        NB2_CUDA(cudaGetLastError());
        ++c->launches;
        k_mod_smooth<<<(S_in + 63) / 64, 64, 0, c->stream>>>(buf[at], buf[at ^ 1], c->modSegOff.as<unsigned>(), S_in,
                                                              st.n_points, c->modQuat.as<double4>(), flag);
        NB2_CUDA(cudaGetLastError());
        ++c->launches;
        smoothing = true;
      }
      else
      {
        if (st.kind == NB2_MOD_TOOL_DRAG_ORIENTATION || st.kind == NB2_MOD_BIASED_TOOL_DRAG_ORIENTATION)
        {
          const unsigned P = static_cast<unsigned>(cur.paths());
          uploadVector(c, c->modSegPath, segmentPaths(cur));
          uploadVector(c, c->modPathOff, cur.path_off);
          c->modPathSign.ensure(sizeof(int) * std::max(P, 1u));
          k_mod_path_sign<<<P, kSignThreads, 0, c->stream>>>(buf[at], c->modSegOff.as<unsigned>(),
                                                               c->modPathOff.as<unsigned>(), P, c->modPathSign.as<int>(), nullptr);
          NB2_CUDA(cudaGetLastError());
          ++c->launches;
        }
        launchMap(c, st, buf[at], buf[at ^ 1], W_in, S_in);
      }
      at ^= 1;
      c->timer.mark(stageName(st.kind), c->stream);
    }
    else if (st.out.seg_off != cur.seg_off)
      uploadVector(c, c->modSegOff, st.out.seg_off);  // a bookkeeping-only pass that moved segment boundaries (JoinCloseSegments)
    cur = st.out;
  }

  const uint64_t W = cur.waypoints();
  nb2_result* r = allocResult(c, cur.paths(), cur.segments(), W, false, true);
  r->post_ops_applied = in->post_ops_applied;
  r->legacy = in->legacy;
  r->params = in->params;
  r->lattice = in->lattice;
  try
  {
    if (W)
    {
      c->modPos.ensure(sizeof(float) * 3 * W);
      c->modNrm.ensure(sizeof(float) * 3 * W);
      k_mod_finish<<<static_cast<unsigned>((3 * W + 255) / 256), 256, 0, c->stream>>>(buf[at], static_cast<unsigned>(W),
                                                                                      c->modPos.as<float>(), c->modNrm.as<float>());
      NB2_CUDA(cudaGetLastError());
      ++c->launches;
      c->timer.mark("float views", c->stream);
      NB2_CUDA(cudaMemcpyAsync(r->poses, buf[at], sizeof(double) * 16 * W, cudaMemcpyDeviceToHost, c->stream));
      NB2_CUDA(cudaMemcpyAsync(r->positions, c->modPos.ptr, sizeof(float) * 3 * W, cudaMemcpyDeviceToHost, c->stream));
      NB2_CUDA(cudaMemcpyAsync(r->normals, c->modNrm.ptr, sizeof(float) * 3 * W, cudaMemcpyDeviceToHost, c->stream));
      c->timer.mark("d2h", c->stream);
    }
    c->hostSmall.ensure(4096);
    unsigned* h_flag = c->hostSmall.as<unsigned>();
    h_flag[0] = 0;
    if (smoothing)
      NB2_CUDA(cudaMemcpyAsync(h_flag, c->counters.as<unsigned>() + 17, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    NB2_CUDA(cudaStreamSynchronize(c->stream));
    if (h_flag[0])
      fail(NB2_ERR_NON_FINITE, "Orientation contains NaN values");
  }
  catch (...)
  {
    freeResult(r);
    throw;
  }
  std::memcpy(r->path_offsets, cur.path_off.data(), sizeof(uint32_t) * cur.path_off.size());
  std::memcpy(r->segment_offsets, cur.seg_off.data(), sizeof(uint32_t) * cur.seg_off.size());
  if (!cur.path_plane.empty())
    std::memcpy(r->path_plane, cur.path_plane.data(), sizeof(int64_t) * cur.path_plane.size());
  *out = r;
  NB2_API_END
}

}  // extern "C"
