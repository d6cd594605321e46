// PlaneSlicerLegacyRasterPlanner's modifiers on the device (sm_100a), applied to the slicer's compact result while it
// is still in HBM (plane_slicer_legacy_raster_planner.cpp:22-33):
//
//   JoinCloseSegmentsToolPathModifier      join_close_segments_modifier.cpp:11-39   (min_hole_size)
//   MinimumSegmentLengthToolPathModifier   minimum_segment_length_modifier.cpp:12-37 (min_segment_size)
//   UniformSpacingModifier, degree 1       uniform_spacing_modifier.cpp:18-70, splines.cpp:38-58,149-229 (point_spacing)
//
// A degree-1 chord-length interpolating B-spline is the polyline itself, so resampling is an arc-length lookup in the
// running sum of the chord lengths plus a linear interpolation inside the span (the closed form the host restatement
// in toolpath.cpp uses as well).  The reference's arithmetic is sequential along a segment — the running sum of chord
// lengths, the sample distances s += point_spacing, the z axis propagated from pose to pose — so the unit of
// parallelism is the segment: one thread per tool path (L1) or per surviving segment (L3) walks its waypoints in
// order, in double, with the host code's operation order (host_math.h), and the rasters of a plan (hundreds to
// thousands) supply the parallelism.  Only the resampled poses ever leave the GPU.
//
//   L1  k_legacy_measure  : per tool path: join consecutive segments closer than min_hole_size, running chord length of
//                           every waypoint (kept for L3), drop segments shorter than min_segment_size, count samples
//   L2  k_legacy_scan     : exclusive prefix over the tool paths of {surviving paths, segments, samples}
//   L3  k_legacy_resample : per surviving segment: sample distances, arc-length lookup with a forward-moving cursor,
//                           poses with the propagated reference z; writes 4x4 poses, float positions and z axes
#include "host_math.h"
#include "nb2_internal.h"

#include <cfloat>

namespace nb2
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ D3 pointOf(const float* __restrict__ pos, unsigned long long i)
{
  return { static_cast<double>(pos[3 * i]), static_cast<double>(pos[3 * i + 1]), static_cast<double>(pos[3 * i + 2]) };
}

/** first sample distance and number of interior samples of a segment of length l (computeEquidistantKnots,
 *  splines.cpp:190-229): the distances are accumulated, s += spacing, exactly as the reference does */
__device__ __forceinline__ double firstSample(double l, double spacing)
{
  const double rem = fmod(l, spacing);
  return rem < spacing / 2.0 ? spacing / 2.0 + rem / 2.0 : rem / 2.0;
}

__global__ void __launch_bounds__(128)
    k_legacy_measure(const float* __restrict__ pos, const unsigned* __restrict__ segLen, const unsigned* __restrict__ planeMeta,
                     const uint2* __restrict__ prefix, unsigned nP, double min_hole, double min_len, double spacing,
                     double* __restrict__ cum, uint4* __restrict__ mseg, uint4* __restrict__ planeL, unsigned* __restrict__ err)
{
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nP)
    return;
  const unsigned Sk = planeMeta[2 * k + 1];
  const uint2 before = prefix[k];
  unsigned kept = 0, samples = 0;
  unsigned long long w = before.x;  // first waypoint of the next original segment
  unsigned j = 0;
  while (j < Sk)
  {
    // merged segment: original segments j .. as long as the gap to the next one is below min_hole_size
    const unsigned long long mb = w;
    unsigned long long me = w + segLen[before.y + j];
    ++j;
    while (j < Sk)
    {
      const D3 gap = pointOf(pos, me - 1) - pointOf(pos, me);  // a.back() - b.front() (b starts right after a)
      if (!(norm(gap) < min_hole))
        break;
      me += segLen[before.y + j];
      ++j;
    }
    w = me;
    // running chord length, sequential as computeLength / the knot distances of the reference
    double acc = 0.0;
    D3 prev = pointOf(pos, mb);
    cum[mb] = 0.0;
    for (unsigned long long i = mb + 1; i < me; ++i)
    {
      const D3 p = pointOf(pos, i);
      acc += norm(p - prev);
      cum[i] = acc;
      prev = p;
    }
    if (acc < min_len)
      continue;  // MinimumSegmentLength drops it
    if (me - mb < 3)
    {
      atomicOr(err, 1u);  // uniform_spacing_modifier.cpp:20-21 throws: fewer than degree + 2 waypoints
      continue;
    }
    unsigned cnt = 2;  // u = 0 and u = 1
    for (double s = firstSample(acc, spacing); s < acc; s += spacing)
      ++cnt;
    mseg[before.y + kept] = make_uint4(static_cast<unsigned>(mb - before.x), static_cast<unsigned>(me - before.x), cnt, samples);
    ++kept;
    samples += cnt;
  }
  planeL[k] = make_uint4(kept ? 1u : 0u, kept, samples, 0u);
}

/** Exclusive prefix over the planes of {surviving paths, segments, samples}; totals at [nP]. */
__global__ void __launch_bounds__(1024) k_legacy_scan(const uint4* __restrict__ planeL, unsigned nP, uint4* __restrict__ out)
{
  __shared__ uint4 s_warp[32];
  __shared__ uint4 s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    s_carry = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (unsigned base = 0; base < nP; base += blockDim.x)
  {
    const unsigned i = base + threadIdx.x;
    const uint4 v = i < nP ? planeL[i] : make_uint4(0u, 0u, 0u, 0u);
    uint4 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const unsigned a = __shfl_up_sync(kFull, x.x, o), b = __shfl_up_sync(kFull, x.y, o), c = __shfl_up_sync(kFull, x.z, o);
      if (lane >= o)
        x.x += a, x.y += b, x.z += c;
    }
    if (lane == 31)
      s_warp[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
      uint4 t = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const unsigned a = __shfl_up_sync(kFull, t.x, o), b = __shfl_up_sync(kFull, t.y, o), c = __shfl_up_sync(kFull, t.z, o);
        if (lane >= o)
          t.x += a, t.y += b, t.z += c;
      }
      s_warp[lane] = t;
    }
    __syncthreads();
    const uint4 carry = s_carry;
    const uint4 wp = warp ? s_warp[warp - 1] : make_uint4(0u, 0u, 0u, 0u);
    const uint4 incl = make_uint4(x.x + wp.x + carry.x, x.y + wp.y + carry.y, x.z + wp.z + carry.z, 0u);
    if (i < nP)
      out[i] = make_uint4(incl.x - v.x, incl.y - v.y, incl.z - v.z, 0u);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1)
      s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0)
    out[nP] = s_carry;
}

__global__ void __launch_bounds__(128)
    k_legacy_resample(const float* __restrict__ pos, const float* __restrict__ nrm, const double* __restrict__ cum,
                      const unsigned* __restrict__ planeMeta, const uint2* __restrict__ prefix,
                      const uint4* __restrict__ mseg, const uint4* __restrict__ planeL, const uint4* __restrict__ planeLPrefix,
                      unsigned nP, unsigned maxSegsPerPlane, double spacing, double* __restrict__ outPose,
                      float* __restrict__ outPos, float* __restrict__ outNrm, unsigned* __restrict__ outSegLen)
{
  // thread = (plane, surviving segment of that plane)
  const unsigned long long tid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned k = static_cast<unsigned>(tid / maxSegsPerPlane), m = static_cast<unsigned>(tid % maxSegsPerPlane);
  if (k >= nP || m >= planeL[k].y)
    return;
  const uint2 before = prefix[k];
  const uint4 sg = mseg[before.y + m];
  const unsigned long long mb = static_cast<unsigned long long>(before.x) + sg.x, me = static_cast<unsigned long long>(before.x) + sg.y;
  const unsigned long long n = me - mb;
  const uint4 lp = planeLPrefix[k];
  unsigned long long o = static_cast<unsigned long long>(lp.z) + sg.w;  // first output sample of this segment
  outSegLen[lp.y + m] = sg.z;
  const double* kd = cum + mb;
  const double l = kd[n - 1];

  // reference z: the z column of the first waypoint's pose (createTransform: normal.normalized())
  D3 ref_z = normalized(D3{ static_cast<double>(nrm[3 * mb]), static_cast<double>(nrm[3 * mb + 1]), static_cast<double>(nrm[3 * mb + 2]) });
  auto emit = [&](const D3& p, const D3& tangent) {
    const D3 x = tangent;
    const D3 y = cross(ref_z, x);
    const D3 z = cross(x, y);
    const D3 xn = normalized(x), yn = normalized(y), zn = normalized(z);
    double* T = outPose + 16ull * o;
    T[0] = xn.x, T[1] = xn.y, T[2] = xn.z, T[3] = 0.0;
    T[4] = yn.x, T[5] = yn.y, T[6] = yn.z, T[7] = 0.0;
    T[8] = zn.x, T[9] = zn.y, T[10] = zn.z, T[11] = 0.0;
    T[12] = p.x, T[13] = p.y, T[14] = p.z, T[15] = 1.0;
    outPos[3 * o] = static_cast<float>(p.x), outPos[3 * o + 1] = static_cast<float>(p.y), outPos[3 * o + 2] = static_cast<float>(p.z);
    outNrm[3 * o] = static_cast<float>(zn.x), outNrm[3 * o + 1] = static_cast<float>(zn.y), outNrm[3 * o + 2] = static_cast<float>(zn.z);
    ref_z = zn;
    ++o;
  };
  auto P = [&](unsigned long long i) { return pointOf(pos, mb + i); };

  emit(P(0), P(1) - P(0));  // u = 0
  // the sample distances only grow, so the lookups share one cursor: `at` = first knot >= s - eps
  const double eps = DBL_EPSILON;
  unsigned long long at = 0;
  for (double s = firstSample(l, spacing); s < l; s += spacing)
  {
    while (at < n && kd[at] < s - eps)
      ++at;
    // estimateSplineParameterAtDistance (splines.cpp:107-144): a knot within epsilon of s is hit exactly ...
    unsigned long long hit = n;
    for (unsigned long long i = at; i < n && kd[i] <= s + eps; ++i)
      if (fabs(kd[i] - s) < eps)
      {
        hit = i;
        break;
      }
    if (hit != n)
    {
      const unsigned long long span = hit < n - 2 ? hit : n - 2;
      emit(P(hit), P(span + 1) - P(span));
      continue;
    }
    // ... otherwise the span is the one ending at the first knot >= s
    unsigned long long hi = at;
    while (kd[hi] < s)
      ++hi;
    const unsigned long long lo = hi - 1;
    const double f = (s - kd[lo]) / (kd[hi] - kd[lo]);
    const D3 a = P(lo), b = P(hi);
    emit(a + f * (b - a), b - a);
  }
  emit(P(n - 1), P(n - 1) - P(n - 2));  // u = 1
}
}  // namespace

/** Launches L1 - L3 on the slab the last plan left in the output buffers.  Sizes the legacy output from the totals of
 *  L2 (one small read-back).  Returns {paths, segments, samples}; the per-plane {kept, segments, samples} land in
 *  host_planeL (nP entries). */
void launchLegacy(nb2_context* c, uint64_t nP, uint64_t W, uint64_t S, double point_spacing, double min_segment_size,
                  double min_hole_size, uint64_t totals[3], std::vector<uint32_t>& host_planeL)
{
  c->legCum.ensure(sizeof(double) * std::max<uint64_t>(W, 1));
  c->legSeg.ensure(sizeof(uint4) * std::max<uint64_t>(S, 1));
  c->legPlane.ensure(sizeof(uint4) * 2 * (nP + 1));
  uint4* planeL = c->legPlane.as<uint4>();
  uint4* planeLPrefix = planeL + (nP + 1);
  unsigned* err = c->counters.as<unsigned>() + 14;
  NB2_CUDA(cudaMemsetAsync(err, 0, sizeof(unsigned), c->stream));
  const unsigned np = static_cast<unsigned>(nP);
  k_legacy_measure<<<(np + 127) / 128, 128, 0, c->stream>>>(c->outPos.as<float>(), c->outSegLen.as<unsigned>(),
                                                            c->planeMeta.as<unsigned>(), c->planePrefix.as<uint2>(), np,
                                                            min_hole_size, min_segment_size, point_spacing,
                                                            c->legCum.as<double>(), c->legSeg.as<uint4>(), planeL, err);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  c->timer.mark("legacy measure", c->stream);
  k_legacy_scan<<<1, 1024, 0, c->stream>>>(planeL, np, planeLPrefix);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;

  // totals, the error flag and the per-plane counts (the only sizes the host needs)
  c->hostSmall.ensure(4096 + sizeof(uint4) * (nP + 1));
  auto* h = c->hostSmall.as<uint32_t>();
  NB2_CUDA(cudaMemcpyAsync(h, planeLPrefix + nP, sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaMemcpyAsync(h + 4, err, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaMemcpyAsync(h + 8, planeL, sizeof(uint4) * nP, cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  if (h[4])
    fail(NB2_ERR_SEGMENT_TOO_SHORT, "Spline cannot be fit to tool path segment with less than degree+2 waypoints");
  totals[0] = h[0];
  totals[1] = h[1];
  totals[2] = h[2];
  host_planeL.assign(h + 8, h + 8 + 4 * nP);
  uint32_t max_segs = 0;
  for (uint64_t k = 0; k < nP; ++k)
    max_segs = std::max(max_segs, host_planeL[4 * k + 1]);
  if (totals[2] == 0 || max_segs == 0)
    return;

  c->legPose.ensure(sizeof(double) * 16 * totals[2]);
  c->legPos.ensure(sizeof(float) * 3 * totals[2]);
  c->legNrm.ensure(sizeof(float) * 3 * totals[2]);
  c->legSegLen.ensure(sizeof(unsigned) * totals[1]);
  const unsigned long long threads = nP * static_cast<unsigned long long>(max_segs);
  k_legacy_resample<<<static_cast<unsigned>((threads + 127) / 128), 128, 0, c->stream>>>(
      c->outPos.as<float>(), c->outNrm.as<float>(), c->legCum.as<double>(), c->planeMeta.as<unsigned>(),
      c->planePrefix.as<uint2>(), c->legSeg.as<uint4>(), planeL, planeLPrefix, np, max_segs, point_spacing,
      c->legPose.as<double>(), c->legPos.as<float>(), c->legNrm.as<float>(), c->legSegLen.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  c->timer.mark("legacy resample", c->stream);
}

}  // namespace nb2
