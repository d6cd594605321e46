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
// parallelism is the segment: one warp per tool path (L1) or per surviving segment (L3).  Whatever is independent per
// waypoint (chord lengths, span lookups, the x and y columns) is spread over the lanes; the sequential recurrences are
// evaluated with the reference's additions in the reference's order (host_math.h), so the poses are bit-identical to
// the sequential code.  Only the resampled poses ever leave the GPU.
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

/**
 * L1: one WARP per tool path.  The chord lengths of 32 consecutive waypoints are computed in parallel (coalesced
 * loads); only their running sum is sequential, and every lane forms its own prefix by adding the warp's chord
 * lengths in waypoint order — ((acc + d0) + d1) + ... — i.e. with exactly the additions, in exactly the order, of
 * the reference's loop, so the result is bit-identical to the sequential sum.
 */
__global__ void __launch_bounds__(128)
    k_legacy_measure(const float* __restrict__ pos, const unsigned* __restrict__ segLen, const unsigned* __restrict__ planeMeta,
                     const uint2* __restrict__ prefix, unsigned nP, double min_hole, double min_len, double spacing,
                     double* __restrict__ cum, uint4* __restrict__ mseg, uint4* __restrict__ planeL, unsigned* __restrict__ err)
{
  const unsigned k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (k >= nP)
    return;
  const unsigned Sk = planeMeta[2 * k + 1];
  const uint2 before = prefix[k];
  unsigned kept = 0, samples = 0;
  unsigned long long w = before.x;  // first waypoint of the next original segment
  unsigned j = 0;
  while (j < Sk)  // (uniform across the warp: every lane evaluates the same scalars)
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
    // running chord length, in the order of computeLength / the knot distances of the reference
    double acc = 0.0;
    if (lane == 0)
      cum[mb] = 0.0;
    for (unsigned long long base = mb + 1; base < me; base += 32)
    {
      const unsigned long long i = base + lane;
      double d = 0.0;
      if (i < me)
        d = norm(pointOf(pos, i) - pointOf(pos, i - 1));
      double run = acc;
#pragma unroll 8
      for (int t = 0; t < 32; ++t)
      {
        const double v = __shfl_sync(kFull, d, t);
        if (t <= static_cast<int>(lane))
          run += v;
      }
      if (i < me)
        cum[i] = run;
      const unsigned long long last = min(static_cast<unsigned long long>(31), me - 1 - base);
      acc = __shfl_sync(kFull, run, static_cast<int>(last));
    }
    if (acc < min_len)
      continue;  // MinimumSegmentLength drops it
    if (me - mb < 3)
    {
      if (lane == 0)
        atomicOr(err, 1u);  // uniform_spacing_modifier.cpp:20-21 throws: fewer than degree + 2 waypoints
      continue;
    }
    unsigned cnt = 2;  // u = 0 and u = 1
    for (double s = firstSample(acc, spacing); s < acc; s += spacing)
      ++cnt;
    if (lane == 0)
      mseg[before.y + kept] = make_uint4(static_cast<unsigned>(mb - before.x), static_cast<unsigned>(me - before.x), cnt, samples);
    ++kept;
    samples += cnt;
  }
  if (lane == 0)
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

/**
 * L3: one WARP per surviving segment.
 *   a. the sample distances s_k (s += spacing, sequential) are generated by every lane in lockstep; the lanes then look
 *      their samples' spans up side by side (bisection in the running chord lengths): position and span direction
 *   b. the reference z is propagated from sample to sample out of registers (the only sequential dependency)
 *   c. all lanes finish their poses: x = tangent, y = z_prev x x, columns normalised
 */
__global__ void __launch_bounds__(128)
    k_legacy_resample(const float* __restrict__ pos, const float* __restrict__ nrm, const double* __restrict__ cum,
                      const unsigned* __restrict__ planeMeta, const uint2* __restrict__ prefix,
                      const uint4* __restrict__ mseg, const uint4* __restrict__ planeL, const uint4* __restrict__ planeLPrefix,
                      unsigned nP, unsigned maxSegsPerPlane, double spacing, double* __restrict__ outPose,
                      float* __restrict__ outPos, float* __restrict__ outNrm, unsigned* __restrict__ outSegLen)
{
  const unsigned long long wid = (static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned k = static_cast<unsigned>(wid / maxSegsPerPlane), m = static_cast<unsigned>(wid % maxSegsPerPlane);
  if (k >= nP || m >= planeL[k].y)
    return;
  const uint2 before = prefix[k];
  const uint4 sg = mseg[before.y + m];
  const unsigned long long mb = static_cast<unsigned long long>(before.x) + sg.x, me = static_cast<unsigned long long>(before.x) + sg.y;
  const unsigned long long n = me - mb;
  const uint4 lp = planeLPrefix[k];
  const unsigned long long o0 = static_cast<unsigned long long>(lp.z) + sg.w;  // first output sample of this segment
  const unsigned cnt = sg.z;
  if (lane == 0)
    outSegLen[lp.y + m] = cnt;
  const double* kd = cum + mb;
  const double l = kd[n - 1];
  auto P = [&](unsigned long long i) { return pointOf(pos, mb + i); };
  // position -> T[12..14], span direction -> T[0..2] (normalised in step c)
  auto stage = [&](unsigned long long o, const D3& p, const D3& tangent) {
    double* T = outPose + 16ull * o;
    T[0] = tangent.x, T[1] = tangent.y, T[2] = tangent.z;
    T[12] = p.x, T[13] = p.y, T[14] = p.z;
  };

  // ---- a: positions and span directions ------------------------------------------------------------------------------
  if (lane == 0)
  {
    stage(o0, P(0), P(1) - P(0));                        // u = 0
    stage(o0 + cnt - 1, P(n - 1), P(n - 1) - P(n - 2));  // u = 1
  }
  {
    // a1: the sample distances are a sequential recurrence (s += spacing): every lane runs it, lane q mod 32 parks s_q in
    //     the (not yet written) last element of sample q's pose
    unsigned q = 0;
    for (double s = firstSample(l, spacing); s < l; s += spacing, ++q)
      if ((q & 31u) == lane)
        outPose[16ull * (o0 + 1 + q) + 15] = s;
    __syncwarp();
    // a2: the lanes look their samples up side by side
    const double eps = DBL_EPSILON;
    for (q = lane; q + 2 < cnt; q += 32)
    {
      const double s = outPose[16ull * (o0 + 1 + q) + 15];
      // first knot >= s - eps (bisection; kd is non-decreasing, kd[0] = 0 < s - eps, kd[n - 1] = l > s)
      unsigned long long lo_i = 0, hi_i = n - 1;
      while (hi_i - lo_i > 1)
      {
        const unsigned long long mid = (lo_i + hi_i) >> 1;
        if (kd[mid] < s - eps)
          lo_i = mid;
        else
          hi_i = mid;
      }
      const unsigned long long at = hi_i;
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
        stage(o0 + 1 + q, P(hit), P(span + 1) - P(span));
        continue;
      }
      // ... otherwise the span is the one ending at the first knot >= s
      unsigned long long hi = at;
      while (kd[hi] < s)
        ++hi;
      const unsigned long long lo = hi - 1;
      const double f = (s - kd[lo]) / (kd[hi] - kd[lo]);
      const D3 a = P(lo), b = P(hi);
      stage(o0 + 1 + q, a + f * (b - a), b - a);
    }
  }
  __syncwarp();

  // ---- b + c: poses.  The reference z is propagated from sample to sample (uniform_spacing_modifier.cpp:39-48):
  //      z_k = normalised x_k x (z_{k-1} x x_k), starting from the z column of the segment's first waypoint
  //      (createTransform: normal.normalized()).  Per chunk of 32 samples every lane loads one span direction; the
  //      recurrence then runs out of registers (shuffles), all lanes in lockstep, and lane t keeps the z it needs ------
  D3 ref_z = normalized(D3{ static_cast<double>(nrm[3 * mb]), static_cast<double>(nrm[3 * mb + 1]), static_cast<double>(nrm[3 * mb + 2]) });
  for (unsigned base = 0; base < cnt; base += 32)
  {
    const unsigned q = base + lane;
    const bool mine = q < cnt;
    double* T = outPose + 16ull * (o0 + (mine ? q : 0));
    D3 x = { 1.0, 0.0, 0.0 };
    if (mine)
      x = D3{ T[0], T[1], T[2] };
    D3 my_prev = ref_z, my_z = ref_z;
    const unsigned m_chunk = min(32u, cnt - base);
    for (unsigned t = 0; t < m_chunk; ++t)
    {
      const D3 xt = { __shfl_sync(kFull, x.x, t), __shfl_sync(kFull, x.y, t), __shfl_sync(kFull, x.z, t) };
      const D3 zn = normalized(cross(xt, cross(ref_z, xt)));
      if (lane == t)
      {
        my_prev = ref_z;
        my_z = zn;
      }
      ref_z = zn;
    }
    if (mine)
    {
      const unsigned long long o = o0 + q;
      const D3 xn = normalized(x), yn = normalized(cross(my_prev, x));
      T[0] = xn.x, T[1] = xn.y, T[2] = xn.z, T[3] = 0.0;
      T[4] = yn.x, T[5] = yn.y, T[6] = yn.z, T[7] = 0.0;
      T[8] = my_z.x, T[9] = my_z.y, T[10] = my_z.z, T[11] = 0.0;
      T[15] = 1.0;
      outPos[3 * o] = static_cast<float>(T[12]), outPos[3 * o + 1] = static_cast<float>(T[13]), outPos[3 * o + 2] = static_cast<float>(T[14]);
      outNrm[3 * o] = static_cast<float>(my_z.x), outNrm[3 * o + 1] = static_cast<float>(my_z.y), outNrm[3 * o + 2] = static_cast<float>(my_z.z);
    }
  }
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
  k_legacy_measure<<<(np + 3) / 4, 128, 0, c->stream>>>(c->outPos.as<float>(), c->outSegLen.as<unsigned>(),
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
  const unsigned long long warps = nP * static_cast<unsigned long long>(max_segs);
  k_legacy_resample<<<static_cast<unsigned>((warps + 3) / 4), 128, 0, c->stream>>>(
      c->outPos.as<float>(), c->outNrm.as<float>(), c->legCum.as<double>(), c->planeMeta.as<unsigned>(),
      c->planePrefix.as<uint2>(), c->legSeg.as<uint4>(), planeL, planeLPrefix, np, max_segs, point_spacing,
      c->legPose.as<double>(), c->legPos.as<float>(), c->legNrm.as<float>(), c->legSegLen.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  c->timer.mark("legacy resample", c->stream);
}

}  // namespace nb2
