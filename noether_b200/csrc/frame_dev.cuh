// Scalar frame steps evaluated by one thread on the device, between the reductions and the slicing kernels, so that
// a plan needs no host round trip before its hit count is known (sm_100a).  They run as the epilogue of the last
// block of the reduction that feeds them (k_ingest_moments -> framePcaStep, k_extents -> frameLatticeStep) or, when
// that reduction is already cached, from the single-thread kernels of kernels_frame.cu.
//
// The functions are the ones the host entry points use (frame_math.h), compiled without FMA contraction and with IEEE
// division / square root, so the device and the host produce the same bits.
#pragma once
#include "frame_math.h"

namespace nb2
{
namespace framedev
{
__device__ __forceinline__ float fromOrderedBits(unsigned u)
{
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

/** moments -> mean, covariance, eigen axes (pcl::PCA::initCompute + Eigen's 3x3 solver); arms the extents slots */
static __device__ __noinline__ void framePcaStep(const MomentPartial& t, const float* __restrict__ pivot,
                                                 unsigned long long V, FrameDev* __restrict__ frame)
{
  const float piv[3] = { pivot[0], pivot[1], pivot[2] };
  nb2_pca p;
  for (int i = 0; i < 3; ++i)
    p.scales[i] = 0.f;
  p.pad_ = 0.f;
  if (V >= 3)
    frame::finalizeMoments(t, V, piv, p);
  else
  {
    // too few points for pcl::PCA; the centroid is still defined (centroid_origin_generator.cpp:13)
    for (int i = 0; i < 9; ++i)
      p.evecs[i] = p.cov[i] = 0.f, p.cov64[i] = 0.0;
    const double n = static_cast<double>(V);
    for (int i = 0; i < 3; ++i)
    {
      p.mean64[i] = static_cast<double>(piv[i]) + t.s[i] / n;
      p.mean[i] = static_cast<float>(p.mean64[i]);
      p.evals[i] = 0.f;
      p.aabb_min[i] = t.mn[i];
      p.aabb_max[i] = t.mx[i];
    }
  }
  frame->pca = p;
  frame->non_finite = t.non_finite;
  // extents accumulators of k_extents: ordered bits of the minima start at the top, of the maxima at the bottom
  for (int i = 0; i < 3; ++i)
  {
    frame->ext[i] = 0xffffffffu;
    frame->ext[3 + i] = 0u;
  }
  frame->have_scales = 0;
}

/** Start-of-plan state of the slicing counters (counters[4..15], see sliceImpl) */
__device__ __forceinline__ void resetPlanCounters(unsigned* __restrict__ counters)
{
  for (int i = 4; i < 16; ++i)
    counters[i] = i == 11 ? 0xffffffffu : 0u;
}

static __device__ __noinline__ void publishLattice(FrameDev* frame, const nb2_lattice& lat, int shard_rank,
                                                   int shard_count, unsigned long long plane_begin,
                                                   unsigned long long plane_end, unsigned plane_cap)
{
  LatticeDev L;
  for (int i = 0; i < 3; ++i)
  {
    L.n[i] = lat.cut_normal[i];
    L.start[i] = lat.start_loc[i];
    L.dir[i] = lat.cut_direction[i];
  }
  L.spacing = lat.line_spacing;
  L.num_planes = static_cast<long long>(lat.num_planes);
  unsigned long long kb = plane_begin, ke = plane_end < lat.num_planes ? plane_end : lat.num_planes;
  if (shard_count > 1)
  {
    kb = lat.num_planes * static_cast<unsigned long long>(shard_rank) / shard_count;
    ke = lat.num_planes * (static_cast<unsigned long long>(shard_rank) + 1ull) / shard_count;
  }
  if (kb > ke)
    kb = ke;
  frame->plane_cap_exceeded = 0;
  if (ke - kb > plane_cap)
  {
    // the host sized the per-plane arrays for fewer planes: make the slicing kernels no-ops and let the host retry
    frame->plane_cap_exceeded = 1;
    frame->planes_wanted = ke - kb;
    ke = kb;
  }
  L.plane_begin = static_cast<long long>(kb);
  L.plane_end = static_cast<long long>(ke);
  frame->lat = lat;
  frame->L = L;
}

/** OBB extents -> scales; PrincipalAxis direction / Centroid origin (or explicit values); the plane lattice of
 *  planImpl :556-586; this rank's plane slab.  The extents slots are read past L1 (k_extents' atomics land in L2). */
static __device__ __noinline__ void frameLatticeStep(FrameDev* frame, const FrameParams& prm)
{
  nb2_pca p = frame->pca;
  if (!frame->have_scales)
  {
    for (int i = 0; i < 3; ++i)
      p.scales[i] = fromOrderedBits(__ldcg(frame->ext + 3 + i)) - fromOrderedBits(__ldcg(frame->ext + i));
    for (int i = 0; i < 3; ++i)
      frame->pca.scales[i] = p.scales[i];
    frame->have_scales = 1;
  }
  double dir[3], origin[3];
  if (prm.direction_mode == NB2_DIRECTION_PRINCIPAL_AXIS)
    frame::principalAxisDirection(p, prm.sin_rotation, prm.cos_rotation, dir);
  else
    for (int i = 0; i < 3; ++i)
      dir[i] = prm.direction[i];
  for (int i = 0; i < 3; ++i)
    origin[i] = prm.origin_mode == NB2_ORIGIN_CENTROID ? static_cast<double>(p.mean[i]) : prm.origin[i];
  nb2_lattice lat;
  lat.no_cuts = 0;
  lat.pad_ = 0;
  lat.num_planes = 0;
  const int code = frame::makeLattice(p, dir, origin, prm.line_spacing, prm.bidirectional, lat);
  frame->lattice_status = code;
  if (code != frame::kLatticeOk)
  {
    lat.num_planes = 0;
    lat.line_spacing = 1.0;
  }
  publishLattice(frame, lat, prm.shard_rank, prm.shard_count, 0ull, ~0ull, prm.plane_cap);
}
}  // namespace framedev
}  // namespace nb2
