// Single-thread frame kernels: the scalar arithmetic between the reductions and the slicing kernels, evaluated on the
// device so that a plan needs no host round trip before its hit count is known (sm_100a).
//
//   k_frame_pca      moments -> mean, covariance, eigen axes (pcl::PCA::initCompute + Eigen's 3x3 solver)
//   k_frame_lattice  OBB extents -> scales; PrincipalAxis direction / Centroid origin (or explicit values); the plane
//                    lattice of planImpl :556-586; this rank's plane slab
//   k_frame_set      an explicit lattice handed in by the host (nb2_slice)
//
// The functions are the ones the host entry points use (frame_math.h), compiled without FMA contraction and with IEEE
// division / square root, so the device and the host produce the same bits.
#include "frame_math.h"

namespace nb2
{
namespace
{
__device__ __forceinline__ float fromOrderedBits(unsigned u)
{
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_frame_pca(const MomentPartial* __restrict__ total, const float* __restrict__ pivot, unsigned long long V,
                            FrameDev* __restrict__ frame)
{
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  const MomentPartial t = *total;
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

__device__ void publishLattice(FrameDev* frame, const nb2_lattice& lat, int shard_rank, int shard_count,
                               unsigned long long plane_begin, unsigned long long plane_end, unsigned plane_cap)
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

__global__ void k_frame_lattice(FrameDev* __restrict__ frame, FrameParams prm)
{
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  nb2_pca p = frame->pca;
  if (!frame->have_scales)
  {
    for (int i = 0; i < 3; ++i)
      p.scales[i] = fromOrderedBits(frame->ext[3 + i]) - fromOrderedBits(frame->ext[i]);
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

__global__ void k_frame_set(FrameDev* __restrict__ frame, nb2_lattice lat, unsigned long long plane_begin,
                            unsigned long long plane_end, unsigned plane_cap)
{
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  frame->lattice_status = frame::kLatticeOk;
  publishLattice(frame, lat, 0, 1, plane_begin, plane_end, plane_cap);
}
}  // namespace

void launchFramePca(nb2_context* c, const float* pivot_dev)
{
  const int grid = static_cast<int>(std::min<uint64_t>((c->V + kMomentThreads - 1) / kMomentThreads, kMomentBlocks));
  c->frame.ensure(sizeof(FrameDev));
  k_frame_pca<<<1, 32, 0, c->stream>>>(c->partials.as<MomentPartial>() + grid, pivot_dev, c->V, c->frame.as<FrameDev>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchFrameLattice(nb2_context* c, const FrameParams& prm)
{
  k_frame_lattice<<<1, 32, 0, c->stream>>>(c->frame.as<FrameDev>(), prm);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchFrameSet(nb2_context* c, const nb2_lattice& lat, uint64_t plane_begin, uint64_t plane_end, unsigned plane_cap)
{
  k_frame_set<<<1, 32, 0, c->stream>>>(c->frame.as<FrameDev>(), lat, plane_begin, plane_end, plane_cap);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

}  // namespace nb2
