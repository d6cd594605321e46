// Single-thread frame kernels (sm_100a): the lattice step of frame_dev.cuh for a plan whose extents are already
// cached on the device (a second plan on the same mesh, a retry), and the explicit lattice of nb2_slice.  The first
// plan on a mesh runs the same steps as the last-block epilogues of k_ingest_moments and k_extents instead.
//
//   k_frame_lattice  OBB extents -> scales; PrincipalAxis direction / Centroid origin (or explicit values); the plane
//                    lattice of planImpl :556-586; this rank's plane slab
//   k_frame_set      an explicit lattice handed in by the host (nb2_slice)
//
// Both also put the slicing counters into their start-of-plan state.
#include "frame_dev.cuh"

namespace nb2
{
namespace
{
__global__ void k_frame_lattice(FrameDev* __restrict__ frame, FrameParams prm, unsigned* __restrict__ counters)
{
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  framedev::resetPlanCounters(counters);
  framedev::frameLatticeStep(frame, prm);
}

__global__ void k_frame_set(FrameDev* __restrict__ frame, nb2_lattice lat, unsigned long long plane_begin,
                            unsigned long long plane_end, unsigned plane_cap, unsigned* __restrict__ counters)
{
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  framedev::resetPlanCounters(counters);
  frame->lattice_status = frame::kLatticeOk;
  framedev::publishLattice(frame, lat, 0, 1, plane_begin, plane_end, plane_cap);
}
}  // namespace

void launchFrameLattice(nb2_context* c, const FrameParams& prm)
{
  k_frame_lattice<<<1, 32, 0, c->stream>>>(c->frame.as<FrameDev>(), prm, c->counters.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchFrameSet(nb2_context* c, const nb2_lattice& lat, uint64_t plane_begin, uint64_t plane_end, unsigned plane_cap)
{
  k_frame_set<<<1, 32, 0, c->stream>>>(c->frame.as<FrameDev>(), lat, plane_begin, plane_end, plane_cap,
                                       c->counters.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

}  // namespace nb2
