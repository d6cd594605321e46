// Binary STL decoded and welded on the device.
//
// The reference loads STL meshes with pcl::io::loadPolygonFile (noether_gui/src/widgets/tpp_widget.cpp:426):
// vtkSTLReader (binary: 80-byte header, uint32 count, 50-byte facets {normal, v1, v2, v3, attribute} read until the
// file ends) with Merging on — every facet corner goes through vtkMergePoints::InsertUniquePoint in facet order, so
// coincident corners (float equality, -0 == +0) become one point numbered by first occurrence, and facets that
// collapse (two equal ids) are dropped — then pcl::io::vtk2mesh: a PointXYZ cloud and one pcl::Vertices per facet.
// That is a sequential hash insertion of 3 F points on one core.  Here the file image goes to the GPU in one copy and
//
//   T1  k_stl_insert   one thread per corner: open-addressing insert keyed on the coordinate bits (-0 folded), the slot
//                      keeps the smallest corner index of its point (atomicMin) = the first occurrence
//   T2  k_stl_flags    corner is the first occurrence of its point?            -> scan -> point ids in file order
//   T3  k_stl_points   first occurrences write their {x y z 0} record; every facet looks its three ids up and is kept
//                      when they differ                                        -> scan -> facet ids
//   T4  k_stl_compact  kept facets -> uint32 x 3 triangle list
//
// Facet records are 50 bytes (2-byte aligned), so coordinates are assembled from aligned words with a funnel shift, as
// in kernels_ply.cu.  The result is bit-identical to the sequential reader: same points in the same order, same facets.
#include "nb2_internal.h"

namespace nb2
{
namespace
{
constexpr unsigned kEmpty = 0xffffffffu;
constexpr unsigned long long kStlHeader = 84, kStlFacet = 50;

__device__ __forceinline__ unsigned loadU32(const unsigned* __restrict__ img, unsigned long long byte_off)
{
  const unsigned long long w = byte_off >> 2;
  const unsigned sh = static_cast<unsigned>(byte_off & 3u) * 8u;
  const unsigned lo = __ldg(img + w);
  if (sh == 0)
    return lo;
  const unsigned hi = __ldg(img + w + 1);
  return __funnelshift_r(lo, hi, sh);
}

struct Corner
{
  unsigned x, y, z;  // float bits as stored
};
__device__ __forceinline__ Corner loadCorner(const unsigned* __restrict__ img, unsigned long long corner)
{
  const unsigned long long t = corner / 3;
  const unsigned k = static_cast<unsigned>(corner - 3 * t);
  const unsigned long long at = kStlHeader + kStlFacet * t + 12 + 12 * k;
  return { loadU32(img, at), loadU32(img, at + 4), loadU32(img, at + 8) };
}
/** float equality on finite values: bit equality once -0 is folded onto +0 */
__device__ __forceinline__ unsigned fold(unsigned bits) { return bits == 0x80000000u ? 0u : bits; }
__device__ __forceinline__ bool samePoint(const Corner& a, const Corner& b)
{
  return fold(a.x) == fold(b.x) && fold(a.y) == fold(b.y) && fold(a.z) == fold(b.z);
}
__device__ __forceinline__ unsigned hashPoint(const Corner& c)
{
  unsigned h = fold(c.x) * 0x9e3779b1u;
  h = (h ^ (h >> 15)) + fold(c.y) * 0x85ebca77u;
  h = (h ^ (h >> 13)) + fold(c.z) * 0xc2b2ae3du;
  h ^= h >> 16;
  h *= 0x7feb352du;
  h ^= h >> 15;
  return h;
}

__global__ void __launch_bounds__(256)
    k_stl_insert(const unsigned* __restrict__ img, unsigned long long n_corners, unsigned* __restrict__ slot_rep,
                 unsigned slot_mask, unsigned* __restrict__ corner_slot)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long c = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; c < n_corners; c += stride)
  {
    const Corner me = loadCorner(img, c);
    unsigned s = hashPoint(me) & slot_mask;
    for (;;)
    {
      unsigned rep = slot_rep[s];
      if (rep == kEmpty)
      {
        rep = atomicCAS(slot_rep + s, kEmpty, static_cast<unsigned>(c));
        if (rep == kEmpty)
          break;  // claimed
      }
      // occupied (possibly a moment ago): by this point?  Any representative of a slot has the slot's coordinates.
      if (samePoint(loadCorner(img, rep), me))
      {
        if (c < rep)
          atomicMin(slot_rep + s, static_cast<unsigned>(c));
        break;
      }
      s = (s + 1) & slot_mask;
    }
    corner_slot[c] = s;
  }
}

__global__ void __launch_bounds__(256)
    k_stl_flags(unsigned long long n_corners, const unsigned* __restrict__ slot_rep, const unsigned* __restrict__ corner_slot,
                unsigned* __restrict__ is_first)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long c = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; c <= n_corners; c += stride)
    is_first[c] = (c < n_corners && slot_rep[corner_slot[c]] == static_cast<unsigned>(c)) ? 1u : 0u;
}

/** one thread per facet; corner_slot is overwritten with the corner's point id */
__global__ void __launch_bounds__(256)
    k_stl_points(const unsigned* __restrict__ img, unsigned long long n_facets, const unsigned* __restrict__ slot_rep,
                 unsigned* __restrict__ corner_slot, const unsigned* __restrict__ first_prefix, float4* __restrict__ points,
                 unsigned* __restrict__ keep)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long t = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; t <= n_facets; t += stride)
  {
    if (t == n_facets)
    {
      keep[t] = 0u;  // the scan's total lands here
      continue;
    }
    unsigned id[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
      const unsigned long long c = 3 * t + k;
      const unsigned rep = slot_rep[corner_slot[c]];
      id[k] = first_prefix[rep];
      if (rep == static_cast<unsigned>(c))
      {
        const Corner me = loadCorner(img, c);
        points[id[k]] = make_float4(__uint_as_float(me.x), __uint_as_float(me.y), __uint_as_float(me.z), 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      corner_slot[3 * t + k] = id[k];
    keep[t] = (id[0] != id[1] && id[0] != id[2] && id[1] != id[2]) ? 1u : 0u;
  }
}

__global__ void __launch_bounds__(256)
    k_stl_compact(unsigned long long n_facets, const unsigned* __restrict__ corner_id, const unsigned* __restrict__ keep,
                  const unsigned* __restrict__ keep_prefix, uint32_t* __restrict__ tri)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long t = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n_facets; t += stride)
    if (keep[t])
    {
      const unsigned long long o = 3ull * keep_prefix[t];
      tri[o] = corner_id[3 * t];
      tri[o + 1] = corner_id[3 * t + 1];
      tri[o + 2] = corner_id[3 * t + 2];
    }
}

inline int gridFor(unsigned long long n, int threads, int cap)
{
  unsigned long long b = (n + threads - 1) / threads;
  if (b < 1)
    b = 1;
  if (b > static_cast<unsigned long long>(cap))
    b = cap;
  return static_cast<int>(b);
}
}  // namespace

/** c->fileImg (whole file image, padded) -> c->blob (16-byte records {x y z 0}) and c->tri.  Synchronises once per scan
 *  total (the sizes of the welded mesh are needed on the host).  Returns {points, kept facets}. */
void launchStlDecode(nb2_context* c, uint64_t n_facets, uint64_t* n_points_out, uint64_t* n_kept_out)
{
  const unsigned* img = c->fileImg.as<unsigned>();
  const uint64_t nc = 3 * n_facets;
  uint64_t slots = 1024;
  while (slots < 4 * n_facets)
    slots <<= 1;
  c->stlSlots.ensure(sizeof(unsigned) * slots);
  c->stlCorner.ensure(sizeof(unsigned) * (nc + 1));
  c->stlFlag.ensure(sizeof(unsigned) * (nc + 1));
  c->stlPrefix.ensure(sizeof(unsigned) * (nc + 1));
  NB2_CUDA(cudaMemsetAsync(c->stlSlots.ptr, 0xff, sizeof(unsigned) * slots, c->stream));
  const int cap = c->sm_count * 16;
  k_stl_insert<<<gridFor(nc, 256, cap), 256, 0, c->stream>>>(img, nc, c->stlSlots.as<unsigned>(), static_cast<unsigned>(slots - 1),
                                                               c->stlCorner.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  k_stl_flags<<<gridFor(nc + 1, 256, cap), 256, 0, c->stream>>>(nc, c->stlSlots.as<unsigned>(), c->stlCorner.as<unsigned>(),
                                                                  c->stlFlag.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  launchScanU32(c, c->stlFlag.as<unsigned>(), c->stlPrefix.as<unsigned>(), nc + 1);
  c->hostSmall.ensure(4096);
  unsigned* h = c->hostSmall.as<unsigned>();
  NB2_CUDA(cudaMemcpyAsync(h, c->stlPrefix.as<unsigned>() + nc, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  const uint64_t V = h[0];
  c->blob.ensure(sizeof(float4) * std::max<uint64_t>(V, 1));
  // the flag array is done with: it now holds the per-facet keep flags, the slot array their prefix
  unsigned* keep = c->stlFlag.as<unsigned>();
  c->stlKeepPrefix.ensure(sizeof(unsigned) * (n_facets + 1));
  k_stl_points<<<gridFor(n_facets + 1, 256, cap), 256, 0, c->stream>>>(img, n_facets, c->stlSlots.as<unsigned>(),
                                                                         c->stlCorner.as<unsigned>(), c->stlPrefix.as<unsigned>(),
                                                                         c->blob.as<float4>(), keep);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  launchScanU32(c, keep, c->stlKeepPrefix.as<unsigned>(), n_facets + 1);
  NB2_CUDA(cudaMemcpyAsync(h, c->stlKeepPrefix.as<unsigned>() + n_facets, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  const uint64_t F = h[0];
  c->tri.ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(F, 1));
  if (n_facets)
  {
    k_stl_compact<<<gridFor(n_facets, 256, cap), 256, 0, c->stream>>>(n_facets, c->stlCorner.as<unsigned>(), keep,
                                                                        c->stlKeepPrefix.as<unsigned>(), c->tri.as<uint32_t>());
    NB2_CUDA(cudaGetLastError());
    ++c->launches;
  }
  *n_points_out = V;
  *n_kept_out = F;
}

}  // namespace nb2
