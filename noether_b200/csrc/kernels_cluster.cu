// Euclidean clustering of the resident cloud on the device: the connected components that
// pcl::EuclideanClusterExtraction finds with one k-d tree radius search per point (euclidean_clustering_modifier.cpp:40-64),
// computed as
//
//   C1  k_clu_insert   one thread per point: claim / find the hash slot of its grid cell, push the point on the cell's chain
//   C2  k_clu_link     one thread per point: the points with a smaller index in the 27 surrounding cells, FLANN's float
//                      distance test, lock-free union (larger root under smaller root)
//   C3  k_clu_flatten  label = root = smallest point index of the component = PCL's seed of that cluster
//
// cluster_ops.h holds the per-point work (shared with the host emulation of the CPU tests).  Sizes, the size filter, PCL's
// reverse std::sort by size and the sub-mesh extraction are index bookkeeping on the labels and stay on the host side of
// the C ABI (plugin shim / Python mirror), where the meshes have to be materialised anyway.
#include "nb2_internal.h"
#include "cluster_ops.h"
#include "pcl_normal_ops.h"

#include <algorithm>

namespace nb2
{
namespace
{
using namespace clu;

struct DevCas64
{
  __device__ unsigned long long operator()(unsigned long long* addr, unsigned long long expect, unsigned long long val) const
  {
    const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(addr);
    if (seen != expect)
      return seen;  // a cell's key never changes once written
    return atomicCAS(addr, expect, val);
  }
};
struct DevCas32
{
  __device__ unsigned operator()(unsigned* addr, unsigned expect, unsigned val) const { return atomicCAS(addr, expect, val); }
};
struct DevXchg
{
  __device__ unsigned operator()(unsigned* addr, unsigned val) const { return atomicExch(addr, val); }
};
struct DevPos
{
  const float4* pos;
  __device__ void operator()(unsigned i, float& x, float& y, float& z) const
  {
    const float4 p = __ldg(pos + i);
    x = p.x;
    y = p.y;
    z = p.z;
  }
};

__global__ void __launch_bounds__(256)
    k_clu_insert(GridParams g, const float4* __restrict__ pos, unsigned V, unsigned long long* cell_key, unsigned* cell_head,
                 unsigned slot_mask, unsigned* __restrict__ next, unsigned* __restrict__ parent, unsigned* occupied)
{
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < V; p += stride)
  {
    const float4 q = __ldg(pos + p);
    if (insertPoint(g, p, q.x, q.y, q.z, cell_key, cell_head, slot_mask, next, parent, DevCas64{}, DevXchg{}))
      atomicAdd(occupied, 1u);
  }
}

__global__ void __launch_bounds__(256)
    k_clu_link(GridParams g, const float4* __restrict__ pos, unsigned V, const unsigned long long* __restrict__ cell_key,
               const unsigned* __restrict__ cell_head, unsigned slot_mask, const unsigned* __restrict__ next, unsigned* parent)
{
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < V; p += stride)
    linkPoint(g, p, DevPos{ pos }, cell_key, cell_head, slot_mask, next, parent, DevCas32{});
}

/** points per occupied cell slot (C1b) */
__global__ void __launch_bounds__(256)
    k_clu_cell_sizes(GridParams g, const float4* __restrict__ pos, unsigned V, const unsigned long long* __restrict__ cell_key,
                     unsigned slot_mask, unsigned* __restrict__ cell_size)
{
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < V; p += stride)
  {
    const float4 q = __ldg(pos + p);
    const unsigned s = findCellSlot(g, cellCoord(g, 0, q.x), cellCoord(g, 1, q.y), cellCoord(g, 2, q.z), cell_key, slot_mask);
    if (s != kNone)
      atomicAdd(cell_size + s, 1u);
  }
}

/** The work k_clu_link is about to do, exactly (C1c): every point walks the chains of its 27 surrounding cells, so the number
 *  of chain steps is the sum over the points of the sizes of those cells — whatever the distribution of the points over the
 *  cells.  work[0] = that sum, work[1] = the longest walk of a single point (one thread does it serially). */
__global__ void __launch_bounds__(256)
    k_clu_work(GridParams g, const float4* __restrict__ pos, unsigned V, const unsigned long long* __restrict__ cell_key,
               unsigned slot_mask, const unsigned* __restrict__ cell_size, unsigned long long* __restrict__ work)
{
  __shared__ unsigned long long s_sum[8], s_max[8];
  unsigned long long sum = 0, mx = 0;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < V; p += stride)
  {
    const float4 q = __ldg(pos + p);
    const long long c[3] = { cellCoord(g, 0, q.x), cellCoord(g, 1, q.y), cellCoord(g, 2, q.z) };
    unsigned long long mine = 0;
    for (long long dz = -1; dz <= 1; ++dz)
      for (long long dy = -1; dy <= 1; ++dy)
        for (long long dx = -1; dx <= 1; ++dx)
        {
          const unsigned s = findCellSlot(g, c[0] + dx, c[1] + dy, c[2] + dz, cell_key, slot_mask);
          if (s != kNone)
            mine += cell_size[s];
        }
    sum += mine;
    mx = mine > mx ? mine : mx;
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    sum += __shfl_down_sync(0xffffffffu, sum, o);
    const unsigned long long other = __shfl_down_sync(0xffffffffu, mx, o);
    mx = other > mx ? other : mx;
  }
  if ((threadIdx.x & 31) == 0)
  {
    s_sum[threadIdx.x >> 5] = sum;
    s_max[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    for (int w = 1; w < 8; ++w)
    {
      sum += s_sum[w];
      mx = s_max[w] > mx ? s_max[w] : mx;
    }
    atomicAdd(work, sum);
    atomicMax(work + 1, mx);
  }
}

__global__ void __launch_bounds__(256) k_clu_flatten(unsigned* parent, unsigned V, unsigned* __restrict__ label)
{
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < V; p += stride)
    label[p] = findRoot(parent, p);
}

/** NormalEstimationPCL (normal_estimation_pcl.cpp:15-50), one thread per point: the radius search over the 27 surrounding
 *  cells (FLANN's float distance test, the query point included), the nine shifted sums in double, pcl_normal_ops.h for the
 *  rest.  Fewer than three neighbours: NaN, as computePointNormal returns false. */
__global__ void __launch_bounds__(256)
    k_pcl_normals(GridParams g, const float4* __restrict__ pos, unsigned V, const unsigned long long* __restrict__ cell_key,
                  const unsigned* __restrict__ cell_head, unsigned slot_mask, const unsigned* __restrict__ next, float vx, float vy,
                  float vz, float4* __restrict__ nrm, float* __restrict__ curvature)
{
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < V; p += stride)
  {
    const float4 P = __ldg(pos + p);
    const long long c[3] = { cellCoord(g, 0, P.x), cellCoord(g, 1, P.y), cellCoord(g, 2, P.z) };
    double S[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    unsigned k = 0;
    for (long long dz = -1; dz <= 1; ++dz)
      for (long long dy = -1; dy <= 1; ++dy)
        for (long long dx = -1; dx <= 1; ++dx)
        {
          const unsigned s = findCellSlot(g, c[0] + dx, c[1] + dy, c[2] + dz, cell_key, slot_mask);
          if (s == kNone)
            continue;
          for (unsigned q = cell_head[s]; q != kNone; q = next[q])
          {
            const float4 Q = __ldg(pos + q);
            if (!isNeighbour(P.x, P.y, P.z, Q.x, Q.y, Q.z, g.r2))
              continue;
            // (float differences as computeMeanAndCovarianceMatrix forms them; their products are exact in double)
            const double x = static_cast<double>(Q.x - P.x), y = static_cast<double>(Q.y - P.y), z = static_cast<double>(Q.z - P.z);
            S[0] += x * x;
            S[1] += x * y;
            S[2] += x * z;
            S[3] += y * y;
            S[4] += y * z;
            S[5] += z * z;
            S[6] += x;
            S[7] += y;
            S[8] += z;
            ++k;
          }
        }
    float n[3] = { nanf(""), nanf(""), nanf("") }, curv = nanf("");
    if (k >= 3u)
    {
      const float pp[3] = { P.x, P.y, P.z }, vp[3] = { vx, vy, vz };
      pclne::normalFromSums(S, k, pp, vp, n, curv);
    }
    nrm[p] = make_float4(n[0], n[1], n[2], 0.f);
    curvature[p] = curv;
  }
}

inline int gridFor(unsigned long long n, const nb2_context* c)
{
  const unsigned long long cap = static_cast<unsigned long long>(c->sm_count) * 16;
  return static_cast<int>(std::max<unsigned long long>(1, std::min((n + 255) / 256, cap)));
}
}  // namespace

/** labels of the resident cloud into `labels_host` (uint32[V]); `max_tests` bounds the chain steps of k_clu_link, counted
 *  exactly on the device before that kernel is launched (a skewed cloud — most points in a few cells — cannot slip under an
 *  average-occupancy estimate and turn the link pass quadratic) */
/** C1 .. C1c: the cell table with every point chained to its cell, and the refusal of radii that are large against the point
 *  spacing (`what` names the caller in the message).  Returns the number of slots of the table. */
static uint64_t buildCellGrid(nb2_context* c, const clu::GridParams& g, double max_tests, const char* what)
{
  const uint64_t V = c->V;
  cudaStream_t st = c->stream;
  uint64_t slots = 1024;
  while (slots < 2 * V)
    slots <<= 1;
  c->subKeys.ensure(sizeof(unsigned long long) * slots);  // the subdivision's table buffer doubles as the cell table
  c->cluHead.ensure(sizeof(unsigned) * slots);
  c->cluNext.ensure(sizeof(unsigned) * V);
  c->cluParent.ensure(sizeof(unsigned) * V);
  c->cluLabel.ensure(sizeof(unsigned) * V);
  c->cluCellSize.ensure(sizeof(unsigned) * slots + 2 * sizeof(unsigned long long));
  unsigned long long* work = reinterpret_cast<unsigned long long*>(c->cluCellSize.as<unsigned>() + slots);
  unsigned* occupied = c->counters.as<unsigned>() + 21;
  NB2_CUDA(cudaMemsetAsync(c->subKeys.ptr, 0xff, sizeof(unsigned long long) * slots, st));
  NB2_CUDA(cudaMemsetAsync(c->cluHead.ptr, 0xff, sizeof(unsigned) * slots, st));
  NB2_CUDA(cudaMemsetAsync(occupied, 0, sizeof(unsigned), st));
  NB2_CUDA(cudaMemsetAsync(c->cluCellSize.ptr, 0, sizeof(unsigned) * slots + 2 * sizeof(unsigned long long), st));
  const unsigned n = static_cast<unsigned>(V);
  k_clu_insert<<<gridFor(V, c), 256, 0, st>>>(g, c->pos.as<float4>(), n, c->subKeys.as<unsigned long long>(), c->cluHead.as<unsigned>(),
                                             static_cast<unsigned>(slots - 1), c->cluNext.as<unsigned>(), c->cluParent.as<unsigned>(),
                                             occupied);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  k_clu_cell_sizes<<<gridFor(V, c), 256, 0, st>>>(g, c->pos.as<float4>(), n, c->subKeys.as<unsigned long long>(),
                                                 static_cast<unsigned>(slots - 1), c->cluCellSize.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  k_clu_work<<<gridFor(V, c), 256, 0, st>>>(g, c->pos.as<float4>(), n, c->subKeys.as<unsigned long long>(),
                                           static_cast<unsigned>(slots - 1), c->cluCellSize.as<unsigned>(), work);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  c->hostSmall.ensure(4096);
  unsigned* h = c->hostSmall.as<unsigned>();
  unsigned long long* hw = reinterpret_cast<unsigned long long*>(h + 2);
  NB2_CUDA(cudaMemcpyAsync(h, occupied, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  NB2_CUDA(cudaMemcpyAsync(hw, work, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  NB2_CUDA(cudaStreamSynchronize(st));
  // one thread walks the chains of one point: beyond about a million steps that single walk takes a second
  constexpr unsigned long long kMaxStepsPerPoint = 1ull << 20;
  if (static_cast<double>(hw[0]) > max_tests || hw[1] > kMaxStepsPerPoint)
    fail(NB2_ERR_INVALID_ARGUMENT, std::string(what) + ": the search radius is large against the point spacing (" + std::to_string(V) +
                                       " points in " + std::to_string(h[0]) + " grid cells, " + std::to_string(hw[0]) +
                                       " distance tests in all and " + std::to_string(hw[1]) +
                                       " for the busiest point); every radius search of the reference would return a large part of the cloud too");
  return slots;
}

/** labels of the resident cloud into `labels_host` (uint32[V]) */
void launchClusterLabels(nb2_context* c, const clu::GridParams& g, double max_tests, uint32_t* labels_host)
{
  const uint64_t V = c->V;
  cudaStream_t st = c->stream;
  const unsigned n = static_cast<unsigned>(V);
  const uint64_t slots = buildCellGrid(c, g, max_tests, "EuclideanClustering");
  k_clu_link<<<gridFor(V, c), 256, 0, st>>>(g, c->pos.as<float4>(), n, c->subKeys.as<unsigned long long>(), c->cluHead.as<unsigned>(),
                                           static_cast<unsigned>(slots - 1), c->cluNext.as<unsigned>(), c->cluParent.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  k_clu_flatten<<<gridFor(V, c), 256, 0, st>>>(c->cluParent.as<unsigned>(), n, c->cluLabel.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  NB2_CUDA(cudaMemcpyAsync(labels_host, c->cluLabel.ptr, sizeof(unsigned) * V, cudaMemcpyDeviceToHost, st));
  NB2_CUDA(cudaStreamSynchronize(st));
}

/** NormalEstimationPCL on the resident cloud: the estimated normals become the resident normals (float4 SoA, as after
 *  NormalsFromMeshFaces); packed copies go to the host (either may be NULL) */
void launchNormalEstimationPcl(nb2_context* c, const clu::GridParams& g, double max_tests, const float viewpoint[3], float* normals_host,
                               float* curvature_host)
{
  const uint64_t V = c->V;
  cudaStream_t st = c->stream;
  const unsigned n = static_cast<unsigned>(V);
  c->timer.begin(st);
  const uint64_t slots = buildCellGrid(c, g, max_tests, "NormalEstimationPCL");
  c->timer.mark("cell grid", st);
  c->nrm.ensure(sizeof(float4) * V);
  c->cluLabel.ensure(sizeof(float) * V);  // (the clustering's label buffer holds the curvature)
  k_pcl_normals<<<gridFor(V, c), 256, 0, st>>>(g, c->pos.as<float4>(), n, c->subKeys.as<unsigned long long>(), c->cluHead.as<unsigned>(),
                                              static_cast<unsigned>(slots - 1), c->cluNext.as<unsigned>(), viewpoint[0], viewpoint[1],
                                              viewpoint[2], c->nrm.as<float4>(), c->cluLabel.as<float>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
  c->timer.mark("radius search + normals", st);
  c->nrm_base = c->nrm.as<uint8_t>();  // the estimated normals replace whatever the cloud carried
  c->nrm_stride = 16;
  c->has_normals = true;
  if (normals_host)
  {
    c->outNrm.ensure(sizeof(float) * 3 * V);
    launchPackOut(c, nullptr, c->outNrm.as<float>());
    NB2_CUDA(cudaMemcpyAsync(normals_host, c->outNrm.ptr, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
  }
  if (curvature_host)
    NB2_CUDA(cudaMemcpyAsync(curvature_host, c->cluLabel.ptr, sizeof(float) * V, cudaMemcpyDeviceToHost, st));
  c->timer.mark("d2h", st);
  NB2_CUDA(cudaStreamSynchronize(st));
}
}  // namespace nb2
