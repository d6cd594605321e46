// Mesh ingest, moments, PCA-frame extents and per-vertex lattice classification (sm_100a).
//
//   K0+K1a  k_ingest_moments : interleaved PCLPointCloud2 blob -> SoA float4 positions / normals, fused with the
//                              single-pass fp64 reduction of sum(x), sum(x x^T) and the AABB
//                              (replaces pcl::fromPCLPointCloud2 x3, pcl::io::mesh2vtk, pcl::PCA's mean/covariance
//                              passes, pcl::computeCentroid: plane_slicer_raster_planner.cpp:536-541,589-590,
//                              principal_axis_direction_generator.cpp:16-19, centroid_origin_generator.cpp:10-13)
//   K1c     k_extents        : float extents of the cloud in the PCA frame (pcl::PCA::project + getMinMax3D,
//                              plane_slicer_raster_planner.cpp:544-548)
//   S1      k_classify       : per vertex, the last lattice plane the vertex is "inside" of (scalar >= 0); replaces
//                              the per-plane vtkPlane::EvaluateFunction sweep over all vertices (:600-608)
#include "device_math.cuh"
#include "frame_dev.cuh"
#include "nb2_internal.h"

#include <cstdlib>

#include <cfloat>

namespace nb2
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double warpSum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warpMin(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fminf(v, __shfl_down_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warpMax(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmaxf(v, __shfl_down_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warpSumU(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(kFull, v, o);
  return v;
}

/** Thread-level partial -> block partial (fixed shuffle tree, then warps summed in index order by thread 0). */
__device__ void blockReducePartial(MomentPartial& p, MomentPartial* smem /*[warps]*/)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    p.s[i] = warpSum(p.s[i]);
    p.mn[i] = warpMin(p.mn[i]);
    p.mx[i] = warpMax(p.mx[i]);
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    p.q[i] = warpSum(p.q[i]);
  p.non_finite = warpSumU(p.non_finite);
  if (lane == 0)
    smem[warp] = p;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    MomentPartial t = smem[0];
    for (int w = 1; w < nwarps; ++w)
    {
      for (int i = 0; i < 3; ++i)
      {
        t.s[i] += smem[w].s[i];
        t.mn[i] = fminf(t.mn[i], smem[w].mn[i]);
        t.mx[i] = fmaxf(t.mx[i], smem[w].mx[i]);
      }
      for (int i = 0; i < 6; ++i)
        t.q[i] += smem[w].q[i];
      t.non_finite += smem[w].non_finite;
    }
    p = t;
  }
}

/** What follows the per-thread accumulation of either ingest kernel: block partial, and in the last block to finish the
 *  fold of the block partials in index order and the PCA frame (framePcaStep) */
__device__ __forceinline__ void momentsEpilogue(MomentPartial& p, MomentPartial* s_part, bool& s_last, MomentPartial* partials,
                                                unsigned* done_counter, FrameDev* __restrict__ frame, const float* piv,
                                                unsigned long long V)
{
  blockReducePartial(p, s_part);
  if (threadIdx.x == 0)
  {
    partials[blockIdx.x] = p;
    __threadfence();
    const unsigned ticket = atomicAdd(done_counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last)
    return;
  __threadfence();

  // final fold by the last block: thread t sums partials t, t + blockDim, ... then the same fixed tree
  MomentPartial t;
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    t.s[i] = 0.0;
    t.mn[i] = FLT_MAX;
    t.mx[i] = -FLT_MAX;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    t.q[i] = 0.0;
  t.non_finite = 0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
  {
    const MomentPartial o = partials[b];
    for (int i = 0; i < 3; ++i)
    {
      t.s[i] += o.s[i];
      t.mn[i] = fminf(t.mn[i], o.mn[i]);
      t.mx[i] = fmaxf(t.mx[i], o.mx[i]);
    }
    for (int i = 0; i < 6; ++i)
      t.q[i] += o.q[i];
    t.non_finite += o.non_finite;
  }
  __syncthreads();
  blockReducePartial(t, s_part);
  if (threadIdx.x == 0)
  {
    partials[gridDim.x] = t;  // slot after the block partials holds the total
    *done_counter = 0;
    // mean, covariance and eigen axes straight from the total: no kernel of its own for one thread of arithmetic
    framedev::framePcaStep(t, piv, V, frame);
  }
}

/**
 * One pass over the interleaved cloud.  Thread i handles vertices i, i + stride, ... (fixed grid => fixed summation
 * order => bit-reproducible moments).  Moments are taken about the first vertex (pivot) to avoid cancellation for
 * meshes far from the origin.  The last block to finish folds the block partials in index order and finalises the
 * PCA frame (framePcaStep).
 */
__global__ void __launch_bounds__(kMomentThreads)
    k_ingest_moments(const uint8_t* __restrict__ blob, unsigned long long V, unsigned point_step, unsigned off_xyz,
                     float4* __restrict__ pos, MomentPartial* partials,
                     unsigned* done_counter, FrameDev* __restrict__ frame)
{
  __shared__ MomentPartial s_part[kMomentThreads / 32];
  __shared__ bool s_last;

  const float* piv = reinterpret_cast<const float*>(blob + off_xyz);
  const double px = piv[0], py = piv[1], pz = piv[2];

  MomentPartial p;
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    p.s[i] = 0.0;
    p.mn[i] = FLT_MAX;
    p.mx[i] = -FLT_MAX;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    p.q[i] = 0.0;
  p.non_finite = 0;

  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  auto accumulate = [&](unsigned long long v, float x, float y, float z) {
    pos[v] = make_float4(x, y, z, 0.f);
    if (!(isfinite(x) && isfinite(y) && isfinite(z)))
    {
      ++p.non_finite;
      return;
    }
    const double dx = static_cast<double>(x) - px, dy = static_cast<double>(y) - py, dz = static_cast<double>(z) - pz;
    p.s[0] += dx;
    p.s[1] += dy;
    p.s[2] += dz;
    p.q[0] += dx * dx;
    p.q[1] += dx * dy;
    p.q[2] += dx * dz;
    p.q[3] += dy * dy;
    p.q[4] += dy * dz;
    p.q[5] += dz * dz;
    p.mn[0] = fminf(p.mn[0], x);
    p.mn[1] = fminf(p.mn[1], y);
    p.mn[2] = fminf(p.mn[2], z);
    p.mx[0] = fmaxf(p.mx[0], x);
    p.mx[1] = fmaxf(p.mx[1], y);
    p.mx[2] = fmaxf(p.mx[2], z);
  };
  // four vertices per step: their twelve loads are in flight together; they are accumulated in the order the one-at-a-
  // time loop would have used, so the sums do not depend on the unrolling
  unsigned long long v = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; v + 3 * stride < V; v += 4 * stride)
  {
    float c[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const float* f = reinterpret_cast<const float*>(blob + (v + u * stride) * point_step + off_xyz);
      c[u][0] = __ldg(f), c[u][1] = __ldg(f + 1), c[u][2] = __ldg(f + 2);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      accumulate(v + u * stride, c[u][0], c[u][1], c[u][2]);
  }
  for (; v < V; v += stride)
  {
    const float* f = reinterpret_cast<const float*>(blob + v * point_step + off_xyz);
    accumulate(v, __ldg(f), __ldg(f + 1), __ldg(f + 2));
  }

  momentsEpilogue(p, s_part, s_last, partials, done_counter, frame, piv, V);
}

// ---------------------------------------------------------------------------------------------------------------------
// K0 with bulk asynchronous copies.  The cloud is one contiguous run of records, so the pass above is pure streaming; its
// rate is set by how many bytes a CTA keeps in flight.  Here one thread per CTA keeps kIngestStages tiles of records on
// their way into shared memory with cp.async.bulk (the TMA engine, SASS UBLKCP; completion counted in bytes on an mbarrier
// per stage) while all threads take their coordinates out of the tile that has landed: no registers hold loads in flight,
// and the refill of a stage is issued the moment its last reader has passed the block barrier.
// Tile t belongs to CTA t mod grid and record r of a tile to thread r mod 256, in tile order: a fixed assignment, so the
// moments are bit-reproducible from run to run (the order differs from the kernel above: the sums differ in their last
// bits between the two kernels, not between runs).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kIngestStages = 4;

__device__ __forceinline__ unsigned smemAddr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbarTryWait(unsigned long long* bar, unsigned parity)
{
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smemAddr(bar)), "r"(parity)
               : "memory");
  return ok != 0u;
}
/** global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the mbarrier */
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)),
               "l"(src), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}

__global__ void __launch_bounds__(kMomentThreads)
    k_ingest_moments_tma(const uint8_t* __restrict__ blob, unsigned long long V, unsigned point_step, unsigned off_xyz,
                         unsigned tile_records, float4* __restrict__ pos, MomentPartial* partials, unsigned* done_counter,
                         FrameDev* __restrict__ frame)
{
  extern __shared__ __align__(128) uint8_t s_tiles[];  // kIngestStages tiles of tile_records * point_step bytes
  __shared__ __align__(8) unsigned long long s_full[kIngestStages];
  __shared__ MomentPartial s_part[kMomentThreads / 32];
  __shared__ bool s_last;

  const float* piv = reinterpret_cast<const float*>(blob + off_xyz);
  const double px = piv[0], py = piv[1], pz = piv[2];
  MomentPartial p;
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    p.s[i] = 0.0;
    p.mn[i] = FLT_MAX;
    p.mx[i] = -FLT_MAX;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    p.q[i] = 0.0;
  p.non_finite = 0;
  auto accumulate = [&](unsigned long long v, float x, float y, float z) {
    pos[v] = make_float4(x, y, z, 0.f);
    if (!(isfinite(x) && isfinite(y) && isfinite(z)))
    {
      ++p.non_finite;
      return;
    }
    const double dx = static_cast<double>(x) - px, dy = static_cast<double>(y) - py, dz = static_cast<double>(z) - pz;
    p.s[0] += dx;
    p.s[1] += dy;
    p.s[2] += dz;
    p.q[0] += dx * dx;
    p.q[1] += dx * dy;
    p.q[2] += dx * dz;
    p.q[3] += dy * dy;
    p.q[4] += dy * dz;
    p.q[5] += dz * dz;
    p.mn[0] = fminf(p.mn[0], x);
    p.mn[1] = fminf(p.mn[1], y);
    p.mn[2] = fminf(p.mn[2], z);
    p.mx[0] = fmaxf(p.mx[0], x);
    p.mx[1] = fmaxf(p.mx[1], y);
    p.mx[2] = fmaxf(p.mx[2], z);
  };

  const unsigned tile_bytes = tile_records * point_step;  // a multiple of 1024 (256 records of a multiple of 4 bytes)
  const unsigned long long full_tiles = V / tile_records;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const unsigned long long mine = full_tiles > blockIdx.x ? (full_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;
  if (threadIdx.x == 0)
  {
    for (int s = 0; s < kIngestStages; ++s)
      mbarInit(&s_full[s], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](unsigned long long k) {  // thread 0: tile k of this CTA into stage k % kIngestStages
    const int s = static_cast<int>(k % kIngestStages);
    const unsigned long long t = blockIdx.x + k * gridDim.x;
    mbarExpectTx(&s_full[s], tile_bytes);
    bulkLoad(s_tiles + static_cast<size_t>(s) * tile_bytes, blob + t * tile_bytes, tile_bytes, &s_full[s]);
  };
  if (threadIdx.x == 0)
    for (unsigned long long k = 0; k < mine && k < kIngestStages; ++k)
      issue(k);
  for (unsigned long long k = 0; k < mine; ++k)
  {
    const int s = static_cast<int>(k % kIngestStages);
    const unsigned parity = static_cast<unsigned>((k / kIngestStages) & 1ull);
    while (!mbarTryWait(&s_full[s], parity))
    {
    }
    const unsigned long long v0 = (blockIdx.x + k * gridDim.x) * tile_records;
    const uint8_t* tile = s_tiles + static_cast<size_t>(s) * tile_bytes + off_xyz;
    for (unsigned r = threadIdx.x; r < tile_records; r += kMomentThreads)
    {
      const float* f = reinterpret_cast<const float*>(tile + static_cast<size_t>(r) * point_step);
      accumulate(v0 + r, f[0], f[1], f[2]);
    }
    __syncthreads();  // every reader of stage s has passed
    if (threadIdx.x == 0 && k + kIngestStages < mine)
    {
      // (the refill is written by the asynchronous proxy: order it behind the generic-proxy reads above)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(k + kIngestStages);
    }
  }
  // the records behind the last full tile (fewer than a tile), by the CTA next in line, straight from global memory
  if (blockIdx.x == full_tiles % gridDim.x)
    for (unsigned long long v = full_tiles * tile_records + threadIdx.x; v < V; v += kMomentThreads)
    {
      const float* f = reinterpret_cast<const float*>(blob + v * point_step + off_xyz);
      accumulate(v, __ldg(f), __ldg(f + 1), __ldg(f + 2));
    }
  momentsEpilogue(p, s_part, s_last, partials, done_counter, frame, piv, V);
}

/** Monotone map float -> unsigned (a < b  <=>  orderedBits(a) < orderedBits(b)), and back */
__device__ __forceinline__ unsigned orderedBits(float f)
{
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

/** Float extents in the PCA frame: proj_c = e_c . (x - mean) with Eigen's 3-term order a0 + (a1 + a2); min/max are
 *  order independent so any reduction tree reproduces pcl::getMinMax3D exactly.  With `with_lattice` the last block
 *  goes on to the lattice step of the plan (frame_dev.cuh). */
__global__ void __launch_bounds__(256)
    k_extents(const float4* __restrict__ pos, unsigned long long V, FrameDev* __restrict__ frame,
              unsigned* __restrict__ counters, FrameParams prm, int with_lattice)
{
  __shared__ float s_mn[3][8], s_mx[3][8];
  __shared__ bool s_last;
  // mean and axes as framePcaStep left them (E column-major: axis c = (E[3c], E[3c+1], E[3c+2]))
  const float mx_ = frame->pca.mean[0], my_ = frame->pca.mean[1], mz_ = frame->pca.mean[2];
  const float e00 = frame->pca.evecs[0], e01 = frame->pca.evecs[1], e02 = frame->pca.evecs[2];
  const float e10 = frame->pca.evecs[3], e11 = frame->pca.evecs[4], e12 = frame->pca.evecs[5];
  const float e20 = frame->pca.evecs[6], e21 = frame->pca.evecs[7], e22 = frame->pca.evecs[8];
  unsigned* const out6 = frame->ext;
  float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  auto project = [&](const float4 p) {
    const float dx = __fsub_rn(p.x, mx_), dy = __fsub_rn(p.y, my_), dz = __fsub_rn(p.z, mz_);
    const float p0 = __fadd_rn(__fmul_rn(e00, dx), __fadd_rn(__fmul_rn(e01, dy), __fmul_rn(e02, dz)));
    const float p1 = __fadd_rn(__fmul_rn(e10, dx), __fadd_rn(__fmul_rn(e11, dy), __fmul_rn(e12, dz)));
    const float p2 = __fadd_rn(__fmul_rn(e20, dx), __fadd_rn(__fmul_rn(e21, dy), __fmul_rn(e22, dz)));
    mn[0] = fminf(mn[0], p0);
    mx[0] = fmaxf(mx[0], p0);
    mn[1] = fminf(mn[1], p1);
    mx[1] = fmaxf(mx[1], p1);
    mn[2] = fminf(mn[2], p2);
    mx[2] = fmaxf(mx[2], p2);
  };
  unsigned long long v = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread
  for (; v + 3 * stride < V; v += 4 * stride)
  {
    const float4 a = __ldg(pos + v), b = __ldg(pos + v + stride), c = __ldg(pos + v + 2 * stride),
                 d = __ldg(pos + v + 3 * stride);
    project(a);
    project(b);
    project(c);
    project(d);
  }
  for (; v < V; v += stride)
    project(__ldg(pos + v));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c)
  {
    mn[c] = warpMin(mn[c]);
    mx[c] = warpMax(mx[c]);
    if (lane == 0)
    {
      s_mn[c][warp] = mn[c];
      s_mx[c][warp] = mx[c];
    }
  }
  __syncthreads();
  // min / max are order independent: fold the block result into the six global slots with integer atomics on the
  // order-preserving bit pattern of the floats
  if (threadIdx.x < 6)
  {
    const int c = threadIdx.x % 3;
    const bool is_max = threadIdx.x >= 3;
    float a = is_max ? s_mx[c][0] : s_mn[c][0];
    for (int w = 1; w < (blockDim.x >> 5); ++w)
      a = is_max ? fmaxf(a, s_mx[c][w]) : fminf(a, s_mn[c][w]);
    if (is_max)
      atomicMax(out6 + threadIdx.x, orderedBits(a));
    else
      atomicMin(out6 + threadIdx.x, orderedBits(a));
  }
  if (!with_lattice)
    return;
  // the plan's lattice follows from the extents: the last block to finish builds it (and puts the slicing counters
  // into their start-of-plan state) instead of a single-thread kernel of its own
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    s_last = atomicAdd(counters + 1, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0)
    return;
  __threadfence();
  counters[1] = 0;
  framedev::resetPlanCounters(counters);
  framedev::frameLatticeStep(frame, prm);
}

/**
 * K_v = the largest lattice index k in [0, P) whose plane scalar at vertex v is >= 0, or -1.  The scalar
 * n . (x - (start + (k s) n)) evaluated with the reference's operation order is non-increasing in k (each rounding is
 * monotone), so "inside" is a prefix of the planes and one integer per vertex encodes the whole per-plane sweep of
 * vtkCutter.  A triangle is cut by plane k  <=>  min K < k <= max K over its vertices, with case bit i = (k <= K_i).
 */
__global__ void __launch_bounds__(256) k_classify(const float4* __restrict__ pos, unsigned long long V,
                                                  const FrameDev* __restrict__ frame, int* __restrict__ vertK,
                                                  unsigned* __restrict__ planeCnt)
{
  const LatticeDev L = frame->L;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  const long long P = L.num_planes;
  // zero the per-plane hit counters of this slab for k_count_hits
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i <= L.plane_end - L.plane_begin;
       i += static_cast<long long>(stride))
    planeCnt[i] = 0u;
  const double inv_spacing = 1.0 / L.spacing;
  const double start_norm = fabs(L.start[0]) + fabs(L.start[1]) + fabs(L.start[2]);
  // (the estimate takes the lattice normal for a unit vector; an explicit lattice may carry another: exact path then)
  const bool unit_normal = fabs(L.n[0] * L.n[0] + L.n[1] * L.n[1] + L.n[2] * L.n[2] - 1.0) * static_cast<double>(P > 0 ? P : 1) < 2.5e-7;
  for (unsigned long long v = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += stride)
  {
    const float4 p = __ldg(pos + v);
    // estimate
    const double d = L.n[0] * (static_cast<double>(p.x) - L.start[0]) + L.n[1] * (static_cast<double>(p.y) - L.start[1]) +
                     L.n[2] * (static_cast<double>(p.z) - L.start[2]);
    const double t = d * inv_spacing;
    double q = floor(t);  // an estimate only: the loops below settle the exact index
    // The estimate and the reference's expression (the scalar against the origin of plane k, start + k * spacing * n) differ
    // by rounding only: a few 2^-53 of (|p| + |start|) / spacing in units of the plane index (and k * spacing * (n.n - 1)).
    // A vertex whose t lies at least `margin` away from both neighbouring integers is therefore settled by the estimate;
    // the two exact evaluations below — most of this kernel's fp64 work — are needed by about two vertices in a million.
    const double scale = (fabs(static_cast<double>(p.x)) + fabs(static_cast<double>(p.y)) + fabs(static_cast<double>(p.z)) + start_norm) * inv_spacing;
    const double margin = 1e-6;
    if (unit_normal && scale < 1e8 && q >= 0.0 && q < static_cast<double>(P - 1) && t - q > margin && (q + 1.0) - t > margin)
    {
      vertK[v] = static_cast<int>(q);
      continue;
    }
    q = fmin(fmax(q, -1.0), static_cast<double>(P - 1));
    long long k = static_cast<long long>(q);
    // exact fix-up against the reference's own expression
    while (k + 1 < P && planeScalar(L, planeOrigin(L, k + 1), p.x, p.y, p.z) >= 0.0)
      ++k;
    while (k >= 0 && !(planeScalar(L, planeOrigin(L, k), p.x, p.y, p.z) >= 0.0))
      --k;
    vertK[v] = static_cast<int>(k);
  }
}

__global__ void __launch_bounds__(256)
    k_set_normals(const float* __restrict__ packed, unsigned long long V, float4* __restrict__ nrm)
{
  const unsigned long long v = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v < V)
    nrm[v] = make_float4(packed[3 * v], packed[3 * v + 1], packed[3 * v + 2], 0.f);
}

/** three floats per vertex at src + v * stride -> packed float[3 V] */
__global__ void __launch_bounds__(256)
    k_pack_out(const uint8_t* __restrict__ src, unsigned stride, unsigned long long V, float* __restrict__ dst)
{
  const unsigned long long v = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v < V)
  {
    const float* a = reinterpret_cast<const float*>(src + v * stride);
    dst[3 * v] = a[0];
    dst[3 * v + 1] = a[1];
    dst[3 * v + 2] = a[2];
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

void launchIngestMoments(nb2_context* c, uint32_t point_step, uint32_t off_xyz, int32_t off_nrm)
{
  const int grid = gridFor(c->V, kMomentThreads, kMomentBlocks);
  c->partials.ensure(sizeof(MomentPartial) * (kMomentBlocks + 1));
  c->frame.ensure(sizeof(FrameDev));
  // normals stay where they are: the slicer gathers the few it needs straight from the cloud (off_nrm is recorded by
  // the caller in nrm_base / nrm_stride)
  c->rec_step = point_step;
  c->rec_off_xyz = off_xyz;
  c->rec_off_nrm = off_nrm;
  // the positions written here are read again by every later pass of the plan: keep as much of the array as the
  // set-aside part of L2 holds (the window covers the array; hitRatio thins it out to what fits)
  if (c->l2_persist_max)
  {
    cudaStreamAttrValue attr{};
    const size_t bytes = sizeof(float4) * c->V;
    attr.accessPolicyWindow.base_ptr = c->pos.ptr;
    const size_t window = std::min(bytes, c->l2_window_max);
    attr.accessPolicyWindow.num_bytes = window;
    attr.accessPolicyWindow.hitRatio = window <= c->l2_persist_max ? 1.0f : static_cast<float>(c->l2_persist_max) / static_cast<float>(window);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess)
      cudaGetLastError();
  }
  // bulk-copy staging (k_ingest_moments_tma) when asked for, the records can travel as 16-byte-aligned tiles and there are
  // enough of them to fill the pipeline; the plain kernel otherwise
  // Measured on cfg3 (profiles/r02_summary.md): 60 us either way — the pass already streams at 4.6 TB/s under ncu and the
  // rest is its serial epilogue, so loads in flight are not what bounds it.  The bulk-copy variant is therefore opt-in
  // (NB2_INGEST_TMA=1); the plain kernel's summation order, which does not depend on the record layout, stays the default.
  static const bool tma_on = [] {
    const char* e = std::getenv("NB2_INGEST_TMA");
    return e && std::atoi(e) != 0;
  }();
  const unsigned tile_records = 2 * kMomentThreads;  // (fixed: the assignment of records to threads must not depend on the layout)
  const size_t tile_bytes = static_cast<size_t>(tile_records) * point_step;
  const bool tma = tma_on && point_step % 4 == 0 && (reinterpret_cast<uintptr_t>(c->blob.ptr) & 15u) == 0 &&
                   c->V >= 64ull * tile_records && kIngestStages * tile_bytes <= (size_t(96) << 10);
  if (tma)
  {
    const size_t smem = kIngestStages * tile_bytes;
    if (c->ingest_smem < smem)
    {
      NB2_CUDA(cudaFuncSetAttribute(k_ingest_moments_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      c->ingest_smem = smem;
    }
    const unsigned long long tiles = c->V / tile_records;
    const int g = static_cast<int>(std::min<unsigned long long>(std::max<unsigned long long>(tiles, 1ull), static_cast<unsigned long long>(kMomentBlocks)));
    k_ingest_moments_tma<<<g, kMomentThreads, smem, c->stream>>>(c->blob.as<uint8_t>(), c->V, point_step, off_xyz, tile_records,
                                                                 c->pos.as<float4>(), c->partials.as<MomentPartial>(),
                                                                 c->counters.as<unsigned>(), c->frame.as<FrameDev>());
    c->ingest_grid = g;
  }
  else
  {
    k_ingest_moments<<<grid, kMomentThreads, 0, c->stream>>>(c->blob.as<uint8_t>(), c->V, point_step, off_xyz,
                                                             c->pos.as<float4>(),
                                                             c->partials.as<MomentPartial>(), c->counters.as<unsigned>(),
                                                             c->frame.as<FrameDev>());
    c->ingest_grid = grid;
  }
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchExtents(nb2_context* c, const FrameParams* lattice_prm)
{
  const int grid = gridFor(c->V, 256, c->sm_count * 8);
  k_extents<<<grid, 256, 0, c->stream>>>(c->pos.as<float4>(), c->V, c->frame.as<FrameDev>(), c->counters.as<unsigned>(),
                                         lattice_prm ? *lattice_prm : FrameParams{}, lattice_prm ? 1 : 0);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchClassify(nb2_context* c)
{
  c->vertK.ensure(sizeof(int) * c->V);
  c->planeCnt.ensure(sizeof(unsigned) * (static_cast<size_t>(c->plane_cap) + 1));
  const int grid = gridFor(c->V, 256, c->sm_count * 16);
  k_classify<<<grid, 256, 0, c->stream>>>(c->pos.as<float4>(), c->V, c->frame.as<FrameDev>(), c->vertK.as<int>(),
                                          c->planeCnt.as<unsigned>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchSetNormals(nb2_context* c, const float* dev_packed)
{
  const int grid = static_cast<int>((c->V + 255) / 256);
  c->nrm.ensure(sizeof(float4) * c->V);
  k_set_normals<<<grid, 256, 0, c->stream>>>(dev_packed, c->V, c->nrm.as<float4>());
  c->nrm_base = c->nrm.as<uint8_t>();
  c->nrm_stride = 16;
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchPackOut(nb2_context* c, float* xyz_packed, float* nrm_packed)
{
  const int grid = static_cast<int>((c->V + 255) / 256);
  if (xyz_packed)
  {
    k_pack_out<<<grid, 256, 0, c->stream>>>(c->pos.as<uint8_t>(), 16u, c->V, xyz_packed);
    ++c->launches;
  }
  if (nrm_packed && !c->nrm_base)
    NB2_CUDA(cudaMemsetAsync(nrm_packed, 0, sizeof(float) * 3 * c->V, c->stream));  // a mesh without normals
  else if (nrm_packed)
  {
    k_pack_out<<<grid, 256, 0, c->stream>>>(c->nrm_base, c->nrm_stride, c->V, nrm_packed);
    ++c->launches;
  }
  NB2_CUDA(cudaGetLastError());
}

}  // namespace nb2
