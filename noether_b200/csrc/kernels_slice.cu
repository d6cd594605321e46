// The vtkCutter + vtkStripper + processPolyLines replacement (sm_100a).
//
// The reference loops over planes and lets vtkCutter visit every vertex and every triangle per plane
// (plane_slicer_raster_planner.cpp:596-628).  Here the loop is inverted:
//
//   S2  k_count_hits   : each triangle projects onto the plane lattice through the per-vertex indices K_v
//                        (k_classify) and bumps a per-plane hit counter; vertex indices are range-checked on the way
//   S3  (last block of S2) : exclusive scan of the per-plane counters -> bucket offsets (+ total hits, largest plane)
//   S4  k_emit_hits    : the triangles are streamed a second time; every (triangle, plane) hit is contoured exactly as
//                        vtkTriangle::Contour does, where the triangle's vertices are hot in cache, and its two cut
//                        points (float32 position + first-inserter rank) are appended to the plane's bucket
//   S5  k_link_planes  : one persistent CTA per plane, shared memory only: merge cut points exactly as vtkMergePoints
//                        does (float-coordinate equality, shared-memory hash), pair the lines at shared points, rank
//                        the chains (Helman-JaJa list ranking), orient and order them, and write, per output
//                        waypoint, the rank (triangle, line end) of the cut point that inserted it first — into a
//                        window of the output that belongs to the plane alone, so no plane waits for another one
//   S5b (last CTA of S5)   : exclusive prefix of the per-plane waypoint / segment counts
//   S6  k_waypoints    : one thread per output waypoint: re-contour the first inserter, interpolate the vertex normals
//                        along the cut edge, write packed position / normal / generating mesh edge
//
// Every kernel from S2 on reads the lattice, the plane count and the hit count from device memory, so a plan is enqueued
// without the host waiting for any of them.
//
// Output of S6 is the reference's planImpl result before convertToPoses expands it to 4x4 matrices.
#include "device_math.cuh"
#include "link2_ops.h"
#include "nb2_internal.h"

#include <cfloat>
#include <cstdlib>

namespace nb2
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;
constexpr int kCountThreads = 256;
constexpr int kSmemBins = 8192;  // per-block privatised plane histogram (32 KB)

/** Four consecutive triangles per thread: three 16-byte loads when the index array is 16-byte aligned. */
__device__ __forceinline__ int loadQuad(const uint32_t* __restrict__ tri, unsigned long long F, unsigned long long q,
                                        bool aligned, uint32_t id[12])
{
  const unsigned long long t0 = 4ull * q;
  if (aligned && t0 + 4 <= F)
  {
    const uint4* p = reinterpret_cast<const uint4*>(tri) + 3ull * q;
    const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    id[0] = a.x, id[1] = a.y, id[2] = a.z, id[3] = a.w;
    id[4] = b.x, id[5] = b.y, id[6] = b.z, id[7] = b.w;
    id[8] = c.x, id[9] = c.y, id[10] = c.z, id[11] = c.w;
    return 4;
  }
  const int n = static_cast<int>(F - t0 < 4ull ? F - t0 : 4ull);
#pragma unroll
  for (int i = 0; i < 12; ++i)
    id[i] = (i < 3 * n) ? __ldg(tri + 3ull * t0 + i) : 0u;
  return n;
}

/** Lattice planes [k0, k1] (clamped to the slab) that cut a triangle whose vertices classify as ka, kb, kc */
__device__ __forceinline__ void planeSpan(int ka, int kb, int kc, const LatticeDev& L, long long& k0, long long& k1)
{
  const int lo = min(ka, min(kb, kc)), hi = max(ka, max(kb, kc));
  k0 = max(static_cast<long long>(lo) + 1, L.plane_begin);
  k1 = min(static_cast<long long>(hi), L.plane_end - 1);
}

/** Block-wide exclusive scan of the per-plane counters (any block size that is a multiple of 32; the counters are read
 *  past L1, they were built with atomics by other blocks).  out[0] = total hits, out[1] = largest plane. */
__device__ void scanPlanesBlock(const unsigned* cnt, long long nP, unsigned* __restrict__ off,
                                unsigned* __restrict__ cursor, unsigned* __restrict__ out2)
{
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_carry;
  __shared__ unsigned s_max[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    s_carry = 0;
  unsigned vmax = 0;
  __syncthreads();
  for (long long base = 0; base < nP; base += blockDim.x)
  {
    const long long i = base + threadIdx.x;
    const unsigned v = (i < nP) ? __ldcg(cnt + i) : 0u;
    vmax = max(vmax, v);
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const unsigned y = __shfl_up_sync(kFull, x, o);
      if (lane >= o)
        x += y;
    }
    if (lane == 31)
      s_warp[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
      unsigned w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const unsigned y = __shfl_up_sync(kFull, w, o);
        if (lane >= o)
          w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const unsigned carry = s_carry;
    const unsigned incl = x + (warp ? s_warp[warp - 1] : 0u) + carry;
    if (i < nP)
    {
      off[i] = incl - v;
      cursor[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1)
      s_carry = incl;
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    vmax = max(vmax, __shfl_down_sync(kFull, vmax, o));
  if (lane == 0)
    s_max[warp] = vmax;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned m = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w)
      m = max(m, s_max[w]);
    off[nP] = s_carry;
    out2[0] = s_carry;
    out2[1] = m;
  }
}

template <bool kListQuads>
__global__ void __launch_bounds__(kCountThreads)
    k_count_hits(const uint32_t* __restrict__ tri, unsigned long long F, unsigned long long V,
                 const int* __restrict__ vertK, const FrameDev* __restrict__ frame, unsigned* __restrict__ planeCnt,
                 int aligned, unsigned* __restrict__ maxIndex, unsigned* __restrict__ doneTicket,
                 unsigned* __restrict__ planeOff, unsigned* __restrict__ planeCursor, unsigned* __restrict__ sizes,
                 unsigned* __restrict__ quadList, unsigned* __restrict__ quadCount)
{
  __shared__ unsigned s_bins[kSmemBins];
  __shared__ bool s_last;
  const LatticeDev L = frame->L;
  const long long nP = L.plane_end - L.plane_begin;
  if (nP <= 0)
  {
    // nothing to cut: k_classify zeroed planeCnt[0]; the sizes stay at their start-of-plan zeros
    if (blockIdx.x == 0 && threadIdx.x == 0)
      planeOff[0] = 0u;
    return;
  }
  const bool use_smem = nP <= kSmemBins;
  if (use_smem)
  {
    for (int i = threadIdx.x; i < nP; i += blockDim.x)
      s_bins[i] = 0;
    __syncthreads();
  }
  const unsigned long long Q = (F + 3) / 4;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  if constexpr (!kListQuads)
  {
    for (unsigned long long q = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < Q; q += stride)
    {
      uint32_t id[12];
      const int n = loadQuad(tri, F, q, aligned != 0, id);
      // range check of the vertex indices (the reference would read out of bounds; here the plan fails instead)
      uint32_t m = id[0];
#pragma unroll
      for (int i = 1; i < 12; ++i)
        m = max(m, id[i]);
      if (m >= V)
      {
        atomicMax(maxIndex, m);
        continue;
      }
      int K[12];
#pragma unroll
      for (int i = 0; i < 12; ++i)
        K[i] = __ldg(vertK + id[i]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        if (i >= n)
          break;
        long long k0, k1;
        planeSpan(K[3 * i], K[3 * i + 1], K[3 * i + 2], L, k0, k1);
        for (long long k = k0; k <= k1; ++k)
        {
          if (use_smem)
            atomicAdd(&s_bins[k - L.plane_begin], 1u);
          else
            atomicAdd(&planeCnt[k - L.plane_begin], 1u);
        }
      }
    }
  }
  else
  {
  // (the loop is uniform per block when the quads with hits are listed: the warps vote below)
  for (unsigned long long q0 = static_cast<unsigned long long>(blockIdx.x) * blockDim.x; q0 < Q; q0 += stride)
  {
    const unsigned long long q = q0 + threadIdx.x;
    bool any_hit = false;
    if (q < Q)
    {
      uint32_t id[12];
      const int n = loadQuad(tri, F, q, aligned != 0, id);
      // range check of the vertex indices (the reference would read out of bounds; here the plan fails instead)
      uint32_t m = id[0];
#pragma unroll
      for (int i = 1; i < 12; ++i)
        m = max(m, id[i]);
      if (m >= V)
        atomicMax(maxIndex, m);
      else
      {
        int K[12];
#pragma unroll
        for (int i = 0; i < 12; ++i)
          K[i] = __ldg(vertK + id[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
          if (i >= n)
            break;
          long long k0, k1;
          planeSpan(K[3 * i], K[3 * i + 1], K[3 * i + 2], L, k0, k1);
          any_hit |= k1 >= k0;
          for (long long k = k0; k <= k1; ++k)
          {
            if (use_smem)
              atomicAdd(&s_bins[k - L.plane_begin], 1u);
            else
              atomicAdd(&planeCnt[k - L.plane_begin], 1u);
          }
        }
      }
    }
    {
      // A slab of a sharded plan is cut by a fraction of the triangles only: the quads that are go on a list, so that
      // k_emit_hits does not stream the whole triangle array again (one atomic per warp; the order is arbitrary)
      const unsigned vote = __ballot_sync(kFull, any_hit);
      if (vote)
      {
        const unsigned lane = threadIdx.x & 31u;
        unsigned base = 0;
        if (lane == 0)
          base = atomicAdd(quadCount, __popc(vote));
        base = __shfl_sync(kFull, base, 0);
        if (any_hit)
          quadList[base + __popc(vote & ((1u << lane) - 1u))] = static_cast<unsigned>(q);
      }
    }
  }
  }
  if (use_smem)
  {
    __syncthreads();
    for (int i = threadIdx.x; i < nP; i += blockDim.x)
    {
      const unsigned v = s_bins[i];
      if (v)
        atomicAdd(&planeCnt[i], v);
    }
  }
  // S3: the last block to finish turns the counters into bucket offsets (+ total hits, largest plane)
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    s_last = atomicAdd(doneTicket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last)
    return;
  __threadfence();
  if (threadIdx.x == 0)
    *doneTicket = 0u;
  scanPlanesBlock(planeCnt, nP, planeOff, planeCursor, sizes);
}

/**
 * S4.  Contour every (triangle, plane) hit where the triangles are streamed: the three vertices of a triangle are
 * shared with its neighbours in the stream, so the position gathers hit L1/L2.  A hit appends two endpoint records
 * {x, y, z, rank} to its plane's bucket; rank = 2 * triangle + line end is the order in which vtkCutter (cells in id
 * order, vtkTriangle::Contour inserting pts[0] then pts[1]) would have inserted the point.  The order of the hits
 * inside a bucket is arbitrary (integer atomics on the cursors); every later decision is made on ranks, never on bucket
 * positions, so the result does not depend on it.
 *
 * Only a fraction of the triangles is cut by any plane, and contouring is ~200 fp64 instructions, so the block first
 * lists the hits of a tile of triangles in shared memory (block scan of the per-thread hit counts) and then contours
 * them with all lanes busy.
 */
constexpr int kEmitThreads = 256;
constexpr int kEmitList = 2048;  // hits contoured per round of a tile

__global__ void __launch_bounds__(kEmitThreads, 4)
    k_emit_hits(const uint32_t* __restrict__ tri, unsigned long long F, const int* __restrict__ vertK,
                const float4* __restrict__ pos, const FrameDev* __restrict__ frame, const unsigned* __restrict__ planeOff,
                unsigned* __restrict__ cursor, float4* __restrict__ rec, int aligned, const unsigned* __restrict__ hits,
                unsigned long long hitCap, unsigned* __restrict__ overflow, const unsigned* __restrict__ quadList,
                const unsigned* __restrict__ quadCount)
{
  __shared__ uint2 s_list[kEmitList];  // {triangle, slab-local plane}
  // The buffers were sized before the hit count was known (from the previous plan on this context): when this plan
  // has more hits than they hold, nothing is written; the host sees the flag, grows the buffers and repeats the stage.
  if (*hits > hitCap)
  {
    if (blockIdx.x == 0 && threadIdx.x == 0)
      atomicOr(overflow, 1u);
    return;
  }
  const LatticeDev L = frame->L;
  __shared__ unsigned s_warp[kEmitThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // all quads of the triangle stream, or (sharded plans) the quads k_count_hits listed as cut by this slab
  const unsigned long long Q = quadList ? *quadCount : (F + 3) / 4;
  const unsigned long long tiles = (Q + kEmitThreads - 1) / kEmitThreads;
  for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x)
  {
    const unsigned long long slot = tile * kEmitThreads + threadIdx.x;
    const unsigned long long q = (quadList && slot < Q) ? __ldg(quadList + slot) : slot;
    int k0[4];
    unsigned cnt[4] = { 0u, 0u, 0u, 0u };
    if (slot < Q)
    {
      uint32_t id[12];
      const int n = loadQuad(tri, F, q, aligned != 0, id);
      int K[12];
#pragma unroll
      for (int i = 0; i < 12; ++i)
        K[i] = __ldg(vertK + id[i]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        long long a, b;
        planeSpan(K[3 * i], K[3 * i + 1], K[3 * i + 2], L, a, b);
        k0[i] = static_cast<int>(a - L.plane_begin);
        cnt[i] = (i < n && b >= a) ? static_cast<unsigned>(b - a + 1) : 0u;
      }
    }
    // block-wide exclusive scan of the per-thread hit counts
    const unsigned mine = cnt[0] + cnt[1] + cnt[2] + cnt[3];
    unsigned x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const unsigned y = __shfl_up_sync(kFull, x, o);
      if (lane >= o)
        x += y;
    }
    __syncthreads();  // the previous tile's readers of s_warp / s_list are done
    if (lane == 31)
      s_warp[warp] = x;
    __syncthreads();
    unsigned before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kEmitThreads / 32; ++w)
    {
      const unsigned t = s_warp[w];
      before += (w < warp) ? t : 0u;
      total += t;
    }
    const unsigned first = before + x - mine;  // tile-local index of this thread's first hit

    for (unsigned w0 = 0; w0 < total; w0 += kEmitList)
    {
      if (w0)
        __syncthreads();
      // scatter this thread's hits that fall into the window [w0, w0 + kEmitList)
      unsigned at = first;
#pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        if (cnt[i] && at + cnt[i] > w0 && at < w0 + kEmitList)
        {
          const unsigned jb = at < w0 ? w0 - at : 0u;
          const unsigned je = min(cnt[i], w0 + kEmitList - at);
          const unsigned t = static_cast<unsigned>(4ull * q + i);
          for (unsigned j = jb; j < je; ++j)
            s_list[at + j - w0] = make_uint2(t, static_cast<unsigned>(k0[i]) + j);
        }
        at += cnt[i];
      }
      __syncthreads();
      const unsigned m = min(total - w0, static_cast<unsigned>(kEmitList));
      for (unsigned h = threadIdx.x; h < m; h += kEmitThreads)
      {
        const uint2 hit = s_list[h];
        const unsigned t = hit.x;
        const uint32_t ids[3] = { __ldg(tri + 3ull * t), __ldg(tri + 3ull * t + 1), __ldg(tri + 3ull * t + 2) };
        const float4 p[3] = { __ldg(pos + ids[0]), __ldg(pos + ids[1]), __ldg(pos + ids[2]) };
        const unsigned slot = __ldg(planeOff + hit.y) + atomicAdd(&cursor[hit.y], 1u);
        const PlaneOrigin org = planeOrigin(L, L.plane_begin + static_cast<long long>(hit.y));
        double s[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          s[j] = planeScalar(L, org, p[j].x, p[j].y, p[j].z);
        float c0[3], c1[3];
        cutBothPositions(contourCase(s), p, s, c0, c1);
        rec[2ull * slot] = make_float4(c0[0], c0[1], c0[2], __uint_as_float(2u * t));
        rec[2ull * slot + 1] = make_float4(c1[0], c1[1], c1[2], __uint_as_float(2u * t + 1u));
      }
    }
  }
}

// =================================================================================================================
// S5: per-plane linking
// =================================================================================================================
constexpr unsigned kEmpty = 0xffffffffu;

struct Slot
{
  unsigned first;  // endpoint that claimed the slot (key holder); later reused as the pairing cell
  unsigned tag;    // smallest first-inserter rank (2 * triangle + line end) among the merged endpoints
};

struct LinkParams
{
  const float4* rec;  // [2 * H] endpoint records {x, y, z, rank bits}, grouped by plane (k_emit_hits)
  const unsigned* planeOff;
  const FrameDev* frame;  // the lattice lives in device memory (no host round trip before this kernel)
  LatticeDev L;           // ... and is copied here by the kernel
  unsigned* ticket;
  unsigned* planeMeta;  // [2 * nP] waypoints, segments of each plane
  // A plane with n hits yields at most 2 n waypoints and n segments, so plane k writes its waypoint ranks at
  // wpTag[2 * planeOff[k] ...] and its segment lengths at segLen[planeOff[k] ...]: no plane waits for another one.
  // scanMetaBlock + k_waypoints close the gaps.
  unsigned* wpTag;   // [2 * H] rank (2 * triangle + line end) of the first inserter of each waypoint
  unsigned* segLen;  // [H]
  unsigned* status;  // [0] error bits, [1] first failing plane
  uint8_t* scratch;  // global work areas for planes that do not fit shared memory
  unsigned long long scratchStride;
  unsigned smemHitCap;   // largest plane (in hits) that fits the shared-memory work area
  unsigned scratchHits;  // largest plane the global work areas were sized for
  unsigned* overflow;    // bit 0: hit buffers too small (k_emit_hits), bit 1: a plane exceeds scratchHits
  unsigned fastHitCap;   // largest plane (in cut lines) whose second-generation work area fits the CTA's shared memory
  unsigned plainRounds;  // claim rounds with plain stores before the compare-and-swap loop takes what is left (>= 1)
  unsigned* generalCount;  // planes linked by the general path, the specification, because no fast path took them (diagnostics)
  unsigned long long* phaseCycles;  // [16] diagnostics (NB2_LINK_PROFILE), normally null
  unsigned* done;        // CTAs that ran out of planes (the last one scans the per-plane counts)
  uint2* planePrefix;    // [nP + 1] {waypoints, segments} before each plane
  unsigned* totals;      // [2] waypoints, segments of the slab
  // The plan's report, written straight into pinned host memory by the last CTA (no copy engine between the plan and
  // the host): the frame state, counters[4..13] and the leading per-plane counts, then the plan's stamp.
  const unsigned* counters;
  unsigned* reportFrame;   // [sizeof(FrameDev) / 4]
  unsigned* reportSizes;   // [16]: counters[4..13], ..., [15] stamp
  unsigned* reportMeta;    // [2 * reportPlanes]
  unsigned reportPlanes;
  unsigned stamp;
};

constexpr unsigned kErrNonManifold = 1u;  // (bit 1, "closed loop", is reserved: closed contours are walked, not refused)
// pairing cell of a hash slot: [31:20] number of line ends at the point, [19:0] last line end that arrived
constexpr unsigned kCellShift = 20u, kCellMask = (1u << kCellShift) - 1u;
constexpr unsigned kMaxCrowded = 16u;      // points with more than two incident lines handled per plane
constexpr unsigned kMaxCrowdedEnds = 32u;  // line ends at one such point
// s_misc: [0] chains, [1] error bits, [2] splitters, [4] plane waypoints, [5] ranking flag, [6] crowded points,
//         [7] closed loops present, [8..] crowded slot ids
constexpr int kMiscWords = 8 + kMaxCrowded;
constexpr int kLinkThreads = 512;
// list ranking: both darts of every (1 << kSplitShift)-th line of the bucket are splitters (Helman & JaJa)
constexpr unsigned kSplitShift = 4u, kSplitMask = (1u << kSplitShift) - 1u;

__device__ __forceinline__ unsigned hashPoint(float x, float y, float z)
{
  // vtkMergePoints compares with operator== : fold -0 onto +0 before hashing
  const unsigned a = __float_as_uint(x == 0.f ? 0.f : x), b = __float_as_uint(y == 0.f ? 0.f : y),
                 c = __float_as_uint(z == 0.f ? 0.f : z);
  unsigned h = a * 0x9E3779B1u;
  h ^= (b + 0x7F4A7C15u) * 0x85EBCA77u;
  h ^= (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 13;
  return h;
}

/** Work area of one plane; Idx is uint16_t in shared memory (<= 32767 hits) and uint32_t in the global fallback. */
template <typename Idx>
struct PlaneWork
{
  float* epos;      // [3 * E]   endpoint positions (phase A/H); later the list-ranking arrays
  size_t eposBytes; //           size of that region: 12 E with 16-bit indices, 32 E with 32-bit indices
  unsigned* etag;   // [E]       slot index, later the merged point's first-inserter rank
  Idx* twin;        // [E]       endpoint sharing the same cut point in another line
  Slot* slots;      // [cap]     hash of cut points; later list-ranking arrays, then chain heads
  unsigned cap;     // power of two >= E
};

template <typename Idx>
__device__ void linkPlane(const LinkParams& P, const PlaneWork<Idx>& W, long long kLocal, unsigned n, unsigned* s_misc)
{
  constexpr Idx kNone = static_cast<Idx>(~static_cast<Idx>(0));
  const unsigned E = 2 * n;
  const unsigned tid = threadIdx.x, nth = blockDim.x;
  const unsigned off = P.planeOff[kLocal];
  const unsigned mask = W.cap - 1;
  const float4* __restrict__ rec = P.rec + 2ull * off;  // this plane's endpoint records
  // rank (2 * triangle + line end) and float32 position of endpoint e, straight from the records (L2 resident)
  auto endTag = [&](unsigned e) { return __float_as_uint(__ldg(reinterpret_cast<const float*>(rec + e) + 3)); };
  auto endPos = [&](unsigned e) { return __ldg(rec + e); };

  // s_misc: [0] nHeads, [1] local error bits, [2] baseW, [3] baseS, [4] totalW, [5] any-unfinished flag
  if (tid < kMiscWords)
    s_misc[tid] = 0;

  // ---- A: stage the plane's endpoint records (coalesced 16-byte loads) ---------------------------------------------
  for (unsigned e = tid; e < E; e += nth)
  {
    const float4 r = __ldg(rec + e);
    W.epos[3 * e] = r.x;
    W.epos[3 * e + 1] = r.y;
    W.epos[3 * e + 2] = r.z;
    W.etag[e] = __float_as_uint(r.w);
  }
  for (unsigned s = tid; s < W.cap; s += nth)
  {
    W.slots[s].first = kEmpty;
    W.slots[s].tag = kEmpty;
  }
  __syncthreads();

  // ---- H1: merge cut points with equal float coordinates (vtkMergePoints::InsertUniquePoint) -----------------------
  for (unsigned e = tid; e < E; e += nth)
  {
    const float x = W.epos[3 * e], y = W.epos[3 * e + 1], z = W.epos[3 * e + 2];
    const unsigned tag = W.etag[e];
    unsigned s = hashPoint(x, y, z) & mask;
    while (true)
    {
      unsigned f = *reinterpret_cast<volatile unsigned*>(&W.slots[s].first);
      if (f == kEmpty)
      {
        f = atomicCAS(&W.slots[s].first, kEmpty, e);
        if (f == kEmpty)
          break;
      }
      if (W.epos[3 * f] == x && W.epos[3 * f + 1] == y && W.epos[3 * f + 2] == z)
        break;
      s = (s + 1) & mask;
    }
    atomicMin(&W.slots[s].tag, tag);
    W.etag[e] = s;
  }
  __syncthreads();

  // ---- H2: pair the two line ends meeting at each point -------------------------------------------------------------
  // The pairing cell of a slot packs {number of line ends that arrived, last arrival}: the second arrival pairs with
  // the first.  Points with more than two incident lines (a plane through a saddle-like vertex, or a doubled line
  // where the plane grazes a mesh edge) are re-paired below in tag order so the result does not depend on arrival order.
  for (unsigned s = tid; s < W.cap; s += nth)
    W.slots[s].first = 0u;
  for (unsigned e = tid; e < E; e += nth)
    W.twin[e] = kNone;
  __syncthreads();
  for (unsigned e = tid; e < E; e += nth)
  {
    // a line whose two ends merged into one point is dropped (vtkTriangle::Contour: pts[0] != pts[1])
    if (W.etag[e] == W.etag[e ^ 1u])
      continue;
    unsigned* cell = &W.slots[W.etag[e]].first;
    unsigned old = *reinterpret_cast<volatile unsigned*>(cell);
    while (true)
    {
      const unsigned seen = atomicCAS(cell, old, (((old >> kCellShift) + 1u) << kCellShift) | e);
      if (seen == old)
        break;
      old = seen;
    }
    if ((old >> kCellShift) == 1u)
    {
      const unsigned prev = old & kCellMask;
      W.twin[e] = static_cast<Idx>(prev);
      W.twin[prev] = static_cast<Idx>(e);
    }
  }
  __syncthreads();
  for (unsigned e = tid; e < E; e += nth)
  {
    if (W.etag[e] == W.etag[e ^ 1u])
      continue;
    const unsigned cellv = W.slots[W.etag[e]].first;
    if ((cellv >> kCellShift) > 2u)
    {
      W.twin[e] = kNone;
      if ((cellv & kCellMask) == e)  // one registration per crowded point
      {
        const unsigned i = atomicAdd(&s_misc[6], 1u);
        if (i < kMaxCrowded)
          s_misc[8 + i] = W.etag[e];
        else
          atomicOr(&s_misc[1], kErrNonManifold);
      }
    }
  }
  __syncthreads();
  if (s_misc[6] > 0 && !(s_misc[1] & kErrNonManifold))
  {
    // rare: one thread per crowded point gathers its line ends, sorts them by tag and pairs them consecutively
    if (tid < s_misc[6])
    {
      const unsigned slot = s_misc[8 + tid];
      unsigned mem[kMaxCrowdedEnds];
      unsigned cnt = 0;
      bool overflow = false;
      for (unsigned e = 0; e < E; ++e)
      {
        if (W.etag[e] != slot || W.etag[e ^ 1u] == slot)
          continue;
        if (cnt == kMaxCrowdedEnds)
        {
          overflow = true;
          break;
        }
        // insertion sort by line-end tag
        const unsigned tg = endTag(e);
        unsigned j = cnt++;
        while (j > 0)
        {
          const unsigned o = mem[j - 1];
          const unsigned to = endTag(o);
          if (to < tg)
            break;
          mem[j] = o;
          --j;
        }
        mem[j] = e;
      }
      if (overflow)
        atomicOr(&s_misc[1], kErrNonManifold);
      else
        for (unsigned i = 0; i + 1 < cnt; i += 2)
        {
          W.twin[mem[i]] = static_cast<Idx>(mem[i + 1]);
          W.twin[mem[i + 1]] = static_cast<Idx>(mem[i]);
        }
    }
  }
  __syncthreads();
  // slot index -> first-inserter rank of the merged point.  Each thread owns the pair (2j, 2j+1), so reading both slots
  // before overwriting them is hazard free.  Dropped lines are marked with twin == self.
  {
    for (unsigned j = tid; j < n; j += nth)
    {
      const unsigned s0 = W.etag[2 * j], s1 = W.etag[2 * j + 1];
      const unsigned t0 = W.slots[s0].tag, t1 = W.slots[s1].tag;
      W.etag[2 * j] = t0;
      W.etag[2 * j + 1] = t1;
      if (s0 == s1)
      {
        W.twin[2 * j] = static_cast<Idx>(2 * j);  // dead marker
        W.twin[2 * j + 1] = static_cast<Idx>(2 * j + 1);
      }
    }
  }
  __syncthreads();
  if (s_misc[1] & kErrNonManifold)
    return;

  // ---- R: list ranking of the darts (Helman & JaJa: splitters, sequential sublist walks, pointer jumping on the
  //         splitters only) ----------------------------------------------------------------------------------------
  // dart e = "traverse line e >> 1 starting at its end e"; next(e) = twin(e ^ 1).  Wanted per dart: rnk = number of
  // hops to the last dart of its list, lst = that last dart.  Every open chain gives two lists (one per direction).
  // Splitters are the list heads (darts starting at an open end) plus both darts of every 16th line of the bucket; each
  // splitter walks its sublist (about 16 hops) up to the next splitter, then only the ~E/16 splitters are ranked by
  // pointer jumping: O(E) work instead of Wyllie's O(E log E).
  Idx* const r1 = reinterpret_cast<Idx*>(W.epos);
  Idx* const lst = r1;                 // during the walk: owner = compact index of the splitter whose sublist holds e
  Idx* const rnk = r1 + E;             // during the walk: hops from that splitter
  Idx* const chainEnd = r1 + 2ull * E;
  Idx* cArr[7];                        // compact splitter arrays: dart, next[2], dist[2], last[2]
  {
    const unsigned inR1 = static_cast<unsigned>(W.eposBytes / (sizeof(Idx) * static_cast<size_t>(E > 0 ? E : 1)));
    Idx* const r2 = reinterpret_cast<Idx*>(W.slots);
    unsigned a1 = 3, a2 = 0;
    for (int i = 0; i < 7; ++i)
      cArr[i] = (a1 < inR1) ? r1 + static_cast<size_t>(a1++) * E : r2 + static_cast<size_t>(a2++) * E;
  }
  Idx* const cDart = cArr[0];
  Idx* cNext[2] = { cArr[1], cArr[2] };
  Idx* cDist[2] = { cArr[3], cArr[4] };
  Idx* cLast[2] = { cArr[5], cArr[6] };
  for (unsigned e = tid; e < E; e += nth)
  {
    lst[e] = kNone;
    chainEnd[e] = kNone;
  }
  __syncthreads();
  for (unsigned e0 = 0; e0 < E; e0 += nth)
  {
    const unsigned e = e0 + tid;
    bool sel = false;
    if (e < E)
    {
      const Idx tw = W.twin[e];
      sel = tw != static_cast<Idx>(e) && (tw == kNone || ((e >> 1) & kSplitMask) == 0u);
    }
    const unsigned m = __ballot_sync(kFull, sel);
    unsigned base = 0;
    if ((tid & 31u) == 0u && m)
      base = atomicAdd(&s_misc[2], __popc(m));
    base = __shfl_sync(kFull, base, 0);
    if (sel)
    {
      const unsigned c = base + __popc(m & ((1u << (tid & 31u)) - 1u));
      cDart[c] = static_cast<Idx>(e);
      lst[e] = static_cast<Idx>(c);
    }
  }
  __syncthreads();
  const unsigned Ns = s_misc[2];
  for (unsigned c = tid; c < Ns; c += nth)
  {
    unsigned d = cDart[c], j = 0;
    while (true)
    {
      if (j)
        lst[d] = static_cast<Idx>(c);
      rnk[d] = static_cast<Idx>(j);
      const Idx nd = W.twin[d ^ 1u];
      if (nd == kNone)  // d ends at an open end: last dart of the list
      {
        cNext[0][c] = kNone;
        cDist[0][c] = static_cast<Idx>(j);
        cLast[0][c] = static_cast<Idx>(d);
        break;
      }
      if (((nd >> 1) & kSplitMask) == 0u)  // the next dart is a splitter (a list head never has a predecessor)
      {
        cNext[0][c] = lst[nd];
        cDist[0][c] = static_cast<Idx>(j + 1u);
        cLast[0][c] = kNone;
        break;
      }
      d = nd;
      ++j;
    }
  }
  __syncthreads();
  int cur = 0;
  {
    unsigned rounds = 1;
    while ((1u << rounds) < Ns + 1u)
      ++rounds;
    ++rounds;
    for (unsigned r = 0; r < rounds; ++r)
    {
      unsigned pending = 0;
      for (unsigned c = tid; c < Ns; c += nth)
      {
        const Idx nx = cNext[cur][c];
        if (nx != kNone)
        {
          const Idx nn = cNext[cur][nx];
          cNext[cur ^ 1][c] = nn;
          cDist[cur ^ 1][c] = static_cast<Idx>(cDist[cur][c] + cDist[cur][nx]);
          cLast[cur ^ 1][c] = cLast[cur][nx];
          pending |= (nn != kNone);
        }
        else
        {
          cNext[cur ^ 1][c] = kNone;
          cDist[cur ^ 1][c] = cDist[cur][c];
          cLast[cur ^ 1][c] = cLast[cur][c];
        }
      }
      if (pending)
        s_misc[5] = r + 1;
      __syncthreads();
      cur ^= 1;
      const unsigned more = s_misc[5];
      __syncthreads();  // everyone has read the flag before the next round may overwrite it
      if (more != r + 1)
        break;
    }
  }
  {
    // owner / hops -> lst / rnk, in place.  Darts no splitter reached, and darts whose splitters never reached an end,
    // lie on closed loops: lst = kNone marks them for the loop walker below.
    unsigned cyclic = 0;
    for (unsigned e = tid; e < E; e += nth)
    {
      if (W.twin[e] == static_cast<Idx>(e))  // dropped line
      {
        lst[e] = static_cast<Idx>(e);
        rnk[e] = 0;
        continue;
      }
      const Idx c = lst[e];
      if (c == kNone || cNext[cur][c] != kNone)
      {
        lst[e] = kNone;
        rnk[e] = 0;
        cyclic = 1;
      }
      else
      {
        rnk[e] = static_cast<Idx>(cDist[cur][c] - rnk[e]);
        lst[e] = cLast[cur][c];
      }
    }
    if (cyclic)
      s_misc[7] = 1u;
  }
  __syncthreads();

  // ---- C: chain heads, orientation, order ---------------------------------------------------------------------------
  struct Head
  {
    double key;    // first point . cut_direction
    unsigned e;    // head dart
    unsigned tag;  // tag (2 * triangle + line end) of the chain's first line end: tie-break of the order
  };
  Head* heads = reinterpret_cast<Head*>(W.slots);  // slots are dead; 16 B per head, cap >= E >= 2 * heads
  unsigned* lenSorted = reinterpret_cast<unsigned*>(W.twin);

  if (s_misc[7])
  {
    // Closed loops are rare (the reference targets open, roughly planar surfaces): one thread walks them.  A loop
    // starts at end 0 of its lowest-tag line and first walks that line, as vtkStripper does from its seed; it is
    // reversed, keeping the start point, when that first step points against the cut direction.
    if (tid == 0)
    {
      for (unsigned e0 = 0; e0 < E; ++e0)
      {
        if (lst[e0] != kNone)
          continue;
        unsigned c = e0, best = endTag(e0), len = 0;
        for (unsigned d = e0;;)
        {
          const unsigned t = endTag(d);
          if (t < best)
          {
            best = t;
            c = d;
          }
          ++len;
          d = W.twin[d ^ 1u];
          if (d == e0)
            break;
        }
        bool forward = false;
        if ((best & 1u) == 0u)
        {
          const float4 a = endPos(c), b = endPos(c ^ 1u);
          const double dd = dotLR(dsub(static_cast<double>(b.x), static_cast<double>(a.x)),
                                  dsub(static_cast<double>(b.y), static_cast<double>(a.y)),
                                  dsub(static_cast<double>(b.z), static_cast<double>(a.z)), P.L.dir);
          forward = !(dd < 0.0);
          // walk c = d_0 .. d_{len-1}; selected darts get list ranks, the others become inert singletons
          unsigned d = c;
          const unsigned tailF = [&] {
            unsigned q = c;
            for (unsigned i = 0; i + 1 < len; ++i)
              q = W.twin[q ^ 1u];
            return q;
          }();
          for (unsigned i = 0; i < len; ++i)
          {
            const unsigned nd = W.twin[d ^ 1u];
            const unsigned m = d ^ 1u;  // reverse dart d'_{len-1-i}
            if (forward)
            {
              rnk[d] = static_cast<Idx>(len - 1 - i), lst[d] = static_cast<Idx>(tailF);
              rnk[m] = 0, lst[m] = static_cast<Idx>(m);
            }
            else
            {
              rnk[d] = 0, lst[d] = static_cast<Idx>(d);
              rnk[m] = static_cast<Idx>(i), lst[m] = static_cast<Idx>(c ^ 1u);
            }
            d = nd;
          }
          const unsigned h = atomicAdd(&s_misc[0], 1u);
          heads[h].key = dotLR(static_cast<double>(a.x), static_cast<double>(a.y), static_cast<double>(a.z), P.L.dir);
          heads[h].e = forward ? c : (tailF ^ 1u);
          heads[h].tag = best;
        }
        // a loop whose lowest tag is odd is the mirror image of a loop handled (or to be handled) above; the walk of
        // its canonical twin rewrites these darts, so nothing to do here
      }
      // whatever is still marked as cyclic now belongs to mirror loops whose canonical twin was visited later
      for (unsigned e0 = 0; e0 < E; ++e0)
        if (lst[e0] == kNone)
          rnk[e0] = 0, lst[e0] = static_cast<Idx>(e0);
    }
    __syncthreads();
  }

  // twin[] is still needed to find heads: collect first, then reuse
  for (unsigned e = tid; e < E; e += nth)
  {
    const Idx tw = W.twin[e];
    if (tw != kNone)
      continue;  // interior end, or dead line (twin == self)
    // open end: head of the dart list that starts here
    const unsigned tail = lst[e];
    const unsigned tagFirst = endTag(e), tagLast = endTag(tail ^ 1u);
    const float4 a = endPos(e), b = endPos(tail ^ 1u);
    const double d = dotLR(dsub(static_cast<double>(b.x), static_cast<double>(a.x)),
                           dsub(static_cast<double>(b.y), static_cast<double>(a.y)),
                           dsub(static_cast<double>(b.z), static_cast<double>(a.z)), P.L.dir);
    if (d > 0.0 || (d == 0.0 && tagFirst < tagLast))
    {
      const unsigned h = atomicAdd(&s_misc[0], 1u);
      heads[h].key = dotLR(static_cast<double>(a.x), static_cast<double>(a.y), static_cast<double>(a.z), P.L.dir);
      heads[h].e = e;
      heads[h].tag = tagFirst;
    }
  }
  __syncthreads();
  const unsigned S = s_misc[0];
  // order of each head = number of heads that sort before it (key, then first-inserter rank)
  unsigned* headOrder = reinterpret_cast<unsigned*>(r1 + 3ull * E);  // dead splitter array: room for n entries
  for (unsigned h = tid; h < S; h += nth)
  {
    unsigned ord = 0;
    const double key = heads[h].key;
    const unsigned tg = heads[h].tag;
    for (unsigned g = 0; g < S; ++g)
    {
      const double kg = heads[g].key;
      ord += (kg < key) || (kg == key && heads[g].tag < tg);
    }
    headOrder[h] = ord;
    lenSorted[ord] = static_cast<unsigned>(rnk[heads[h].e]) + 2u;  // lines + 1 points
  }
  __syncthreads();
  // exclusive scan of the chain lengths (few chains per plane: one warp, serial over 32-wide chunks)
  if (tid < 32)
  {
    unsigned carry = 0;
    for (unsigned base = 0; base < S; base += 32)
    {
      const unsigned i = base + tid;
      const unsigned v = i < S ? lenSorted[i] : 0u;
      unsigned x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const unsigned y = __shfl_up_sync(kFull, x, o);
        if (tid >= static_cast<unsigned>(o))
          x += y;
      }
      if (i < S)
        lenSorted[i] = carry + x - v;  // exclusive offset; length recoverable from the next entry
      carry += __shfl_sync(kFull, x, 31);
    }
    if (tid == 0)
      s_misc[4] = carry;  // waypoints of this plane
  }
  __syncthreads();
  const unsigned Wk = s_misc[4];

  if (tid == 0)
  {
    P.planeMeta[2 * kLocal] = Wk;
    P.planeMeta[2 * kLocal + 1] = S;
  }
  const unsigned long long baseW = 2ull * off, baseS = off;  // the plane's own output windows (see LinkParams)

  // ---- O: output ----------------------------------------------------------------------------------------------------
  for (unsigned h = tid; h < S; h += nth)
  {
    const unsigned ord = headOrder[h];
    const unsigned e = heads[h].e;
    const unsigned lines = static_cast<unsigned>(rnk[e]) + 1u;
    chainEnd[lst[e]] = static_cast<Idx>(lenSorted[ord] + lines);  // plane-local index of the chain's last point
    P.segLen[baseS + ord] = lines + 1u;
  }
  __syncthreads();
  for (unsigned e = tid; e < E; e += nth)
  {
    if (W.etag[e] == W.etag[e ^ 1u])
      continue;  // dropped line
    const Idx ce = chainEnd[lst[e]];
    if (ce == kNone)
      continue;  // dart of the non-selected direction
    const unsigned r = rnk[e];
    const unsigned last = static_cast<unsigned>(ce);
    P.wpTag[baseW + last - 1u - r] = W.etag[e];
    if (r == 0)
      P.wpTag[baseW + last] = W.etag[e ^ 1u];
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Fast path of S5 for the planes met in practice: the work area is in shared memory (16-bit indices), no cut point has
// more than two proper line ends and no contour closes on itself.  Same decisions as linkPlane above (which remains the
// specification and handles every other plane: the fast path returns false before writing any output and the caller
// re-runs the plane through linkPlane), organised for instruction count: one thread per cut LINE (its two endpoints
// are independent work, and adjacent 16- / 32-bit entries move as one 32- / 64-bit access), fused sweeps, splitters
// every 8th line.
// -----------------------------------------------------------------------------------------------------------------
constexpr unsigned kFastSplitMask = 7u;

/** Cheap hash of a float3 whose upper half serves as a 16-bit fingerprint (-0 folded onto +0: operator== semantics) */
__device__ __forceinline__ unsigned hashPointFast(float x, float y, float z)
{
  const unsigned a = __float_as_uint(x == 0.f ? 0.f : x), b = __float_as_uint(y == 0.f ? 0.f : y),
                 c = __float_as_uint(z == 0.f ? 0.f : z);
  // (rotated against one another: with plain sums the sign bits of two coordinates cancel, and the points (x, y, z) and
  //  (-x, y, -z) of a point-symmetric part would always collide)
  unsigned h = a * 0x9E3779B1u + __funnelshift_l(b, b, 11) * 0x85EBCA77u + __funnelshift_l(c, c, 22) * 0xC2B2AE3Du;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  return h ^ (h >> 13);
}

__device__ bool linkPlaneFast(const LinkParams& P, uint8_t* smem, long long kLocal, unsigned n, unsigned cap8,
                              unsigned* s_misc)
{
  using Idx = uint16_t;
  constexpr unsigned kNone = 0xffffu;
  const unsigned E = 2 * n;
  const unsigned tid = threadIdx.x, nth = blockDim.x;
  const unsigned off = P.planeOff[kLocal];
  const float4* __restrict__ rec = P.rec + 2ull * off;

  // Work area, same footprint as PlaneWork<uint16_t> (12 E | 4 E | 8 cap8 | 2 E bytes):
  //   [0, 12 E)    endpoint positions px py pz while the cut points are merged; afterwards twin | lst | rnk | chainEnd
  //                (2 E each) and the seven compact splitter arrays (4 E); the pairing cells alias lst + rnk
  //   etag         own rank of each endpoint; the key holder's entry becomes the smallest rank of its point
  //   slots        4-byte slots {key holder : 16, fingerprint : 16}, twice as many as the generic path's 8-byte ones
  //   slotIdx      slot of each endpoint (16 bit)
  float* const px = reinterpret_cast<float*>(smem);
  float* const py = px + E;
  float* const pz = py + E;
  unsigned* const etag = reinterpret_cast<unsigned*>(smem + ((12ull * E + 15) & ~size_t(15)));
  unsigned* const slots = etag + E;
  const unsigned cap = 2 * cap8, mask = cap - 1;
  Idx* const slotIdx = reinterpret_cast<Idx*>(slots + cap);
  Idx* const twin = reinterpret_cast<Idx*>(smem);
  Idx* const lst = twin + E;       // owner splitter during the walk, then last dart of the list
  Idx* const rnk = lst + E;        // hops from the owner, then hops to the last dart
  Idx* const chainEnd = rnk + E;
  unsigned* const pair = reinterpret_cast<unsigned*>(lst);  // [E] pairing cells, indexed by the key holder of the point
  // compact splitter arrays in [8 E, 12 E): one 64-bit entry {next splitter : 16, last dart : 16, hops : 16} and the
  // splitter's dart, 10 bytes per splitter
  const unsigned nsCap = (2u * E) / 5u;  // splitters they hold (more: leave the plane to linkPlane)
  unsigned long long* const cEnt = reinterpret_cast<unsigned long long*>(chainEnd + E);
  Idx* const cDart = reinterpret_cast<Idx*>(cEnt + nsCap);
  auto packEnt = [](unsigned next, unsigned last, unsigned hops) {
    return static_cast<unsigned long long>(next | (last << 16)) | (static_cast<unsigned long long>(hops) << 32);
  };
  if (n < 64 || cap > 65536u)
    return false;  // tiny planes are not worth a special case; 16-bit slot indices need cap <= 2^16

  if (tid < kMiscWords)
    s_misc[tid] = 0;

  // The records of the plane this CTA will most likely take next (tickets advance in step: kLocal + grid size) are
  // started towards L2 now, a whole plane ahead of their use; a wrong guess only warms L2 for another CTA.
  {
    const long long kNext = kLocal + gridDim.x;
    if (kNext < P.L.plane_end - P.L.plane_begin)
    {
      const unsigned o = P.planeOff[kNext], m = P.planeOff[kNext + 1] - o;
      const char* base = reinterpret_cast<const char*>(P.rec + 2ull * o);
      for (unsigned b = tid * 128u; b < m * 32u; b += nth * 128u)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b));
    }
  }

  // ---- A: stage the records; empty hash ---------------------------------------------------------------------------------
  for (unsigned j = tid; j < n; j += nth)
  {
    const float4 r0 = __ldg(rec + 2 * j), r1 = __ldg(rec + 2 * j + 1);
    *reinterpret_cast<float2*>(px + 2 * j) = make_float2(r0.x, r1.x);
    *reinterpret_cast<float2*>(py + 2 * j) = make_float2(r0.y, r1.y);
    *reinterpret_cast<float2*>(pz + 2 * j) = make_float2(r0.z, r1.z);
    *reinterpret_cast<uint2*>(etag + 2 * j) = make_uint2(__float_as_uint(r0.w), __float_as_uint(r1.w));
  }
  for (unsigned s = 2 * tid; s < cap; s += 2 * nth)  // (the slot array is 8-byte aligned)
    *reinterpret_cast<uint2*>(slots + s) = make_uint2(kEmpty, kEmpty);
  __syncthreads();

  // ---- H1: merge cut points with equal float coordinates; smallest rank per point ------------------------------------
  // A probe compares the 16-bit fingerprint first and touches the positions only on a match; the table is at most a
  // quarter full.  Probe sequences differ in length, so a lane does not wait for its neighbours: the moment its
  // endpoint is placed it takes its next one (endpoints tid, tid + nth, ...), which keeps the lanes of a warp busy.
  {
    unsigned e = tid, s = 0, fp = 0, tg = 0;
    float x = 0.f, y = 0.f, z = 0.f;
    bool active = e < E;
    auto fetch = [&]() {
      x = px[e], y = py[e], z = pz[e];
      tg = etag[e];
      const unsigned h = hashPointFast(x, y, z);
      fp = h >> 16;
      s = h & mask;
    };
    if (active)
      fetch();
    while (active)
    {
      unsigned v = *reinterpret_cast<volatile unsigned*>(slots + s);
      if (v == kEmpty)
        v = atomicCAS(slots + s, kEmpty, e | (fp << 16));
      unsigned f = e;
      bool placed = v == kEmpty;  // claimed: e is the key holder of a new point
      if (!placed && (v >> 16) == fp)
      {
        const unsigned g = v & 0xffffu;
        if (px[g] == x && py[g] == y && pz[g] == z)
        {
          f = g;
          placed = true;
        }
      }
      if (placed)
      {
        if (f != e)
          atomicMin(etag + f, tg);
        slotIdx[e] = static_cast<Idx>(s);
        e += nth;
        active = e < E;
        if (active)
          fetch();
      }
      else
        s = (s + 1) & mask;
    }
  }
  __syncthreads();

  // ---- H2: pair the two line ends meeting at each point (the positions are dead: their region now holds the pairing
  //          cells, the twins and the list-ranking arrays) -------------------------------------------------------------------
  for (unsigned j = tid; j < n; j += nth)
  {
    pair[2 * j] = 0u;
    pair[2 * j + 1] = 0u;
    *reinterpret_cast<unsigned*>(twin + 2 * j) = 0xffffffffu;
  }
  __syncthreads();
  for (unsigned j = tid; j < n; j += nth)
  {
    const unsigned sl = *reinterpret_cast<const unsigned*>(slotIdx + 2 * j);
    if ((sl & 0xffffu) == (sl >> 16))
      continue;  // both ends merged into one point: the line is dropped (vtkTriangle::Contour: pts[0] != pts[1])
#pragma unroll
    for (int w = 0; w < 2; ++w)
    {
      const unsigned e = 2 * j + w;
      unsigned* cell = pair + (slots[w ? (sl >> 16) : (sl & 0xffffu)] & 0xffffu);
      unsigned old = *reinterpret_cast<volatile unsigned*>(cell);
      while (true)
      {
        const unsigned seen = atomicCAS(cell, old, (((old >> kCellShift) + 1u) << kCellShift) | e);
        if (seen == old)
          break;
        old = seen;
      }
      const unsigned arrivals = old >> kCellShift;
      if (arrivals == 1u)
      {
        const unsigned prev = old & kCellMask;
        twin[e] = static_cast<Idx>(prev);
        twin[prev] = static_cast<Idx>(e);
      }
      else if (arrivals >= 2u)
        s_misc[6] = 1u;  // a point with more than two line ends: linkPlane's business
    }
  }
  __syncthreads();
  if (s_misc[6])
    return false;

  // ---- F: merged ranks, dropped lines, list-ranking set-up and splitter selection, one sweep ---------------------------
  for (unsigned j0 = 0; j0 < n; j0 += nth)
  {
    const unsigned j = j0 + tid;
    bool sel0 = false, sel1 = false;
    if (j < n)
    {
      const unsigned sl = *reinterpret_cast<const unsigned*>(slotIdx + 2 * j);
      const unsigned f0 = slots[sl & 0xffffu] & 0xffffu, f1 = slots[sl >> 16] & 0xffffu;
      // (a key holder rewrites its own entry with the value it already has: no hazard with the readers of that entry)
      *reinterpret_cast<uint2*>(etag + 2 * j) = make_uint2(etag[f0], etag[f1]);
      *reinterpret_cast<unsigned*>(lst + 2 * j) = 0xffffffffu;
      *reinterpret_cast<unsigned*>(chainEnd + 2 * j) = 0xffffffffu;
      if (f0 == f1)
        *reinterpret_cast<unsigned*>(twin + 2 * j) = (2 * j) | ((2 * j + 1) << 16);  // dead markers: twin == self
      else
      {
        const unsigned tw = *reinterpret_cast<const unsigned*>(twin + 2 * j);
        const bool every = (j & kFastSplitMask) == 0u;
        sel0 = every || (tw & 0xffffu) == kNone;  // a dart that starts at an open end heads a list
        sel1 = every || (tw >> 16) == kNone;
      }
    }
    const unsigned m0 = __ballot_sync(kFull, sel0), m1 = __ballot_sync(kFull, sel1);
    unsigned base = 0;
    if ((tid & 31u) == 0u && (m0 | m1))
      base = atomicAdd(&s_misc[2], __popc(m0) + __popc(m1));
    base = __shfl_sync(kFull, base, 0);
    const unsigned below = (1u << (tid & 31u)) - 1u;
    if (sel0)
    {
      const unsigned c = base + __popc(m0 & below);
      if (c < nsCap)
        cDart[c] = static_cast<Idx>(2 * j), lst[2 * j] = static_cast<Idx>(c);
    }
    if (sel1)
    {
      const unsigned c = base + __popc(m0) + __popc(m1 & below);
      if (c < nsCap)
        cDart[c] = static_cast<Idx>(2 * j + 1), lst[2 * j + 1] = static_cast<Idx>(c);
    }
  }
  __syncthreads();
  const unsigned Ns = s_misc[2];
  if (Ns > nsCap)
    return false;  // a crowd of short chains: linkPlane's general layout takes it

  // ---- R1: every splitter walks its sublist up to the next splitter (or the end of the list) --------------------------
  for (unsigned c = tid; c < Ns; c += nth)
  {
    unsigned d = cDart[c], j = 0;
    while (true)
    {
      if (j)
        lst[d] = static_cast<Idx>(c);
      rnk[d] = static_cast<Idx>(j);
      const unsigned nd = twin[d ^ 1u];
      if (nd == kNone)  // d ends at an open end: last dart of the list
      {
        cEnt[c] = packEnt(kNone, d, j);
        break;
      }
      if (((nd >> 1) & kFastSplitMask) == 0u)  // the next dart is a splitter (a list head has no predecessor)
      {
        cEnt[c] = packEnt(lst[nd], kNone, j + 1u);
        break;
      }
      d = nd;
      ++j;
    }
  }
  __syncthreads();

  // ---- R2: pointer jumping on the splitters, in place: an entry is one 64-bit word, so whichever version of its
  //          successor a thread reads (this round's or the previous one's) is a consistent {next, last, hops} triple ------
  {
    unsigned rounds = 1;
    while ((1u << rounds) < Ns + 1u)
      ++rounds;
    ++rounds;
    volatile unsigned long long* ent = cEnt;
    for (unsigned r = 0; r < rounds; ++r)
    {
      int pending = 0;
      for (unsigned c = tid; c < Ns; c += nth)
      {
        const unsigned long long me = ent[c];
        const unsigned nx = static_cast<unsigned>(me) & 0xffffu;
        if (nx != kNone)
        {
          const unsigned long long o = ent[nx];
          ent[c] = (o & 0xffffffffull) | ((me & 0xffffffff00000000ull) + (o & 0xffffffff00000000ull));
          pending |= (static_cast<unsigned>(o) & 0xffffu) != kNone;
        }
      }
      if (!__syncthreads_or(pending))
        break;
    }
  }

  // ---- R3: owner / hops -> last dart / hops to it, in place; open ends become chain heads -------------------------------
  struct Head
  {
    double key;    // first point . cut_direction
    unsigned e;    // head dart
    unsigned tag;  // rank (2 * triangle + line end) of the chain's first line end: tie-break of the order
  };
  Head* const heads = reinterpret_cast<Head*>(slots);  // the hash is dead; 16 B per head, 4 cap >= 8 E bytes
  for (unsigned j = tid; j < n; j += nth)
  {
    const unsigned tw = *reinterpret_cast<const unsigned*>(twin + 2 * j);
    if ((tw & 0xffffu) == 2 * j)  // dropped line
    {
      *reinterpret_cast<unsigned*>(lst + 2 * j) = (2 * j) | ((2 * j + 1) << 16);
      *reinterpret_cast<unsigned*>(rnk + 2 * j) = 0u;
      continue;
    }
    const unsigned ow = *reinterpret_cast<const unsigned*>(lst + 2 * j);
    const unsigned lo = *reinterpret_cast<const unsigned*>(rnk + 2 * j);
    const unsigned c0 = ow & 0xffffu, c1 = ow >> 16;
    bool cyclic = c0 == kNone || c1 == kNone;
    unsigned l0 = 0, l1 = 0, r0 = 0, r1 = 0;
    if (!cyclic)
    {
      const unsigned long long e0 = cEnt[c0], e1 = cEnt[c1];
      cyclic = (static_cast<unsigned>(e0) & 0xffffu) != kNone || (static_cast<unsigned>(e1) & 0xffffu) != kNone;
      l0 = (static_cast<unsigned>(e0) >> 16) & 0xffffu, l1 = (static_cast<unsigned>(e1) >> 16) & 0xffffu;
      r0 = static_cast<unsigned>(e0 >> 32) - (lo & 0xffffu), r1 = static_cast<unsigned>(e1 >> 32) - (lo >> 16);
    }
    if (cyclic)
    {
      s_misc[7] = 1u;  // a closed contour: linkPlane walks those
      continue;
    }
    *reinterpret_cast<unsigned*>(lst + 2 * j) = l0 | (l1 << 16);
    *reinterpret_cast<unsigned*>(rnk + 2 * j) = (r0 & 0xffffu) | (r1 << 16);
#pragma unroll
    for (int w = 0; w < 2; ++w)
    {
      if ((w ? (tw >> 16) : (tw & 0xffffu)) != kNone)
        continue;
      // dart e starts at an open end: it heads one of the two lists of its chain; keep the list that runs along the
      // cut direction (first -> last point), ties to the lower rank
      const unsigned e = 2 * j + w, tail = w ? l1 : l0;
      const float4 a = __ldg(rec + e), b = __ldg(rec + (tail ^ 1u));
      const unsigned tagFirst = __float_as_uint(a.w), tagLast = __float_as_uint(b.w);
      const double d = dotLR(dsub(static_cast<double>(b.x), static_cast<double>(a.x)),
                             dsub(static_cast<double>(b.y), static_cast<double>(a.y)),
                             dsub(static_cast<double>(b.z), static_cast<double>(a.z)), P.L.dir);
      if (d > 0.0 || (d == 0.0 && tagFirst < tagLast))
      {
        const unsigned h = atomicAdd(&s_misc[0], 1u);
        heads[h].key = dotLR(static_cast<double>(a.x), static_cast<double>(a.y), static_cast<double>(a.z), P.L.dir);
        heads[h].e = e;
        heads[h].tag = tagFirst;
      }
    }
  }
  __syncthreads();
  if (s_misc[7])
    return false;

  // ---- C: order of the chains ---------------------------------------------------------------------------------------------
  const unsigned S = s_misc[0];
  unsigned* const headOrder = reinterpret_cast<unsigned*>(cEnt);    // the splitter arrays are dead: 4 E bytes
  unsigned* const lenSorted = reinterpret_cast<unsigned*>(slotIdx);  // 2 E bytes: room for the at most E / 2 chains
  for (unsigned h = tid; h < S; h += nth)
  {
    unsigned ord = 0;
    const double key = heads[h].key;
    const unsigned tg = heads[h].tag;
    for (unsigned g = 0; g < S; ++g)
    {
      const double kg = heads[g].key;
      ord += (kg < key) || (kg == key && heads[g].tag < tg);
    }
    headOrder[h] = ord;
    lenSorted[ord] = static_cast<unsigned>(rnk[heads[h].e]) + 2u;  // lines + 1 points
  }
  __syncthreads();
  if (tid < 32)
  {
    unsigned carry = 0;
    for (unsigned base = 0; base < S; base += 32)
    {
      const unsigned i = base + tid;
      const unsigned v = i < S ? lenSorted[i] : 0u;
      unsigned x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const unsigned y = __shfl_up_sync(kFull, x, o);
        if (tid >= static_cast<unsigned>(o))
          x += y;
      }
      if (i < S)
        lenSorted[i] = carry + x - v;
      carry += __shfl_sync(kFull, x, 31);
    }
    if (tid == 0)
    {
      P.planeMeta[2 * kLocal] = carry;
      P.planeMeta[2 * kLocal + 1] = S;
    }
  }
  __syncthreads();
  const unsigned long long baseW = 2ull * off, baseS = off;

  // ---- O: output ------------------------------------------------------------------------------------------------------------
  for (unsigned h = tid; h < S; h += nth)
  {
    const unsigned ord = headOrder[h];
    const unsigned e = heads[h].e;
    const unsigned lines = static_cast<unsigned>(rnk[e]) + 1u;
    chainEnd[lst[e]] = static_cast<Idx>(lenSorted[ord] + lines);  // plane-local index of the chain's last point
    P.segLen[baseS + ord] = lines + 1u;
  }
  __syncthreads();
  for (unsigned j = tid; j < n; j += nth)
  {
    const unsigned la = *reinterpret_cast<const unsigned*>(lst + 2 * j);
    const unsigned rk = *reinterpret_cast<const unsigned*>(rnk + 2 * j);
    const uint2 tg = *reinterpret_cast<const uint2*>(etag + 2 * j);
    const unsigned ce0 = chainEnd[la & 0xffffu], ce1 = chainEnd[la >> 16];
    // exactly one of the two darts of a line belongs to the selected direction of its chain (none for a dropped line)
    if (ce0 != kNone)
    {
      const unsigned r = rk & 0xffffu;
      P.wpTag[baseW + ce0 - 1u - r] = tg.x;
      if (r == 0)
        P.wpTag[baseW + ce0] = tg.y;
    }
    if (ce1 != kNone)
    {
      const unsigned r = rk >> 16;
      P.wpTag[baseW + ce1 - 1u - r] = tg.y;
      if (r == 0)
        P.wpTag[baseW + ce1] = tg.x;
    }
  }
  return true;
}

// -----------------------------------------------------------------------------------------------------------------
// S5, second generation (link2_ops.h): no shared-memory atomics, 4 bytes of hash per endpoint staged instead of 16 bytes
// of coordinates, one dart per line, ranks from the chain heads; two planes resident per SM.
// -----------------------------------------------------------------------------------------------------------------
struct DevOps
{
  static __device__ __forceinline__ void orr(unsigned* a, unsigned v)
  {
    if ((*reinterpret_cast<volatile unsigned*>(a) & v) != v)
      atomicOr(a, v);
  }
  static __device__ __forceinline__ unsigned add(unsigned* a, unsigned v) { return atomicAdd(a, v); }
  static __device__ __forceinline__ unsigned cas(unsigned* a, unsigned cmp, unsigned v) { return atomicCAS(a, cmp, v); }
  /** a place in a compacted list for every thread that wants one: one atomic per warp.  Warp-wide: all lanes call it. */
  static __device__ __forceinline__ unsigned claim(unsigned* counter, bool wanted)
  {
    const unsigned m = __ballot_sync(kFull, wanted);
    if (m == 0u)
      return 0u;
    const unsigned lane = threadIdx.x & 31u;
    unsigned base = 0;
    if (lane == static_cast<unsigned>(__ffs(m)) - 1u)
      base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(kFull, base, __ffs(m) - 1);
    return base + __popc(m & ((1u << lane) - 1u));
  }
  static __device__ __forceinline__ unsigned vload(const unsigned* a) { return *reinterpret_cast<const volatile unsigned*>(a); }
  static __device__ __forceinline__ link2::Rec4 load(const link2::Rec4* p)
  {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return link2::Rec4{ v.x, v.y, v.z, v.w };
  }
  static __device__ __forceinline__ unsigned long long ld64(const unsigned long long* p)
  {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
  }
  static __device__ __forceinline__ void st64(unsigned long long* p, unsigned long long v)
  {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
  }
};

// NB2_LINK_PROFILE (environment, read once by the launcher): clock cycles per phase of the fast path, summed over the
// planes by thread 0 of every CTA (diagnostics only; LinkParams::phaseCycles is null otherwise)
struct PhaseClock
{
  unsigned long long* acc;
  long long last;
  __device__ explicit PhaseClock(unsigned long long* a) : acc(a), last(0)
  {
    if (acc && threadIdx.x == 0)
      last = clock64();
  }
  __device__ __forceinline__ void lap(int phase)
  {
    if (acc && threadIdx.x == 0)
    {
      const long long now = clock64();
      atomicAdd(acc + phase, static_cast<unsigned long long>(now - last));
      last = now;
    }
  }
};

/** The second-generation fast path on one plane: false when the plane needs another path (what it wrote by then lies in
 *  the plane's own output windows and is overwritten) */
__device__ __forceinline__ bool linkPlane2(const LinkParams& P, uint8_t* smem, long long kLocal, unsigned n)
{
  using namespace link2;
  const unsigned tid = threadIdx.x, nth = blockDim.x;
  if (!fitsThreadSet(n, nth))
    return false;
  const unsigned off = P.planeOff[kLocal];
  const Rec4* rec = reinterpret_cast<const Rec4*>(P.rec + 2ull * off);
  const Work W = carve(smem, n);
  PhaseClock clk(P.phaseCycles);
  phaseInit<DevOps>(W, tid, nth);
  __syncthreads();
  clk.lap(0);
  unsigned todo = phaseHash<DevOps>(W, rec, tid, nth);
  __syncthreads();
  clk.lap(1);
  // Claim rounds with plain stores; the first claim rode along with the hashing.  What the plain rounds leave unsettled
  // finishes with compare-and-swap (phaseStragglers).  Measured on cfg3 (profiles/r02_summary.md): one plain round plus
  // stragglers (0.226 ms) and many plain rounds (0.241 ms) are about level; a variant that probed with compare-and-swap
  // inside an unrolled settle phase took 0.45 - 0.8 ms (the probe loop replicated per endpoint of the batch, few lanes each).
  bool left = true;
  for (unsigned round = 0; round < P.plainRounds && left; ++round)
  {
    if (round)
    {
      phaseClaim<DevOps>(W, todo, tid, nth);
      __syncthreads();
    }
    todo = phaseSettle<DevOps>(W, todo, tid, nth);
    if (round + 1u == P.plainRounds)
      break;  // (the stragglers follow without a barrier: occupied slots are final, empty ones are taken by compare-and-swap)
    left = __syncthreads_or(todo != 0u ? 1 : 0) != 0;
  }
  if (left)
  {
    phaseStragglers<DevOps>(W, todo, tid, nth);
    __syncthreads();
  }
  clk.lap(2);
  phasePair<DevOps>(W, tid, nth);
  __syncthreads();
  if (W.misc[M_ANYDEAD])
  {
    phaseDead<DevOps>(W, tid, nth);
    __syncthreads();
    if (W.misc[M_FAIL])
      return false;
  }
  clk.lap(3);
  phaseLink<DevOps>(W, tid, nth);
  __syncthreads();
  clk.lap(4);
  if (W.misc[M_FAIL])
    return false;
  phaseWalk<DevOps>(W, tid, nth);
  __syncthreads();
  clk.lap(5);
  for (unsigned round = 0, rounds = jumpRounds(W.misc[M_NS]); round < rounds; ++round)
    if (!__syncthreads_or(phaseJump<DevOps>(W, tid, nth) ? 1 : 0))
      break;
  clk.lap(6);
  phaseChains<DevOps>(W, tid, nth);
  __syncthreads();
  if (W.misc[M_FAIL])
    return false;
  phaseHeads<DevOps>(W, rec, P.L.dir, tid, nth);
  __syncthreads();
  if (W.misc[M_FAIL])
    return false;
  phaseOrder<DevOps>(W, tid, nth);
  __syncthreads();
  // exclusive prefix of the chain sizes (few chains per plane: one warp, serial over 32-wide chunks)
  if (tid < 32)
  {
    const unsigned S = W.misc[M_NH];
    unsigned carry = 0;
    for (unsigned base = 0; base < S; base += 32)
    {
      const unsigned i = base + tid;
      const unsigned v = i < S ? W.sortedLen[i] : 0u;
      unsigned x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const unsigned y = __shfl_up_sync(kFull, x, o);
        if (tid >= static_cast<unsigned>(o))
          x += y;
      }
      if (i < S)
        W.sortedLen[i] = carry + x - v;
      carry += __shfl_sync(kFull, x, 31);
    }
    if (tid == 0)
      W.misc[M_WK] = carry;
  }
  __syncthreads();
  clk.lap(7);
  Out o;
  o.wpTag = P.wpTag + 2ull * off;  // the plane's own output windows (see LinkParams)
  o.segLen = P.segLen + off;
  phaseOutput<DevOps>(W, rec, o, tid, nth);
  __syncthreads();
  clk.lap(8);
  if (W.misc[M_FAIL])
    return false;
  if (tid == 0)
  {
    P.planeMeta[2 * kLocal] = W.misc[M_WK];
    P.planeMeta[2 * kLocal + 1] = W.misc[M_NH];
  }
  return true;
}

/** Exclusive prefix of the per-plane waypoint and segment counts (one block; nP is at most a few thousand; the counts
 *  were written by other blocks and are read past L1).
 *  prefix[k] = {waypoints, segments} before plane k; prefix[nP] = totals, also written to totals[0..1]. */
__device__ void scanMetaBlock(const unsigned* planeMeta, long long nP, uint2* __restrict__ prefix,
                              unsigned* __restrict__ totals)
{
  __shared__ uint2 s_warp[32];
  __shared__ uint2 s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    s_carry = make_uint2(0u, 0u);
  __syncthreads();
  for (long long base = 0; base < nP; base += blockDim.x)
  {
    const long long i = base + threadIdx.x;
    const uint2 v = (i < nP) ? __ldcg(reinterpret_cast<const uint2*>(planeMeta) + i) : make_uint2(0u, 0u);
    uint2 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const unsigned yx = __shfl_up_sync(kFull, x.x, o), yy = __shfl_up_sync(kFull, x.y, o);
      if (lane >= o)
        x.x += yx, x.y += yy;
    }
    if (lane == 31)
      s_warp[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
      uint2 w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const unsigned yx = __shfl_up_sync(kFull, w.x, o), yy = __shfl_up_sync(kFull, w.y, o);
        if (lane >= o)
          w.x += yx, w.y += yy;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const uint2 carry = s_carry;
    const uint2 wprev = warp ? s_warp[warp - 1] : make_uint2(0u, 0u);
    const uint2 incl = make_uint2(x.x + wprev.x + carry.x, x.y + wprev.y + carry.y);
    if (i < nP)
      prefix[i] = make_uint2(incl.x - v.x, incl.y - v.y);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1)
      s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    prefix[nP] = s_carry;
    totals[0] = s_carry.x;
    totals[1] = s_carry.y;
  }
}


__device__ __forceinline__ void linkPlanesBody(const LinkParams& Pin)
{
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ unsigned s_misc[kMiscWords];
  __shared__ unsigned s_ticket;
  if (*Pin.overflow & 1u)
    return;  // k_emit_hits wrote nothing
  LinkParams P = Pin;
  P.L = Pin.frame->L;
  const long long nP = P.L.plane_end - P.L.plane_begin;

  while (true)
  {
    __syncthreads();
    if (threadIdx.x == 0)
      s_ticket = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const long long kLocal = s_ticket;
    if (kLocal >= nP)
      break;
    const unsigned n = P.planeOff[kLocal + 1] - P.planeOff[kLocal];
    if (n == 0)
    {
      if (threadIdx.x == 0)
        P.planeMeta[2 * kLocal] = P.planeMeta[2 * kLocal + 1] = 0u;
      continue;
    }
    if (n > P.smemHitCap && n > P.scratchHits)
    {
      // larger than the global work areas were sized for: same protocol as the hit buffers
      if (threadIdx.x == 0)
      {
        atomicOr(P.overflow, 2u);
        P.planeMeta[2 * kLocal] = P.planeMeta[2 * kLocal + 1] = 0u;
      }
      continue;
    }
    const unsigned E = 2 * n;
    unsigned cap = 16;
    while (cap < E)
      cap <<= 1;
    if (n <= P.fastHitCap && linkPlane2(P, smem, kLocal, n))
      continue;
    // The second-generation path declined (a cut point with more than two line ends, a closed contour, inconsistently
    // wound triangles, a crowd of zero-length lines, a plane beyond its work area): the first-generation paths take it
    __syncthreads();
    if (n <= P.smemHitCap && linkPlaneFast(P, smem, kLocal, n, cap, s_misc))
    {
      // done
    }
    else if (n <= P.smemHitCap)
    {
      __syncthreads();
      if (threadIdx.x == 0)
        atomicAdd(P.generalCount, 1u);  // (diagnostics: planes no fast path took)
      PlaneWork<uint16_t> W;
      W.eposBytes = 12ull * E;
      W.epos = reinterpret_cast<float*>(smem);
      W.etag = reinterpret_cast<unsigned*>(smem + ((W.eposBytes + 15) & ~size_t(15)));
      W.slots = reinterpret_cast<Slot*>(reinterpret_cast<uint8_t*>(W.etag) + 4ull * E);
      W.twin = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(W.slots) + 8ull * cap);
      W.cap = cap;
      linkPlane<uint16_t>(P, W, kLocal, n, s_misc);
    }
    else if (n <= 16384u && linkPlaneFast(P, P.scratch + P.scratchStride * blockIdx.x, kLocal, n, cap, s_misc))
    {
      // a plane too large for shared memory but within 16-bit indices: the fast path again, its work area in this
      // CTA's slice of the global scratch (L2 resident)
    }
    else
    {
      __syncthreads();
      if (threadIdx.x == 0)
        atomicAdd(P.generalCount, 1u);
      uint8_t* g = P.scratch + P.scratchStride * blockIdx.x;
      PlaneWork<uint32_t> W;
      W.eposBytes = 32ull * E;  // 8 index arrays of 4 E bytes
      W.epos = reinterpret_cast<float*>(g);
      W.etag = reinterpret_cast<unsigned*>(g + ((W.eposBytes + 15) & ~size_t(15)));
      W.slots = reinterpret_cast<Slot*>(reinterpret_cast<uint8_t*>(W.etag) + 4ull * E);
      W.twin = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(W.slots) + 8ull * cap);
      W.cap = cap;
      linkPlane<uint32_t>(P, W, kLocal, n, s_misc);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_misc[1])
    {
      atomicOr(&P.status[0], s_misc[1]);
      atomicMin(&P.status[1], static_cast<unsigned>(kLocal));
      P.planeMeta[2 * kLocal] = P.planeMeta[2 * kLocal + 1] = 0u;
    }
  }
  // S5b: the last CTA to run out of planes closes the per-plane counts into prefixes and totals for k_waypoints
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    s_ticket = atomicAdd(P.done, 1u);
  }
  __syncthreads();
  if (s_ticket != gridDim.x - 1)
    return;
  __threadfence();
  if (threadIdx.x == 0)
    *P.done = 0u;
  scanMetaBlock(P.planeMeta, nP, P.planePrefix, P.totals);
  if (!P.reportSizes)
    return;
  __syncthreads();
  const unsigned* f = reinterpret_cast<const unsigned*>(P.frame);
  for (unsigned i = threadIdx.x; i < sizeof(FrameDev) / 4u; i += blockDim.x)
    P.reportFrame[i] = __ldcg(f + i);
  if (threadIdx.x < 11)
    P.reportSizes[threadIdx.x] = __ldcg(P.counters + 4 + threadIdx.x);  // ([10] = counters[14]: planes the fast path declined)
  const long long words = 2 * min(nP, static_cast<long long>(P.reportPlanes));
  for (long long i = threadIdx.x; i < words; i += blockDim.x)
    P.reportMeta[i] = __ldcg(P.planeMeta + i);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    P.reportSizes[15] = P.stamp;
}

/** second-generation path (NB2_LINK_GEN=2): two CTAs of 512 threads share an SM and hide each other's barriers */
__global__ void __launch_bounds__(kLinkThreads, 2) k_link_planes(LinkParams Pin) { linkPlanesBody(Pin); }
/** first-generation paths (the default): a plane gets the whole SM — 1024 threads, all of its shared memory — which also
 *  gives the shortest latency per plane, all that counts for the slab of a sharded plan (fewer planes than SMs) */
__global__ void __launch_bounds__(2 * kLinkThreads, 1) k_link_planes_wide(LinkParams Pin) { linkPlanesBody(Pin); }

// =================================================================================================================
// S6: waypoints
// =================================================================================================================
/**
 * One thread per output waypoint; blockIdx.x = slab-local plane, blockIdx.y = 256-waypoint chunk of that plane (the
 * grid covers the 2 n_max waypoints the largest plane could have; chunks past a plane's count exit at once).
 * The waypoint is the cut point its first inserter (triangle rank >> 1, line end rank & 1) produced: contour that
 * triangle against the plane again — same operands, same operation order, hence the same float32 position the linker
 * merged on — and interpolate the vertex normals along the cut edge (vtkDataSetAttributes::InterpolateEdge; only the
 * first inserter's attributes are kept by vtkCutter).  The same kernel packs the per-plane segment lengths.
 */
constexpr unsigned kWpChunk = 1024u;  // waypoints per block and chunk (four per thread)

// (five CTAs per SM: 48 registers; measured 119 us against 125 / 131 / 143 us at four / six / eight on cfg3)
// kRec24: the cloud lies in HBM as 24-byte {x y z nx ny nz} records (what the staged upload leaves there): position and
// normal of a vertex then come out of one record — three 8-byte loads, one or two sectors — instead of a 16-byte position
// and three 4-byte normal components from two arrays (nrmBase points at the record's normal, 12 bytes behind its x)
template <bool kRec24>
__global__ void __launch_bounds__(256, 5)
    k_waypoints(const uint32_t* __restrict__ tri, const float4* __restrict__ pos, const uint8_t* __restrict__ nrmBase,
                unsigned nrmStride, const FrameDev* __restrict__ frame, const unsigned* __restrict__ sizes,
                const unsigned* __restrict__ overflow, const unsigned* __restrict__ planeOff,
                const unsigned* __restrict__ planeMeta, const uint2* __restrict__ prefix,
                const unsigned* __restrict__ wpTag, const unsigned* __restrict__ segLen, float* __restrict__ outPos,
                float* __restrict__ outNrm, uint32_t* __restrict__ outEdgeLo, uint32_t* __restrict__ outEdgeHi,
                unsigned* __restrict__ outSegLen, int emitEdges)
{
  if (*overflow)
    return;
  const LatticeDev L = frame->L;
  const unsigned nP = static_cast<unsigned>(L.plane_end - L.plane_begin);
  // blockIdx.x walks the planes, blockIdx.y the chunks of kWpChunk waypoints of a plane (at most 2 n_max waypoints;
  // sizes[1] = n_max).  The grid was sized from the previous plan, so both loops normally run once; they cover any
  // other size too.  A thread takes kWpChunk / 256 waypoints one after the other, which amortises the block's start-up
  // (the lattice constants come from device memory).
  const unsigned chunks = (2u * sizes[1] + kWpChunk - 1u) / kWpChunk;
  for (unsigned k = blockIdx.x; k < nP; k += gridDim.x)
  {
    const unsigned Wk = __ldg(planeMeta + 2 * k);
    const unsigned Sk = __ldg(planeMeta + 2 * k + 1);
    const unsigned off = __ldg(planeOff + k);
    const uint2 before = __ldg(prefix + k);
    const PlaneOrigin org = planeOrigin(L, L.plane_begin + static_cast<long long>(k));
    for (unsigned ci = blockIdx.y; ci < chunks && ci * kWpChunk < Wk; ci += gridDim.y)
      for (unsigned local = ci * kWpChunk + threadIdx.x; local < min(Wk, (ci + 1u) * kWpChunk); local += 256u)
      {
        if (local < Sk)
          outSegLen[before.y + local] = __ldg(segLen + off + local);
        const unsigned long long i = static_cast<unsigned long long>(before.x) + local;
        const unsigned tg = __ldg(wpTag + 2ull * off + local);
        const unsigned t = tg >> 1;
        const uint32_t ids[3] = { __ldg(tri + 3ull * t), __ldg(tri + 3ull * t + 1), __ldg(tri + 3ull * t + 2) };
        // positions and normals of all three vertices are requested together: the two the cut edge needs are only
        // known after the plane scalars, and waiting for them would put a fourth dependent gather on every waypoint
        float4 p[3], nv[3];
        if constexpr (kRec24)
        {
#pragma unroll
          for (int j = 0; j < 3; ++j)
          {
            const float2* q = reinterpret_cast<const float2*>(nrmBase - 12 + static_cast<size_t>(ids[j]) * 24u);
            const float2 a = __ldg(q), b = __ldg(q + 1), c2 = __ldg(q + 2);
            p[j] = make_float4(a.x, a.y, b.x, 0.f);
            nv[j] = make_float4(b.y, c2.x, c2.y, 0.f);
          }
        }
        else
        {
#pragma unroll
          for (int j = 0; j < 3; ++j)
            p[j] = __ldg(pos + ids[j]);
#pragma unroll
          for (int j = 0; j < 3; ++j)
          {
            const float* q = reinterpret_cast<const float*>(nrmBase + static_cast<size_t>(ids[j]) * nrmStride);
            nv[j] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.f);
          }
        }
        double s[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          s[j] = planeScalar(L, org, p[j].x, p[j].y, p[j].z);
        const CutPoint c = interpolateCut(contourCase(s), tg & 1u, ids, p, s, nv);
        outPos[3 * i] = c.x;
        outPos[3 * i + 1] = c.y;
        outPos[3 * i + 2] = c.z;
        outNrm[3 * i] = c.nx;
        outNrm[3 * i + 1] = c.ny;
        outNrm[3 * i + 2] = c.nz;
        if (emitEdges)
        {
          outEdgeLo[i] = c.lo;
          outEdgeHi[i] = c.hi;
        }
      }
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

size_t smemWorkBytes(unsigned hits)
{
  const size_t E = 2 * static_cast<size_t>(hits);
  size_t cap = 16;
  while (cap < E)
    cap <<= 1;
  return ((12 * E + 15) & ~size_t(15)) + 4 * E + 8 * cap + 2 * E + 16;
}
size_t globalWorkBytes(unsigned hits)
{
  const size_t E = 2 * static_cast<size_t>(hits);
  size_t cap = 16;
  while (cap < E)
    cap <<= 1;
  return ((32 * E + 15) & ~size_t(15)) + 4 * E + 8 * cap + 4 * E + 256;
}
}  // namespace

void launchCountHits(nb2_context* c)
{
  const int grid = gridFor((c->F + 3) / 4, kCountThreads, c->sm_count * 8);
  const int aligned = (reinterpret_cast<uintptr_t>(c->tri.ptr) & 15u) == 0 ? 1 : 0;
  const size_t n = static_cast<size_t>(c->plane_cap) + 1;
  c->planeOff.ensure(sizeof(unsigned) * n);
  c->planeCursor.ensure(sizeof(unsigned) * n);
  if (c->list_quads)
    c->quadList.ensure(sizeof(unsigned) * ((c->F + 3) / 4 + 1));
  // counters[2]: blocks done (the last one scans); {H, n_max} land in counters[4..5]
  if (c->list_quads)
    k_count_hits<true><<<grid, kCountThreads, 0, c->stream>>>(c->tri.as<uint32_t>(), c->F, c->V, c->vertK.as<int>(),
                                                        c->frame.as<FrameDev>(), c->planeCnt.as<unsigned>(), aligned,
                                                        c->counters.as<unsigned>() + 12, c->counters.as<unsigned>() + 2,
                                                        c->planeOff.as<unsigned>(), c->planeCursor.as<unsigned>(),
                                                        c->counters.as<unsigned>() + 4,
                                                        c->list_quads ? c->quadList.as<unsigned>() : nullptr,
                                                        c->counters.as<unsigned>() + 9);
  else
    k_count_hits<false><<<grid, kCountThreads, 0, c->stream>>>(c->tri.as<uint32_t>(), c->F, c->V, c->vertK.as<int>(),
                                                        c->frame.as<FrameDev>(), c->planeCnt.as<unsigned>(), aligned,
                                                        c->counters.as<unsigned>() + 12, c->counters.as<unsigned>() + 2,
                                                        c->planeOff.as<unsigned>(), c->planeCursor.as<unsigned>(),
                                                        c->counters.as<unsigned>() + 4,
                                                        c->list_quads ? c->quadList.as<unsigned>() : nullptr,
                                                        c->counters.as<unsigned>() + 9);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

// The kernels after the hit count read every size they need from device memory (FrameDev, counters), so the host can
// enqueue them without waiting for the count: buffers are sized from capacities (hit_cap hits, scratch_hits hits per
// plane for planes that miss shared memory) and the kernels raise counters[13] when a capacity does not suffice.
void launchEmitHits(nb2_context* c, uint64_t hit_cap)
{
  const int aligned = (reinterpret_cast<uintptr_t>(c->tri.ptr) & 15u) == 0 ? 1 : 0;
  // two waves of four resident CTAs per SM (measured 141 us against 145 us with grids of 8, 16 or 32 CTAs per SM)
  const int grid = gridFor((c->F + 3) / 4, kEmitThreads, c->sm_count * 4);
  k_emit_hits<<<grid, kEmitThreads, 0, c->stream>>>(c->tri.as<uint32_t>(), c->F, c->vertK.as<int>(), c->pos.as<float4>(),
                                                     c->frame.as<FrameDev>(), c->planeOff.as<unsigned>(),
                                                     c->planeCursor.as<unsigned>(), c->hitRec.as<float4>(), aligned,
                                                     c->counters.as<unsigned>() + 4, hit_cap,
                                                     c->counters.as<unsigned>() + 13,
                                                     c->list_quads ? c->quadList.as<unsigned>() : nullptr,
                                                     c->counters.as<unsigned>() + 9);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

/** largest plane (cut lines) whose work area, as sized by `bytes_for`, fits `budget` bytes */
template <typename F>
static unsigned largestPlane(size_t budget, unsigned hi, F bytes_for)
{
  unsigned lo = 0;
  while (lo < hi)
  {
    const unsigned mid = lo + (hi - lo + 1) / 2;
    if (bytes_for(mid) <= budget)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

void launchLinkPlanes(nb2_context* c, uint64_t hit_cap, uint32_t scratch_hits, const LinkReport& report)
{
  // Shared memory per CTA: the second-generation work area for the largest plane expected (the previous plan's largest
  // plane plus a margin; the first plan of a context knows its own), but no more than lets two CTAs share an SM — the
  // planes of two CTAs hide each other's barriers and global-memory phases.  Larger planes get the whole SM; planes beyond
  // even that, and planes the fast path declines, go to the first-generation paths: in shared memory when they fit
  // there, in the CTA's slice of the global scratch otherwise.
  const unsigned n_last = static_cast<unsigned>(std::min<uint64_t>(link2::kMaxLines, c->grid_nmax));
  const unsigned n_hint = static_cast<unsigned>(std::min<uint64_t>(link2::kMaxLines, static_cast<uint64_t>(n_last) + n_last / 8 + 64));
  const size_t per_sm = c->smem_per_sm ? c->smem_per_sm : (size_t(228) << 10);
  const size_t two_per_sm = per_sm / 2 - 1024 - 512;  // (1 KB reserved per CTA, static shared memory)
  const size_t optin = c->smem_optin > 2048 ? c->smem_optin - 2048 : 0;
  // Which fast path, and with it the launch shape.  Measured on cfg3 (profiles/r02_summary.md), link time in us for the whole
  // lattice / a half / a third / an eighth of it (the slabs of 1, 2, 3, 8 GPUs):
  //   first generation, one 1024-thread CTA per SM with all its shared memory   222 / 137 / 107 /  45
  //   second generation, two 512-thread CTAs per SM                             226 / 160 / 152 / 158
  // The second generation's plane takes twice as long (more, cheaper phases) and only breaks even when two planes overlap on
  // every SM, so the first generation is the default; NB2_LINK_GEN=2 selects the second (its phases are the ones
  // tests/cpp/link_emulation.cpp replays on the host).
  static const int link_gen = [] {
    const char* e = std::getenv("NB2_LINK_GEN");
    return e && std::atoi(e) == 2 ? 2 : 1;
  }();
  const bool wide = link_gen == 1;
  size_t smem_bytes = std::max<size_t>(link2::workBytes(n_hint), size_t(16) << 10);
  if (wide)
    smem_bytes = optin;
  else if (smem_bytes > two_per_sm)
  {
    // the margin must not cost the second CTA: when the last plan's largest plane fits two to an SM, that is the budget
    smem_bytes = link2::workBytes(n_last + 16) <= two_per_sm ? two_per_sm : std::min(smem_bytes, optin);
  }
  smem_bytes = std::min(smem_bytes, optin);
  const unsigned fast_cap = largestPlane(smem_bytes, link2::kMaxLines, [](unsigned n) { return link2::workBytes(n); });
  const unsigned smem_cap = largestPlane(smem_bytes, 32767u, [](unsigned n) { return smemWorkBytes(n); });
  // (planes of a few hundred cut lines do not fill 512 threads: smaller CTAs, more of them per SM)
  const int threads = wide ? 2 * kLinkThreads : (n_hint >= 2048u ? kLinkThreads : kLinkThreads / 2);
  // attribute and occupancy are asked for when the configuration changes, not at every launch (host time between the
  // launches of a plan is what a sharded plan's short kernels cannot hide)
  int per_sm_ctas = 1;
  if (c->link_cfg_threads != threads || c->link_cfg_smem != smem_bytes || c->link_cfg_wide != wide)
  {
    if (wide)
    {
      NB2_CUDA(cudaFuncSetAttribute(k_link_planes_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
      NB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_ctas, k_link_planes_wide, threads, smem_bytes));
    }
    else
    {
      NB2_CUDA(cudaFuncSetAttribute(k_link_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
      NB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_ctas, k_link_planes, threads, smem_bytes));
    }
    c->link_cfg_threads = threads;
    c->link_cfg_smem = smem_bytes;
    c->link_cfg_wide = wide;
    c->link_cfg_ctas = per_sm_ctas;
  }
  per_sm_ctas = c->link_cfg_ctas;
  int grid = c->sm_count * std::max(1, per_sm_ctas);
  // global scratch of the first-generation paths: any plane may end up there
  const uint32_t scratch_n = std::max<uint32_t>(scratch_hits, n_hint);
  size_t stride = 0;
  if (scratch_n > smem_cap)
  {
    stride = (globalWorkBytes(scratch_n) + 255) & ~size_t(255);
    while (grid > c->sm_count && stride * grid > (size_t(8) << 30))
      grid -= c->sm_count;
    c->linkScratch.ensure(stride * grid);
  }
  const size_t planes = static_cast<size_t>(c->plane_cap) + 1;
  c->planeMeta.ensure(sizeof(unsigned) * 2 * planes);
  c->planePrefix.ensure(sizeof(uint2) * planes);
  c->wpTag.ensure(sizeof(unsigned) * 2 * hit_cap);
  c->segLen.ensure(sizeof(unsigned) * hit_cap);

  LinkParams P{};
  P.rec = c->hitRec.as<float4>();
  P.planeOff = c->planeOff.as<unsigned>();
  P.frame = c->frame.as<FrameDev>();
  P.ticket = c->counters.as<unsigned>() + 8;
  P.planeMeta = c->planeMeta.as<unsigned>();
  P.wpTag = c->wpTag.as<unsigned>();
  P.segLen = c->segLen.as<unsigned>();
  P.status = c->counters.as<unsigned>() + 10;
  P.scratch = c->linkScratch.as<uint8_t>();
  P.scratchStride = stride;
  P.smemHitCap = smem_cap;
  P.scratchHits = scratch_n > smem_cap ? scratch_n : 0u;
  P.overflow = c->counters.as<unsigned>() + 13;
  P.fastHitCap = fast_cap;
  static const unsigned plain_rounds = [] {
    const char* e = std::getenv("NB2_LINK_PLAIN_ROUNDS");
    return e ? static_cast<unsigned>(std::max(1, std::atoi(e))) : 1u;
  }();
  P.plainRounds = plain_rounds;
  P.generalCount = c->counters.as<unsigned>() + 14;
  static const bool profile = std::getenv("NB2_LINK_PROFILE") != nullptr;
  if (profile)
  {
    c->linkProfile.ensure(16 * sizeof(unsigned long long));
    NB2_CUDA(cudaMemsetAsync(c->linkProfile.ptr, 0, 16 * sizeof(unsigned long long), c->stream));
    P.phaseCycles = c->linkProfile.as<unsigned long long>();
  }
  P.done = c->counters.as<unsigned>() + 3;
  P.planePrefix = c->planePrefix.as<uint2>();
  P.totals = c->counters.as<unsigned>() + 6;
  P.counters = c->counters.as<unsigned>();
  P.reportFrame = report.frame;
  P.reportSizes = report.sizes;
  P.reportMeta = report.meta;
  P.reportPlanes = report.planes;
  P.stamp = report.stamp;

  if (link_gen == 1)
    P.fastHitCap = 0;  // no plane is offered to the second-generation path
  if (wide)
    k_link_planes_wide<<<grid, threads, smem_bytes, c->stream>>>(P);
  else
    k_link_planes<<<grid, threads, smem_bytes, c->stream>>>(P);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchWaypoints(nb2_context* c, uint64_t w_cap, int emit_edges)
{
  // grid from the sizes of the previous plan on this context (the kernel loops when this plan is larger)
  unsigned gx = c->grid_planes ? c->grid_planes : 1024u, gy = (2u * c->grid_nmax + kWpChunk - 1u) / kWpChunk;
  gx = std::min(gx, 1u << 20);
  gy = std::max(1u, std::min(gy, 4096u));
  // the cloud as 24-byte {xyz, normal} records whose normals are the ones in use (not replaced by computed / set normals)
  const bool rec24 = c->rec_step == 24 && c->rec_off_xyz == 0 && c->rec_off_nrm == 12 && c->nrm_stride == 24 &&
                     c->nrm_base == c->blob.as<uint8_t>() + 12 && (reinterpret_cast<uintptr_t>(c->blob.ptr) & 7u) == 0;
  auto launch = [&](auto kernel) {
    kernel<<<dim3(gx, gy), 256, 0, c->stream>>>(
        c->tri.as<uint32_t>(), c->pos.as<float4>(), c->nrm_base, c->nrm_stride, c->frame.as<FrameDev>(),
        c->counters.as<unsigned>() + 4, c->counters.as<unsigned>() + 13, c->planeOff.as<unsigned>(),
        c->planeMeta.as<unsigned>(), c->planePrefix.as<uint2>(), c->wpTag.as<unsigned>(), c->segLen.as<unsigned>(),
        c->outPos.as<float>(), c->outNrm.as<float>(), c->outEdge.as<uint32_t>(), c->outEdge.as<uint32_t>() + w_cap,
        c->outSegLen.as<unsigned>(), emit_edges);
  };
  if (rec24)
    launch(k_waypoints<true>);
  else
    launch(k_waypoints<false>);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

}  // namespace nb2
