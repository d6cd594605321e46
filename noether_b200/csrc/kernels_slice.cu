// The vtkCutter + vtkStripper + processPolyLines replacement (sm_100a).
//
// The reference loops over planes and lets vtkCutter visit every vertex and every triangle per plane
// (plane_slicer_raster_planner.cpp:596-628).  Here the loop is inverted:
//
//   S2  k_count_hits   : each triangle projects onto the plane lattice through the per-vertex indices K_v
//                        (k_classify) and bumps a per-plane hit counter
//   S3  k_scan_planes  : exclusive scan of the per-plane counters -> bucket offsets (+ total hits, largest plane)
//   S4  k_fill_buckets : each (triangle, plane) hit drops its triangle id into the plane's bucket
//   S5  k_link_planes  : one persistent CTA per plane: contour every hit exactly as vtkTriangle::Contour does,
//                        merge cut points exactly as vtkMergePoints does (float-coordinate equality, staged in a
//                        shared-memory hash), pair the lines at shared points, rank the chains by pointer jumping,
//                        orient and order them, size the output through a decoupled look-back across planes, and
//                        write ordered waypoints (position, interpolated normal, generating mesh edge)
//
// Output of S5 is the reference's planImpl result before convertToPoses expands it to 4x4 matrices.
#include "device_math.cuh"
#include "nb2_internal.h"

#include <cfloat>

namespace nb2
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;
constexpr int kCountThreads = 256;
constexpr int kSmemBins = 8192;  // per-block privatised plane histogram (32 KB)

__device__ __forceinline__ void triangleSpan(const uint32_t* __restrict__ tri, const int* __restrict__ vertK,
                                             unsigned long long t, const LatticeDev& L, long long& k0, long long& k1)
{
  const uint32_t a = __ldg(tri + 3 * t), b = __ldg(tri + 3 * t + 1), c = __ldg(tri + 3 * t + 2);
  const int ka = __ldg(vertK + a), kb = __ldg(vertK + b), kc = __ldg(vertK + c);
  const int lo = min(ka, min(kb, kc)), hi = max(ka, max(kb, kc));
  k0 = max(static_cast<long long>(lo) + 1, L.plane_begin);
  k1 = min(static_cast<long long>(hi), L.plane_end - 1);
}

__global__ void __launch_bounds__(kCountThreads)
    k_count_hits(const uint32_t* __restrict__ tri, unsigned long long F, const int* __restrict__ vertK, LatticeDev L,
                 unsigned* __restrict__ planeCnt, int use_smem)
{
  __shared__ unsigned s_bins[kSmemBins];
  const long long nP = L.plane_end - L.plane_begin;
  if (use_smem)
  {
    for (int i = threadIdx.x; i < nP; i += blockDim.x)
      s_bins[i] = 0;
    __syncthreads();
  }
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long t = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < F; t += stride)
  {
    long long k0, k1;
    triangleSpan(tri, vertK, t, L, k0, k1);
    for (long long k = k0; k <= k1; ++k)
    {
      if (use_smem)
        atomicAdd(&s_bins[k - L.plane_begin], 1u);
      else
        atomicAdd(&planeCnt[k - L.plane_begin], 1u);
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
}

/** Single-block exclusive scan of the per-plane counters.  out[0] = total hits, out[1] = largest plane. */
__global__ void __launch_bounds__(1024)
    k_scan_planes(const unsigned* __restrict__ cnt, long long nP, unsigned* __restrict__ off, unsigned* __restrict__ cursor,
                  unsigned long long* __restrict__ lookback, unsigned* __restrict__ out2)
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
    const unsigned v = (i < nP) ? cnt[i] : 0u;
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
      lookback[i] = 0ull;
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

__global__ void __launch_bounds__(kCountThreads)
    k_fill_buckets(const uint32_t* __restrict__ tri, unsigned long long F, const int* __restrict__ vertK, LatticeDev L,
                   const unsigned* __restrict__ planeOff, unsigned* __restrict__ cursor, uint32_t* __restrict__ bucket)
{
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long t = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < F; t += stride)
  {
    long long k0, k1;
    triangleSpan(tri, vertK, t, L, k0, k1);
    for (long long k = k0; k <= k1; ++k)
    {
      const long long b = k - L.plane_begin;
      const unsigned slot = __ldg(planeOff + b) + atomicAdd(&cursor[b], 1u);
      bucket[slot] = static_cast<uint32_t>(t);
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
  const uint32_t* tri;
  const float4* pos;
  const float4* nrm;
  const unsigned* planeOff;
  const uint32_t* bucket;
  LatticeDev L;
  unsigned long long* lookback;
  unsigned* ticket;
  unsigned* planeMeta;  // [2 * nP] waypoints, segments
  float* outPos;
  float* outNrm;
  uint32_t* outEdgeLo;
  uint32_t* outEdgeHi;
  unsigned* outSegLen;
  unsigned* status;  // [0] error bits, [1] first failing plane
  uint8_t* scratch;  // global work areas for planes that do not fit shared memory
  unsigned long long scratchStride;
  unsigned smemHitCap;  // largest plane (in hits) that fits the shared-memory work area
  unsigned long long wCap, sCap;
  int emitEdges;
};

constexpr unsigned kErrNonManifold = 1u, kErrClosedLoop = 2u, kErrOverflow = 4u;
// pairing cell of a hash slot: [31:20] number of line ends at the point, [19:0] last line end that arrived
constexpr unsigned kCellShift = 20u, kCellMask = (1u << kCellShift) - 1u;
constexpr unsigned kMaxCrowded = 16u;      // points with more than two incident lines handled per plane
constexpr unsigned kMaxCrowdedEnds = 32u;  // line ends at one such point
// s_misc: [0] chains, [1] error bits, [2] baseW, [3] baseS, [4] plane waypoints, [5] ranking flag, [6] crowded points,
//         [7] closed loops present, [8..] crowded slot ids
constexpr int kMiscWords = 8 + kMaxCrowded;

// look-back status word: [63:62] flag, [61:31] waypoints, [30:0] segments
constexpr unsigned long long kFlagAggregate = 1ull << 62, kFlagPrefix = 2ull << 62;
__device__ __forceinline__ unsigned long long packStatus(unsigned long long flag, unsigned long long w, unsigned long long s)
{
  return flag | (w << 31) | s;
}

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
  float* epos;     // [3 * E]   endpoint positions (phase A/H); later ping-pong ranking buffers
  unsigned* etag;  // [E]       slot index, later the merged point's first-inserter rank
  Idx* twin;       // [E]       endpoint sharing the same cut point in another line
  Slot* slots;     // [cap]     hash of cut points; later chain heads
  unsigned cap;    // power of two >= E
};

template <typename Idx>
__host__ __device__ inline size_t planeWorkBytes(unsigned hits)
{
  const size_t E = 2 * static_cast<size_t>(hits);
  size_t cap = 16;
  while (cap < E)
    cap <<= 1;
  const size_t a = 12 * E;  // epos; ranking needs 2 buffers x 3 arrays x sizeof(Idx) x E <= 12 E for Idx <= 2 bytes
  const size_t rank = 6 * sizeof(Idx) * E;
  return (a > rank ? a : rank) + 4 * E + sizeof(Idx) * E + 8 * cap + 64;
}

/** The triangle of first-inserter rank `tag` contoured against plane k: endpoint tag & 1 */
__device__ __forceinline__ CutPoint cutPointOfTag(const LinkParams& P, const PlaneOrigin& o, unsigned tag, bool with_normal)
{
  const unsigned t = tag >> 1;
  uint32_t ids[3] = { __ldg(P.tri + 3ull * t), __ldg(P.tri + 3ull * t + 1), __ldg(P.tri + 3ull * t + 2) };
  float4 p[3] = { __ldg(P.pos + ids[0]), __ldg(P.pos + ids[1]), __ldg(P.pos + ids[2]) };
  double s[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    s[i] = planeScalar(P.L, o, p[i].x, p[i].y, p[i].z);
  return interpolateCut(contourCase(s), tag & 1, ids, p, s, with_normal ? P.nrm : nullptr);
}

template <typename Idx>
__device__ void linkPlane(const LinkParams& P, const PlaneWork<Idx>& W, long long kLocal, unsigned n, unsigned* s_misc)
{
  constexpr Idx kNone = static_cast<Idx>(~static_cast<Idx>(0));
  const unsigned E = 2 * n;
  const unsigned tid = threadIdx.x, nth = blockDim.x;
  const long long k = P.L.plane_begin + kLocal;
  const PlaneOrigin org = planeOrigin(P.L, k);
  const unsigned off = P.planeOff[kLocal];
  const unsigned mask = W.cap - 1;

  // s_misc: [0] nHeads, [1] local error bits, [2] baseW, [3] baseS, [4] totalW, [5] any-unfinished flag
  if (tid < kMiscWords)
    s_misc[tid] = 0;

  // ---- A: contour every hit (vtkTriangle::Contour) ----------------------------------------------------------------
  for (unsigned j = tid; j < n; j += nth)
  {
    const unsigned t = __ldg(P.bucket + off + j);
    uint32_t ids[3] = { __ldg(P.tri + 3ull * t), __ldg(P.tri + 3ull * t + 1), __ldg(P.tri + 3ull * t + 2) };
    float4 p[3] = { __ldg(P.pos + ids[0]), __ldg(P.pos + ids[1]), __ldg(P.pos + ids[2]) };
    double s[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      s[i] = planeScalar(P.L, org, p[i].x, p[i].y, p[i].z);
    const int index = contourCase(s);
#pragma unroll
    for (int w = 0; w < 2; ++w)
    {
      const CutPoint c = interpolateCut(index, w, ids, p, s, nullptr);
      float* e = W.epos + 3 * (2 * j + w);
      e[0] = c.x;
      e[1] = c.y;
      e[2] = c.z;
    }
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
    const unsigned tag = 2u * __ldg(P.bucket + off + (e >> 1)) + (e & 1u);
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
        const unsigned tg = 2u * __ldg(P.bucket + off + (e >> 1)) + (e & 1u);
        unsigned j = cnt++;
        while (j > 0)
        {
          const unsigned o = mem[j - 1];
          const unsigned to = 2u * __ldg(P.bucket + off + (o >> 1)) + (o & 1u);
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

  // ---- R: rank the darts of every chain by pointer jumping (Wyllie), both directions ---------------------------------
  // dart e = "traverse line e>>1 starting at end e"; next(e) = twin(e ^ 1)
  Idx* buf[2][3];
  {
    Idx* base = reinterpret_cast<Idx*>(W.epos);
    for (int b = 0; b < 2; ++b)
      for (int a = 0; a < 3; ++a)
        buf[b][a] = base + (static_cast<size_t>(b) * 3 + a) * E;
  }
  for (unsigned e = tid; e < E; e += nth)
  {
    const bool dead = (W.twin[e] == static_cast<Idx>(e));
    Idx nx = kNone;
    if (!dead)
      nx = W.twin[e ^ 1u];
    buf[0][0][e] = nx;
    buf[0][1][e] = static_cast<Idx>(nx != kNone ? 1 : 0);
    buf[0][2][e] = static_cast<Idx>(e);
  }
  __syncthreads();
  int cur = 0;
  unsigned rounds = 1;
  while ((1u << rounds) < n + 1)
    ++rounds;
  ++rounds;
  for (unsigned r = 0; r < rounds; ++r)
  {
    Idx *nxt = buf[cur][0], *rnk = buf[cur][1], *lst = buf[cur][2];
    Idx *nxt2 = buf[cur ^ 1][0], *rnk2 = buf[cur ^ 1][1], *lst2 = buf[cur ^ 1][2];
    unsigned pending = 0;
    for (unsigned e = tid; e < E; e += nth)
    {
      const Idx nx = nxt[e];
      if (nx != kNone)
      {
        const Idx nn = nxt[nx];
        nxt2[e] = nn;
        rnk2[e] = static_cast<Idx>(rnk[e] + rnk[nx]);
        lst2[e] = lst[nx];
        pending |= (nn != kNone);
      }
      else
      {
        nxt2[e] = kNone;
        rnk2[e] = rnk[e];
        lst2[e] = lst[e];
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
  Idx *nxt = buf[cur][0], *rnk = buf[cur][1], *lst = buf[cur][2];
  Idx* chainEnd = buf[cur ^ 1][0];  // dead buffer, reused
  {
    unsigned bad = 0;
    for (unsigned e = tid; e < E; e += nth)
    {
      bad |= (nxt[e] != kNone);
      chainEnd[e] = kNone;
    }
    if (bad)
      s_misc[7] = 1u;  // some darts never reached an end: closed loops
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
  auto endTag = [&](unsigned e) { return 2u * __ldg(P.bucket + off + (e >> 1)) + (e & 1u); };

  if (s_misc[7])
  {
    // Closed loops are rare (the reference targets open, roughly planar surfaces): one thread walks them.  A loop
    // starts at end 0 of its lowest-tag line and first walks that line, as vtkStripper does from its seed; it is
    // reversed, keeping the start point, when that first step points against the cut direction.
    if (tid == 0)
    {
      for (unsigned e0 = 0; e0 < E; ++e0)
      {
        if (nxt[e0] == kNone)
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
        CutPoint a{};
        if ((best & 1u) == 0u)
        {
          a = cutPointOfTag(P, org, W.etag[c], false);
          const CutPoint b = cutPointOfTag(P, org, W.etag[c ^ 1u], false);
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
              nxt[d] = kNone, rnk[d] = static_cast<Idx>(len - 1 - i), lst[d] = static_cast<Idx>(tailF);
              nxt[m] = kNone, rnk[m] = 0, lst[m] = static_cast<Idx>(m);
            }
            else
            {
              nxt[d] = kNone, rnk[d] = 0, lst[d] = static_cast<Idx>(d);
              nxt[m] = kNone, rnk[m] = static_cast<Idx>(i), lst[m] = static_cast<Idx>(c ^ 1u);
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
        if (nxt[e0] != kNone)
          nxt[e0] = kNone, rnk[e0] = 0, lst[e0] = static_cast<Idx>(e0);
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
    const CutPoint a = cutPointOfTag(P, org, W.etag[e], false);
    const CutPoint b = cutPointOfTag(P, org, W.etag[tail ^ 1u], false);
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
  unsigned* headOrder = reinterpret_cast<unsigned*>(buf[cur ^ 1][1]);  // dead ranking buffer: room for n entries
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

  // ---- L: decoupled look-back over planes: where do this plane's waypoints / segments start? -----------------------
  if (tid == 0)
  {
    unsigned long long exclW = 0, exclS = 0;
    if (kLocal > 0)
    {
      *reinterpret_cast<volatile unsigned long long*>(&P.lookback[kLocal]) = packStatus(kFlagAggregate, Wk, S);
      long long j = kLocal - 1;
      while (true)
      {
        unsigned long long st;
        do
        {
          st = *reinterpret_cast<volatile unsigned long long*>(&P.lookback[j]);
        } while ((st >> 62) == 0ull);
        exclW += (st >> 31) & 0x7fffffffull;
        exclS += st & 0x7fffffffull;
        if ((st >> 62) == 2ull)
          break;
        --j;
      }
    }
    *reinterpret_cast<volatile unsigned long long*>(&P.lookback[kLocal]) = packStatus(kFlagPrefix, exclW + Wk, exclS + S);
    s_misc[2] = static_cast<unsigned>(exclW);
    s_misc[3] = static_cast<unsigned>(exclS);
    P.planeMeta[2 * kLocal] = Wk;
    P.planeMeta[2 * kLocal + 1] = S;
    if (exclW + Wk > P.wCap || exclS + S > P.sCap)
      atomicOr(&s_misc[1], kErrOverflow);
  }
  __syncthreads();
  if (s_misc[1] & kErrOverflow)
    return;
  const unsigned long long baseW = s_misc[2], baseS = s_misc[3];

  // ---- O: output ----------------------------------------------------------------------------------------------------
  for (unsigned h = tid; h < S; h += nth)
  {
    const unsigned ord = headOrder[h];
    const unsigned e = heads[h].e;
    const unsigned lines = static_cast<unsigned>(rnk[e]) + 1u;
    chainEnd[lst[e]] = static_cast<Idx>(lenSorted[ord] + lines);  // plane-local index of the chain's last point
    P.outSegLen[baseS + ord] = lines + 1u;
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
    for (int last = 0; last < 2; ++last)
    {
      if (last && r != 0)
        break;
      const unsigned tag = last ? W.etag[e ^ 1u] : W.etag[e];
      const unsigned long long idx = baseW + (last ? static_cast<unsigned>(ce) : static_cast<unsigned>(ce) - 1u - r);
      const CutPoint c = cutPointOfTag(P, org, tag, true);
      P.outPos[3 * idx] = c.x;
      P.outPos[3 * idx + 1] = c.y;
      P.outPos[3 * idx + 2] = c.z;
      P.outNrm[3 * idx] = c.nx;
      P.outNrm[3 * idx + 1] = c.ny;
      P.outNrm[3 * idx + 2] = c.nz;
      if (P.emitEdges)
      {
        P.outEdgeLo[idx] = c.lo;
        P.outEdgeHi[idx] = c.hi;
      }
    }
  }
}

__global__ void __launch_bounds__(512, 1) k_link_planes(LinkParams P)
{
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ unsigned s_misc[kMiscWords];
  __shared__ unsigned s_ticket;
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

    if (n <= P.smemHitCap)
    {
      const unsigned E = 2 * n;
      unsigned cap = 16;
      while (cap < E)
        cap <<= 1;
      PlaneWork<uint16_t> W;
      const size_t a = 12ull * E;
      W.epos = reinterpret_cast<float*>(smem);
      W.etag = reinterpret_cast<unsigned*>(smem + ((a + 15) & ~size_t(15)));
      W.slots = reinterpret_cast<Slot*>(reinterpret_cast<uint8_t*>(W.etag) + 4ull * E);
      W.twin = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(W.slots) + 8ull * cap);
      W.cap = cap;
      linkPlane<uint16_t>(P, W, kLocal, n, s_misc);
    }
    else
    {
      const unsigned E = 2 * n;
      unsigned cap = 16;
      while (cap < E)
        cap <<= 1;
      uint8_t* g = P.scratch + P.scratchStride * blockIdx.x;
      PlaneWork<uint32_t> W;
      const size_t a = 24ull * E;  // 2 buffers x 3 arrays x 4 B
      W.epos = reinterpret_cast<float*>(g);
      W.etag = reinterpret_cast<unsigned*>(g + ((a + 15) & ~size_t(15)));
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
      // a failed plane must still publish a prefix, or later planes would spin forever
      unsigned long long exclW = 0, exclS = 0;
      const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&P.lookback[kLocal]);
      if ((cur >> 62) != 2ull)
      {
        long long j = kLocal - 1;
        while (j >= 0)
        {
          unsigned long long st;
          do
          {
            st = *reinterpret_cast<volatile unsigned long long*>(&P.lookback[j]);
          } while ((st >> 62) == 0ull);
          exclW += (st >> 31) & 0x7fffffffull;
          exclS += st & 0x7fffffffull;
          if ((st >> 62) == 2ull)
            break;
          --j;
        }
        *reinterpret_cast<volatile unsigned long long*>(&P.lookback[kLocal]) = packStatus(kFlagPrefix, exclW, exclS);
        P.planeMeta[2 * kLocal] = 0;
        P.planeMeta[2 * kLocal + 1] = 0;
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
  return ((24 * E + 15) & ~size_t(15)) + 4 * E + 8 * cap + 4 * E + 256;
}
}  // namespace

void launchCountHits(nb2_context* c, const LatticeDev& L)
{
  const long long nP = L.plane_end - L.plane_begin;
  c->planeCnt.ensure(sizeof(unsigned) * (nP + 1));
  NB2_CUDA(cudaMemsetAsync(c->planeCnt.ptr, 0, sizeof(unsigned) * (nP + 1), c->stream));
  const int grid = gridFor(c->F, kCountThreads, c->sm_count * 8);
  k_count_hits<<<grid, kCountThreads, 0, c->stream>>>(c->tri.as<uint32_t>(), c->F, c->vertK.as<int>(), L,
                                                      c->planeCnt.as<unsigned>(), nP <= kSmemBins ? 1 : 0);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchScanPlanes(nb2_context* c, const LatticeDev& L)
{
  const long long nP = L.plane_end - L.plane_begin;
  c->planeOff.ensure(sizeof(unsigned) * (nP + 1));
  c->planeCursor.ensure(sizeof(unsigned) * (nP + 1));
  c->lookback.ensure(sizeof(unsigned long long) * (nP + 1));
  k_scan_planes<<<1, 1024, 0, c->stream>>>(c->planeCnt.as<unsigned>(), nP, c->planeOff.as<unsigned>(),
                                           c->planeCursor.as<unsigned>(), c->lookback.as<unsigned long long>(),
                                           c->counters.as<unsigned>() + 4);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchFillBuckets(nb2_context* c, const LatticeDev& L)
{
  const int grid = gridFor(c->F, kCountThreads, c->sm_count * 8);
  k_fill_buckets<<<grid, kCountThreads, 0, c->stream>>>(c->tri.as<uint32_t>(), c->F, c->vertK.as<int>(), L,
                                                        c->planeOff.as<unsigned>(), c->planeCursor.as<unsigned>(),
                                                        c->bucket.as<uint32_t>());
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

void launchLinkPlanes(nb2_context* c, const LatticeDev& L, uint32_t n_max, uint64_t w_cap, uint64_t s_cap, int emit_edges)
{
  const long long nP = L.plane_end - L.plane_begin;
  // largest plane that fits the opt-in shared memory (uint16 indices additionally need 2 n < 65535)
  const size_t smem_budget = c->smem_optin > 1024 ? c->smem_optin - 1024 : 0;
  unsigned hit_cap = 0;
  {
    unsigned lo = 0, hi = 32767;  // uint16 indices need 2 n < 65535
    while (lo < hi)
    {
      const unsigned mid = (lo + hi + 1) / 2;
      if (smemWorkBytes(mid) <= smem_budget)
        lo = mid;
      else
        hi = mid - 1;
    }
    hit_cap = lo;
  }
  const unsigned smem_hits = n_max < hit_cap ? n_max : hit_cap;
  const size_t smem_bytes = smemWorkBytes(smem_hits > 0 ? smem_hits : 1);

  int grid = c->sm_count;
  if (grid > nP)
    grid = static_cast<int>(nP > 0 ? nP : 1);
  size_t stride = 0;
  if (n_max > hit_cap)
  {
    stride = (globalWorkBytes(n_max) + 255) & ~size_t(255);
    c->linkScratch.ensure(stride * grid);
  }
  c->planeMeta.ensure(sizeof(unsigned) * 2 * (nP + 1));

  LinkParams P{};
  P.tri = c->tri.as<uint32_t>();
  P.pos = c->pos.as<float4>();
  P.nrm = c->nrm.as<float4>();
  P.planeOff = c->planeOff.as<unsigned>();
  P.bucket = c->bucket.as<uint32_t>();
  P.L = L;
  P.lookback = c->lookback.as<unsigned long long>();
  P.ticket = c->counters.as<unsigned>() + 8;
  P.planeMeta = c->planeMeta.as<unsigned>();
  P.outPos = c->outPos.as<float>();
  P.outNrm = c->outNrm.as<float>();
  P.outEdgeLo = c->outEdge.as<uint32_t>();
  P.outEdgeHi = c->outEdge.as<uint32_t>() + w_cap;
  P.outSegLen = c->outSegLen.as<unsigned>();
  P.status = c->counters.as<unsigned>() + 10;
  P.scratch = c->linkScratch.as<uint8_t>();
  P.scratchStride = stride;
  P.smemHitCap = hit_cap;
  P.wCap = w_cap;
  P.sCap = s_cap;
  P.emitEdges = emit_edges;

  NB2_CUDA(cudaFuncSetAttribute(k_link_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  k_link_planes<<<grid, 512, smem_bytes, c->stream>>>(P);
  NB2_CUDA(cudaGetLastError());
  ++c->launches;
}

}  // namespace nb2
