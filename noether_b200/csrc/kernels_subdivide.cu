// Face subdivision mesh modifiers on the device: FaceMidpointSubdivisionMeshModifier and
// FaceSubdivisionByAreaMeshModifier (face_subdivision_modifier.cpp:202-339) applied to the resident cloud and triangle
// list.  The reference walks the faces depth first on one core (a std::map lookup per edge, a std::vector growth per
// vertex); here every level of the subdivision trees is one set of streaming launches and the reference's vertex and
// face numbering is recovered afterwards (subdivide_ops.h explains how; the per-element work lives there and is shared
// with the host emulation the CPU tests compare against the oracle).
//
//   midpoint, per level l = 1 .. n      S1 k_mid_insert   3 edge requests per triangle -> hash slot keeps min sequence no.
//                                       S2 k_mid_resolve  provisional midpoint ids, creator mask per node, 4 children
//             once                      scan of the creator counts (pre-order node index)  -> final vertex numbers
//             per level                 S3 k_mid_create   records of the created vertices (mean of the two sources)
//             once                      S4 k_mid_remap    output triangles: provisional -> final ids
//   by area,  per level                 A1 k_area_test    area > max_area ?   -> scan -> splits of the level (read back)
//                                       A2 k_area_split   centroid records + 3 children per split triangle
//             per level, bottom-up      A3 k_area_up      leaves / centroids below every triangle
//             per level, top-down       A4 k_area_down    depth-first offsets, final centroid numbers
//             per level                 A5 k_area_emit    kept triangles to their place, final ids
//             once                      A6 k_area_place   centroid records to their final place
//
// All kernels are grid-stride streaming passes (HBM-bound: 12 B per triangle corner triple in, 48 B out per level for the
// midpoint scheme) sized to a multiple of the SM count.
#include "nb2_internal.h"
#include "subdivide_ops.h"

#include <algorithm>
#include <utility>

namespace nb2
{
namespace
{
using namespace sub;

struct DevCas
{
  __device__ unsigned long long operator()(unsigned long long* addr, unsigned long long expect, unsigned long long val) const
  {
    const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(addr);
    if (seen != expect)
      return seen;  // keys never change once written
    return atomicCAS(addr, expect, val);
  }
};
struct DevMin
{
  __device__ void operator()(unsigned* addr, unsigned val) const { atomicMin(addr, val); }
};

#define NB2_GRID_STRIDE(i, n)                                                                                          \
  for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x,                 \
                          stride__ = static_cast<unsigned long long>(gridDim.x) * blockDim.x;                          \
       i < (n); i += stride__)

__global__ void __launch_bounds__(256)
    k_mid_insert(MidpointLevel lv, unsigned long long* table, unsigned slot_mask, unsigned* __restrict__ req_slot)
{
  NB2_GRID_STRIDE(j, lv.n_nodes) midpointInsert(lv, j, table, slot_mask, req_slot, DevCas{}, DevMin{});
}

__global__ void __launch_bounds__(256)
    k_mid_resolve(MidpointLevel lv, const unsigned long long* __restrict__ table, const unsigned* __restrict__ req_slot,
                  uint8_t* __restrict__ node_mask, unsigned* __restrict__ node_cnt, uint32_t* __restrict__ children)
{
  NB2_GRID_STRIDE(j, lv.n_nodes) midpointResolve(lv, j, table, req_slot, node_mask, node_cnt, children);
}

__global__ void __launch_bounds__(256)
    k_mid_create(MidpointLevel lv, const unsigned* __restrict__ node_base, const uint8_t* __restrict__ node_mask, uint8_t* cloud,
                 Layout L)
{
  NB2_GRID_STRIDE(j, lv.n_nodes) midpointCreate(lv, j, node_base, node_mask, cloud, L);
}

__global__ void __launch_bounds__(256)
    k_mid_remap(uint32_t* __restrict__ ids, unsigned long long n, uint32_t V, const unsigned* __restrict__ node_base,
                const uint8_t* __restrict__ node_mask)
{
  NB2_GRID_STRIDE(i, n) ids[i] = midpointFinal(ids[i], V, node_base, node_mask);
}

/** level 0 of the by-area sweep: the faces in the order the stack pops them (last face first) */
__global__ void __launch_bounds__(256)
    k_area_roots(const uint32_t* __restrict__ tri, unsigned long long F, uint32_t* __restrict__ out)
{
  NB2_GRID_STRIDE(i, 3 * F)
  {
    const unsigned long long r = i / 3, k = i - 3 * r;
    out[i] = tri[3 * (F - 1 - r) + k];
  }
}

__global__ void __launch_bounds__(256)
    k_area_test(const uint32_t* __restrict__ tri, unsigned long long n, uint32_t V, const uint8_t* __restrict__ cloud,
                const uint8_t* __restrict__ fresh, Layout L, float max_area, unsigned* __restrict__ split,
                unsigned* __restrict__ non_finite)
{
  NB2_GRID_STRIDE(j, n + 1)
  {
    if (j == n)
    {
      split[j] = 0u;  // the scan's total lands here
      continue;
    }
    const unsigned r = areaTest(tri, j, V, cloud, fresh, L, max_area);
    if (r == 2u)
      atomicAdd(non_finite, 1u);
    split[j] = r == 1u ? 1u : 0u;
  }
}

__global__ void __launch_bounds__(256)
    k_area_split(const uint32_t* __restrict__ tri, unsigned long long n, const unsigned* __restrict__ split,
                 const unsigned* __restrict__ pre, uint32_t base, uint32_t V, const uint8_t* __restrict__ cloud, uint8_t* fresh,
                 Layout L, uint32_t* __restrict__ children)
{
  NB2_GRID_STRIDE(j, n) if (split[j]) areaSplit(tri, j, pre[j], base, V, cloud, fresh, L, children);
}

__global__ void __launch_bounds__(256) k_area_up(AreaLevel lv, AreaLevel below, int has_below)
{
  NB2_GRID_STRIDE(j, lv.n + 1)
  {
    if (j == lv.n)
    {
      lv.leaves[j] = 0u;  // trailing zeros for the scans of level 0
      lv.inner[j] = 0u;
      continue;
    }
    areaUp(lv, has_below ? &below : nullptr, j);
  }
}

__global__ void __launch_bounds__(256) k_area_down(AreaLevel lv, AreaLevel below, unsigned* __restrict__ rank)
{
  NB2_GRID_STRIDE(j, lv.n) areaDown(lv, &below, j, rank);
}

__global__ void __launch_bounds__(256) k_area_emit(AreaLevel lv, uint32_t V, const unsigned* __restrict__ rank, uint32_t* __restrict__ out)
{
  NB2_GRID_STRIDE(j, lv.n) areaEmit(lv, j, V, rank, out);
}

/** centroid records from creation order to their final numbers, one 4-byte word per thread */
__global__ void __launch_bounds__(256)
    k_area_place(const uint32_t* __restrict__ fresh, unsigned long long n_new, uint32_t words, const unsigned* __restrict__ rank,
                 uint32_t* __restrict__ dst /* record V of the new cloud */)
{
  NB2_GRID_STRIDE(i, n_new * words)
  {
    const unsigned long long s = i / words, w = i - s * words;
    dst[static_cast<unsigned long long>(rank[s]) * words + w] = fresh[i];
  }
}

inline int gridFor(unsigned long long n, const nb2_context* c)
{
  const unsigned long long cap = static_cast<unsigned long long>(c->sm_count) * 16;
  unsigned long long b = (n + 255) / 256;
  return static_cast<int>(std::max<unsigned long long>(1, std::min(b, cap)));
}

#define NB2_LAUNCHED(c)                                                                                                \
  do                                                                                                                   \
  {                                                                                                                    \
    NB2_CUDA(cudaGetLastError());                                                                                      \
    ++(c)->launches;                                                                                                   \
  } while (0)

unsigned readBackU32(nb2_context* c, const unsigned* dev)
{
  c->hostSmall.ensure(4096);
  unsigned* h = c->hostSmall.as<unsigned>();
  NB2_CUDA(cudaMemcpyAsync(h, dev, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  NB2_CUDA(cudaStreamSynchronize(c->stream));
  return h[0];
}
}  // namespace

void launchMidpointSubdivision(nb2_context* c, const sub::Layout& L, int n_iter, int slot, uint64_t* V_out, uint64_t* F_out)
{
  const uint64_t V = c->V, F = c->F;
  const TreeShape shape = makeTreeShape(n_iter);
  const uint64_t nodes = F * shape.S[1];           // internal nodes of all trees
  const uint64_t F_out_n = F << (2 * n_iter);      // 4^n F
  // provisional ids are V + 3 * node + k: everything must stay below 2^32 - 1, and the result below the 2^31 limit of
  // the rest of the path (nb2_mesh_upload)
  if (3 * nodes + V >= 0xfffffff0ull || F_out_n >= (1ull << 31))
    fail(NB2_ERR_INVALID_ARGUMENT, "FaceMidpointSubdivision: " + std::to_string(n_iter) + " iterations of " + std::to_string(F) +
                                       " triangles exceed the 2^31 triangle / 2^32 request limit of this path");
  cudaStream_t st = c->stream;
  // levels 2 .. n in the workspace, level n + 1 (the output) in the triangle buffer that becomes the resident one
  uint64_t ws_tris = 0;
  for (int l = 2; l <= n_iter; ++l)
    ws_tris += F << (2 * (l - 1));
  c->subLevels.ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(ws_tris, 1));
  c->subTri[slot].ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(F_out_n, 1));
  const uint64_t top_nodes = F << (2 * (n_iter - 1));
  uint64_t slots = 1024;
  while (slots < 6 * top_nodes)
    slots <<= 1;
  c->subKeys.ensure(2 * sizeof(unsigned long long) * slots);  // 16-byte slots { key, smallest sequence number }
  c->subReq.ensure(sizeof(unsigned) * 3 * std::max<uint64_t>(top_nodes, 1));
  c->subNodeCnt.ensure(sizeof(unsigned) * (nodes + 1));
  c->subNodeBase.ensure(sizeof(unsigned) * (nodes + 1));
  c->subNodeMask.ensure(nodes + 1);
  NB2_CUDA(cudaMemsetAsync(c->subNodeCnt.as<unsigned>() + nodes, 0, sizeof(unsigned), st));

  std::vector<MidpointLevel> levels(n_iter + 1);
  {
    uint32_t* next_ws = c->subLevels.as<uint32_t>();
    const uint32_t* cur = c->tri.as<uint32_t>();
    for (int l = 1; l <= n_iter; ++l)
    {
      MidpointLevel& lv = levels[l];
      lv.shape = shape;
      lv.level = l;
      lv.n_nodes = F << (2 * (l - 1));
      lv.V = static_cast<uint32_t>(V);
      lv.tri = cur;
      uint32_t* children = l == n_iter ? c->subTri[slot].as<uint32_t>() : next_ws;
      if (lv.n_nodes)
      {
        uint64_t lslots = 1024;
        while (lslots < 6 * lv.n_nodes)
          lslots <<= 1;
        NB2_CUDA(cudaMemsetAsync(c->subKeys.ptr, 0xff, 2 * sizeof(unsigned long long) * lslots, st));
        const int grid = gridFor(lv.n_nodes, c);
        k_mid_insert<<<grid, 256, 0, st>>>(lv, c->subKeys.as<unsigned long long>(), static_cast<unsigned>(lslots - 1),
                                           c->subReq.as<unsigned>());
        NB2_LAUNCHED(c);
        k_mid_resolve<<<grid, 256, 0, st>>>(lv, c->subKeys.as<unsigned long long>(), c->subReq.as<unsigned>(), c->subNodeMask.as<uint8_t>(),
                                            c->subNodeCnt.as<unsigned>(), children);
        NB2_LAUNCHED(c);
      }
      cur = children;
      if (l != n_iter)
        next_ws += 3 * (lv.n_nodes << 2);
    }
  }
  launchScanU32(c, c->subNodeCnt.as<unsigned>(), c->subNodeBase.as<unsigned>(), nodes + 1);
  const uint64_t created = readBackU32(c, c->subNodeBase.as<unsigned>() + nodes);
  const uint64_t V_new = V + created;
  if (V_new >= (1ull << 31))
    fail(NB2_ERR_INVALID_ARGUMENT, "FaceMidpointSubdivision: the subdivided cloud would have 2^31 or more points");
  c->subBlob[slot].ensure(V_new * L.step);
  NB2_CUDA(cudaMemcpyAsync(c->subBlob[slot].ptr, c->blob.ptr, V * L.step, cudaMemcpyDeviceToDevice, st));
  for (int l = 1; l <= n_iter; ++l)
    if (levels[l].n_nodes)
    {
      k_mid_create<<<gridFor(levels[l].n_nodes, c), 256, 0, st>>>(levels[l], c->subNodeBase.as<unsigned>(),
                                                                  c->subNodeMask.as<uint8_t>(), c->subBlob[slot].as<uint8_t>(), L);
      NB2_LAUNCHED(c);
    }
  if (F_out_n)
  {
    k_mid_remap<<<gridFor(3 * F_out_n, c), 256, 0, st>>>(c->subTri[slot].as<uint32_t>(), 3 * F_out_n, static_cast<uint32_t>(V),
                                                         c->subNodeBase.as<unsigned>(), c->subNodeMask.as<uint8_t>());
    NB2_LAUNCHED(c);
  }
  *V_out = V_new;
  *F_out = F_out_n;
}

namespace
{
/** one level of the by-area sweep in one allocation: tri[3 n] split[n+1] pre[n+1] leaves[n+1] inner[n+1] off_leaf[n+1] off_inner[n+1] */
struct AreaLevelBuf
{
  DevBuf buf;
  uint64_t n = 0;
  uint32_t base = 0;
  uint32_t* tri() const { return buf.as<uint32_t>(); }
  unsigned* arr(int i) const { return buf.as<unsigned>() + 3 * n + static_cast<uint64_t>(i) * (n + 1); }
  AreaLevel view() const { return { n, base, tri(), arr(0), arr(1), arr(2), arr(3), arr(4), arr(5) }; }
  void alloc(uint64_t count)
  {
    n = count;
    buf.ensure(sizeof(uint32_t) * (3 * n + 6 * (n + 1)));
  }
};
}  // namespace

void launchAreaSubdivision(nb2_context* c, const sub::Layout& L, float max_area, uint64_t max_points, int slot, uint64_t* V_out,
                           uint64_t* F_out)
{
  const uint64_t V = c->V, F = c->F;
  cudaStream_t st = c->stream;
  const uint64_t limit = std::min<uint64_t>(max_points ? max_points : ~0ull, (1ull << 31) - 1);
  // The buffers of the levels come from a pool the context keeps (level i of this call reuses what level i of the last call
  // allocated), so a second subdivision of a mesh of the same size allocates nothing
  std::vector<AreaLevelBuf> levels;
  struct Cleanup
  {
    nb2_context* c;
    std::vector<AreaLevelBuf>& v;
    ~Cleanup()
    {
      for (size_t i = 0; i < v.size(); ++i)
      {
        if (c->areaPool.size() <= i)
          c->areaPool.resize(i + 1);
        std::swap(c->areaPool[i], v[i].buf);
      }
    }
  } cleanup{ c, levels };
  auto newLevel = [&]() -> AreaLevelBuf& {
    levels.emplace_back();
    const size_t i = levels.size() - 1;
    if (i < c->areaPool.size())
      std::swap(c->areaPool[i], levels[i].buf);
    return levels[i];
  };
  levels.reserve(64);
  newLevel();
  levels[0].alloc(F);
  if (F)
  {
    k_area_roots<<<gridFor(3 * F, c), 256, 0, st>>>(c->tri.as<uint32_t>(), F, levels[0].tri());
    NB2_LAUNCHED(c);
  }
  unsigned* non_finite = c->counters.as<unsigned>() + 20;
  NB2_CUDA(cudaMemsetAsync(non_finite, 0, sizeof(unsigned), st));
  uint64_t created = 0;
  c->subFresh.ensure(std::max<uint64_t>(L.step, 4));
  for (;;)
  {
    AreaLevelBuf& lv = levels.back();
    k_area_test<<<gridFor(lv.n + 1, c), 256, 0, st>>>(lv.tri(), lv.n, static_cast<uint32_t>(V), c->blob.as<uint8_t>(),
                                                      c->subFresh.as<uint8_t>(), L, max_area, lv.arr(0), non_finite);
    NB2_LAUNCHED(c);
    launchScanU32(c, lv.arr(0), lv.arr(1), lv.n + 1);
    const uint64_t splits = readBackU32(c, lv.arr(1) + lv.n);
    if (readBackU32(c, non_finite))
      fail(NB2_ERR_NON_FINITE, "Triangle area is not finite");  // face_subdivision_modifier.cpp:332-333
    if (!splits)
      break;
    if (V + created + splits > limit || levels.size() >= 4096)
      fail(NB2_ERR_INVALID_ARGUMENT, "FaceSubdivisionByArea: max_area " + std::to_string(max_area) + " needs more than " +
                                         std::to_string(limit) + " points (the reference would not terminate or run out of memory)");
    // the records made so far must survive the growth of their buffer
    const size_t need = (created + splits) * L.step;
    if (need > c->subFresh.cap)
    {
      DevBuf bigger;
      bigger.ensure(2 * need);
      if (created)
        NB2_CUDA(cudaMemcpyAsync(bigger.ptr, c->subFresh.ptr, created * L.step, cudaMemcpyDeviceToDevice, st));
      NB2_CUDA(cudaStreamSynchronize(st));
      std::swap(c->subFresh, bigger);
      bigger.release();
    }
    newLevel();
    AreaLevelBuf& parent = levels[levels.size() - 2];
    AreaLevelBuf& child = levels.back();
    child.alloc(3 * splits);
    child.base = static_cast<uint32_t>(created + splits);
    k_area_split<<<gridFor(parent.n, c), 256, 0, st>>>(parent.tri(), parent.n, parent.arr(0), parent.arr(1), parent.base,
                                                       static_cast<uint32_t>(V), c->blob.as<uint8_t>(), c->subFresh.as<uint8_t>(), L,
                                                       child.tri());
    NB2_LAUNCHED(c);
    created += splits;
  }
  // bottom-up sizes, top-down offsets
  for (size_t l = levels.size(); l-- > 0;)
  {
    const bool has_below = l + 1 < levels.size();
    k_area_up<<<gridFor(levels[l].n + 1, c), 256, 0, st>>>(levels[l].view(), has_below ? levels[l + 1].view() : levels[l].view(),
                                                           has_below ? 1 : 0);
    NB2_LAUNCHED(c);
  }
  launchScanU32(c, levels[0].arr(2), levels[0].arr(4), levels[0].n + 1);
  launchScanU32(c, levels[0].arr(3), levels[0].arr(5), levels[0].n + 1);
  const uint64_t F_new = readBackU32(c, levels[0].arr(4) + levels[0].n);
  if (F_new >= (1ull << 31))
    fail(NB2_ERR_INVALID_ARGUMENT, "FaceSubdivisionByArea: the subdivided mesh would have 2^31 or more triangles");
  c->subRank.ensure(sizeof(unsigned) * std::max<uint64_t>(created, 1));
  for (size_t l = 0; l + 1 < levels.size(); ++l)
    if (levels[l].n)
    {
      k_area_down<<<gridFor(levels[l].n, c), 256, 0, st>>>(levels[l].view(), levels[l + 1].view(), c->subRank.as<unsigned>());
      NB2_LAUNCHED(c);
    }
  c->subTri[slot].ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(F_new, 1));
  for (size_t l = 0; l < levels.size(); ++l)
    if (levels[l].n)
    {
      k_area_emit<<<gridFor(levels[l].n, c), 256, 0, st>>>(levels[l].view(), static_cast<uint32_t>(V), c->subRank.as<unsigned>(),
                                                           c->subTri[slot].as<uint32_t>());
      NB2_LAUNCHED(c);
    }
  const uint64_t V_new = V + created;
  c->subBlob[slot].ensure(V_new * L.step);
  NB2_CUDA(cudaMemcpyAsync(c->subBlob[slot].ptr, c->blob.ptr, V * L.step, cudaMemcpyDeviceToDevice, st));
  if (created)
  {
    k_area_place<<<gridFor(created * (L.step / 4), c), 256, 0, st>>>(c->subFresh.as<uint32_t>(), created, L.step / 4,
                                                                     c->subRank.as<unsigned>(),
                                                                     c->subBlob[slot].as<uint32_t>() + V * (L.step / 4));
    NB2_LAUNCHED(c);
  }
  // the level buffers are freed by `cleanup` below: everything queued on the stream must be done with them first
  NB2_CUDA(cudaStreamSynchronize(st));
  *V_out = V_new;
  *F_out = F_new;
}
}  // namespace nb2
