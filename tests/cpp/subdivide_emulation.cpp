// TEST INFRASTRUCTURE.  Host emulation of the launch sequence of noether_b200/csrc/kernels_subdivide.cu: the same
// per-element functions (subdivide_ops.h) driven by loops instead of kernels, so that the numbering scheme can be
// compared with the oracle bit for bit on a machine without a GPU.  "Threads" run in a scrambled order (a stride
// permutation) to show that no pass depends on the order its elements are visited in.
#include "../../noether_b200/csrc/subdivide_ops.h"

#include <cstring>
#include <numeric>
#include <string>
#include <vector>

using namespace nb2::sub;

namespace
{
std::vector<uint8_t> g_cloud;
std::vector<uint32_t> g_tri;
std::string g_error;

struct HostCas
{
  unsigned long long operator()(unsigned long long* addr, unsigned long long expect, unsigned long long val) const
  {
    const unsigned long long seen = *addr;
    if (seen == expect)
      *addr = val;
    return seen;
  }
};
struct HostMin
{
  void operator()(unsigned* addr, unsigned val) const
  {
    if (val < *addr)
      *addr = val;
  }
};

/** visit 0 .. n-1 in a scrambled order */
template <typename F>
void scrambled(unsigned long long n, F f)
{
  if (n == 0)
    return;
  unsigned long long stride = 7919;
  while (std::gcd(stride, n) != 1)
    ++stride;
  unsigned long long i = (n * 5) / 7 % n;
  for (unsigned long long c = 0; c < n; ++c, i = (i + stride) % n)
    f(i);
}

void exclusiveScan(const std::vector<unsigned>& in, std::vector<unsigned>& out)
{
  out.resize(in.size());
  unsigned run = 0;
  for (size_t i = 0; i < in.size(); ++i)
  {
    out[i] = run;
    run += in[i];
  }
}
}  // namespace

extern "C" {

const char* emu_sub_last_error() { return g_error.c_str(); }
uint64_t emu_sub_num_points(uint32_t step) { return g_cloud.size() / step; }
uint64_t emu_sub_num_triangles() { return g_tri.size() / 3; }
void emu_sub_export(uint8_t* cloud, uint32_t* tri)
{
  std::memcpy(cloud, g_cloud.data(), g_cloud.size());
  std::memcpy(tri, g_tri.data(), g_tri.size() * 4);
}

int emu_midpoint(const uint8_t* cloud, uint64_t V, uint32_t step, uint32_t off_xyz, int32_t off_nrm, int32_t off_rgba,
                 const uint32_t* tri, uint64_t F, int n_iter)
{
  const Layout L{ step, off_xyz, off_nrm, off_rgba };
  const TreeShape shape = makeTreeShape(n_iter);
  const uint64_t nodes = F * shape.S[1];
  std::vector<std::vector<uint32_t>> level_tri(n_iter + 2);
  level_tri[1].assign(tri, tri + 3 * F);
  std::vector<uint8_t> node_mask(nodes + 1, 0);
  std::vector<unsigned> node_cnt(nodes + 1, 0), node_base;
  std::vector<MidpointLevel> levels(n_iter + 1);
  for (int l = 1; l <= n_iter; ++l)
  {
    MidpointLevel& lv = levels[l];
    lv.shape = shape;
    lv.level = l;
    lv.n_nodes = F << (2 * (l - 1));
    lv.V = static_cast<uint32_t>(V);
    lv.tri = level_tri[l].data();
    level_tri[l + 1].assign(12 * lv.n_nodes, 0);
    uint64_t slots = 1024;
    while (slots < 6 * lv.n_nodes)
      slots <<= 1;
    std::vector<unsigned long long> table(2 * slots, kEmptyKey);  // { key, smallest sequence number } pairs, all bits set
    std::vector<unsigned> req(3 * lv.n_nodes + 1);
    scrambled(lv.n_nodes, [&](unsigned long long j) {
      midpointInsert(lv, j, table.data(), static_cast<unsigned>(slots - 1), req.data(), HostCas{}, HostMin{});
    });
    scrambled(lv.n_nodes, [&](unsigned long long j) {
      midpointResolve(lv, j, table.data(), req.data(), node_mask.data(), node_cnt.data(), level_tri[l + 1].data());
    });
  }
  exclusiveScan(node_cnt, node_base);
  const uint64_t V_new = V + node_base[nodes];
  g_cloud.assign(V_new * step, 0xcd);
  std::memcpy(g_cloud.data(), cloud, V * step);
  for (int l = 1; l <= n_iter; ++l)
    scrambled(levels[l].n_nodes,
              [&](unsigned long long j) { midpointCreate(levels[l], j, node_base.data(), node_mask.data(), g_cloud.data(), L); });
  g_tri = level_tri[n_iter + 1];
  scrambled(g_tri.size(), [&](unsigned long long i) {
    g_tri[i] = midpointFinal(g_tri[i], static_cast<uint32_t>(V), node_base.data(), node_mask.data());
  });
  return 0;
}

int emu_by_area(const uint8_t* cloud, uint64_t V, uint32_t step, uint32_t off_xyz, int32_t off_nrm, int32_t off_rgba,
                const uint32_t* tri, uint64_t F, float max_area, uint64_t max_points)
{
  const Layout L{ step, off_xyz, off_nrm, off_rgba };
  struct Level
  {
    uint64_t n = 0;
    uint32_t base = 0;
    std::vector<uint32_t> tri;
    std::vector<unsigned> split, pre, leaves, inner, off_leaf, off_inner;
    AreaLevel view() { return { n, base, tri.data(), split.data(), pre.data(), leaves.data(), inner.data(), off_leaf.data(), off_inner.data() }; }
    void alloc(uint64_t count)
    {
      n = count;
      tri.assign(3 * n + 1, 0);
      for (auto* v : { &split, &pre, &leaves, &inner, &off_leaf, &off_inner })
        v->assign(n + 1, 0);
    }
  };
  std::vector<Level> levels(1);
  levels[0].alloc(F);
  for (uint64_t i = 0; i < 3 * F; ++i)
    levels[0].tri[i] = tri[3 * (F - 1 - i / 3) + i % 3];
  std::vector<uint8_t> fresh;
  uint64_t created = 0;
  const uint64_t limit = max_points ? max_points : (1ull << 31) - 1;
  for (;;)
  {
    Level& lv = levels.back();
    bool non_finite = false;
    scrambled(lv.n, [&](unsigned long long j) {
      const unsigned r = areaTest(lv.tri.data(), j, static_cast<uint32_t>(V), cloud, fresh.data(), L, max_area);
      non_finite |= r == 2u;
      lv.split[j] = r == 1u;
    });
    lv.split[lv.n] = 0;
    exclusiveScan(lv.split, lv.pre);
    if (non_finite)
    {
      g_error = "Triangle area is not finite";
      return 2;
    }
    const uint64_t splits = lv.pre[lv.n];
    if (!splits)
      break;
    if (V + created + splits > limit)
    {
      g_error = "too many points";
      return 1;
    }
    fresh.resize((created + splits) * step, 0xcd);
    levels.emplace_back();
    Level& parent = levels[levels.size() - 2];
    Level& child = levels.back();
    child.alloc(3 * splits);
    child.base = static_cast<uint32_t>(created + splits);
    scrambled(parent.n, [&](unsigned long long j) {
      if (parent.split[j])
        areaSplit(parent.tri.data(), j, parent.pre[j], parent.base, static_cast<uint32_t>(V), cloud, fresh.data(), L, child.tri.data());
    });
    created += splits;
  }
  for (size_t l = levels.size(); l-- > 0;)
  {
    AreaLevel me = levels[l].view();
    AreaLevel below = l + 1 < levels.size() ? levels[l + 1].view() : me;
    scrambled(levels[l].n, [&](unsigned long long j) { areaUp(me, l + 1 < levels.size() ? &below : nullptr, j); });
    levels[l].leaves[levels[l].n] = 0;
    levels[l].inner[levels[l].n] = 0;
  }
  exclusiveScan(levels[0].leaves, levels[0].off_leaf);
  exclusiveScan(levels[0].inner, levels[0].off_inner);
  const uint64_t F_new = levels[0].off_leaf[levels[0].n];
  std::vector<unsigned> rank(created + 1, 0xffffffffu);
  for (size_t l = 0; l + 1 < levels.size(); ++l)
  {
    AreaLevel me = levels[l].view(), below = levels[l + 1].view();
    scrambled(levels[l].n, [&](unsigned long long j) { areaDown(me, &below, j, rank.data()); });
  }
  g_tri.assign(3 * F_new, 0xffffffffu);
  for (size_t l = 0; l < levels.size(); ++l)
  {
    AreaLevel me = levels[l].view();
    scrambled(levels[l].n, [&](unsigned long long j) { areaEmit(me, j, static_cast<uint32_t>(V), rank.data(), g_tri.data()); });
  }
  g_cloud.assign((V + created) * step, 0xcd);
  std::memcpy(g_cloud.data(), cloud, V * step);
  for (uint64_t s = 0; s < created; ++s)
    std::memcpy(g_cloud.data() + (V + rank[s]) * step, fresh.data() + s * step, step);
  return 0;
}
}
