// TEST INFRASTRUCTURE.  Host emulation of the launch sequence of noether_b200/csrc/kernels_cluster.cu: the same per-point
// functions (cluster_ops.h) driven by loops in a scrambled visiting order instead of kernels, so that the labels can be
// compared with the oracle's breadth-first clusters on a machine without a GPU.
#include "../../noether_b200/csrc/cluster_ops.h"

#include <algorithm>
#include <numeric>
#include <vector>

using namespace nb2::clu;

namespace
{
struct HostCas64
{
  unsigned long long operator()(unsigned long long* addr, unsigned long long expect, unsigned long long val) const
  {
    const unsigned long long seen = *addr;
    if (seen == expect)
      *addr = val;
    return seen;
  }
};
struct HostCas32
{
  unsigned operator()(unsigned* addr, unsigned expect, unsigned val) const
  {
    const unsigned seen = *addr;
    if (seen == expect)
      *addr = val;
    return seen;
  }
};
struct HostXchg
{
  unsigned operator()(unsigned* addr, unsigned val) const
  {
    const unsigned old = *addr;
    *addr = val;
    return old;
  }
};
struct HostPos
{
  const float* xyz;
  void operator()(unsigned i, float& x, float& y, float& z) const
  {
    x = xyz[3 * i];
    y = xyz[3 * i + 1];
    z = xyz[3 * i + 2];
  }
};
template <typename F>
void scrambled(unsigned long long n, F f)
{
  if (n == 0)
    return;
  unsigned long long stride = 7919;
  while (std::gcd(stride, n) != 1)
    ++stride;
  unsigned long long i = (n * 3) / 5 % n;
  for (unsigned long long c = 0; c < n; ++c, i = (i + stride) % n)
    f(static_cast<unsigned>(i));
}
}  // namespace

extern "C" int emu_cluster_labels(const float* xyz, uint64_t V, double tolerance, uint32_t* labels)
{
  if (!(tolerance > 0.0))
  {
    for (uint64_t v = 0; v < V; ++v)
      labels[v] = static_cast<uint32_t>(v);
    return 0;
  }
  GridParams g{};
  g.h = tolerance * (1.0 + 1.0 / 1024.0);
  g.r2 = static_cast<float>(tolerance * tolerance);
  for (int i = 0; i < 3; ++i)
  {
    float mn = xyz[i], mx = xyz[i];
    for (uint64_t v = 1; v < V; ++v)
      mn = std::min(mn, xyz[3 * v + i]), mx = std::max(mx, xyz[3 * v + i]);
    g.mn[i] = static_cast<double>(mn) - g.h;
    if (!((static_cast<double>(mx) - g.mn[i]) / g.h < kMaxCells))
      return 1;
  }
  uint64_t slots = 1024;
  while (slots < 2 * V)
    slots <<= 1;
  std::vector<unsigned long long> cell_key(slots, kEmptyCell);
  std::vector<unsigned> cell_head(slots, kNone), next(V), parent(V);
  const unsigned mask = static_cast<unsigned>(slots - 1);
  scrambled(V, [&](unsigned p) {
    insertPoint(g, p, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], cell_key.data(), cell_head.data(), mask, next.data(), parent.data(),
                HostCas64{}, HostXchg{});
  });
  scrambled(V, [&](unsigned p) {
    linkPoint(g, p, HostPos{ xyz }, cell_key.data(), cell_head.data(), mask, next.data(), parent.data(), HostCas32{});
  });
  scrambled(V, [&](unsigned p) { labels[p] = findRoot(parent.data(), p); });
  return 0;
}
