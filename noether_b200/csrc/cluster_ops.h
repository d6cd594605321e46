// Euclidean clustering of the resident cloud: the per-point work, shared by the CUDA kernels
// (kernels_cluster.cu) and by the host emulation of the launch sequence the CPU tests run
// (tests/cpp/cluster_emulation.cpp).
//
// EuclideanClusteringMeshModifier::modify (euclidean_clustering_modifier.cpp:40-64) grows clusters breadth first through
// k-d tree radius searches (pcl::extractEuclideanClusters).  What it returns depends only on WHICH points end up together —
// the connected components of the graph "FLANN's float squared distance < float(tolerance^2)" — because the sub-meshes are
// cut by membership (subset_extractor.cpp:22-56).  So here: a uniform grid of cells slightly wider than the tolerance
// (hash table of occupied cells, points chained per cell), every point tests the points of its 27 surrounding cells, and a
// lock-free union-find that always hangs the larger root under the smaller one — the root of a component is its smallest
// point index, which is also the seed pcl::extractEuclideanClusters starts that cluster from, so labels sorted ascending
// are the clusters in PCL's order of discovery.
#pragma once
#include "host_math.h"

#include <cstdint>

namespace nb2
{
namespace clu
{
constexpr unsigned long long kEmptyCell = ~0ull;
constexpr unsigned kNone = 0xffffffffu;
// Cells per axis.  cellCoord divides in double (relative error about 2^-52) while the safety margin of the cell size is
// 2^-10 of a cell: beyond about 2^41 cells per axis two points within the tolerance could land two cells apart and never
// be compared.  2^40 keeps a factor of two in hand; wider grids are refused.
constexpr double kMaxCells = 1099511627776.0;

struct GridParams
{
  double mn[3];  // smallest coordinate of the cloud per axis
  double h;      // cell size: tolerance * (1 + 2^-10), so that rounding can never put a neighbour two cells away
  float r2;      // float(tolerance * tolerance): what pcl::KdTreeFLANN hands to FLANN
};

NB2_HD inline long long cellCoord(const GridParams& g, int axis, float x)
{
  return static_cast<long long>(floor((static_cast<double>(x) - g.mn[axis]) / g.h));
}
/** 64-bit key of a cell.  Two cells may share a key: their chains then merge, which only adds candidates to the distance
 *  test — membership never depends on it — so the grid has no limit on its extent. */
NB2_HD inline unsigned long long cellKey(long long cx, long long cy, long long cz)
{
  unsigned long long k = static_cast<unsigned long long>(cx) * 0x9e3779b97f4a7c15ull;
  k ^= (static_cast<unsigned long long>(cy) * 0xc2b2ae3d27d4eb4full) << 21 | (static_cast<unsigned long long>(cy) * 0xc2b2ae3d27d4eb4full) >> 43;
  k ^= (static_cast<unsigned long long>(cz) * 0x165667b19e3779f9ull) << 42 | (static_cast<unsigned long long>(cz) * 0x165667b19e3779f9ull) >> 22;
  k ^= k >> 29;
  k *= 0xbf58476d1ce4e5b9ull;
  k ^= k >> 32;
  return k == kEmptyCell ? k - 1 : k;
}
NB2_HD inline unsigned hashCell(unsigned long long k)
{
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return static_cast<unsigned>(k);
}

/** FLANN L2_Simple<float>: result = 0; result += diff * diff per axis; a neighbour when strictly below radius^2 */
NB2_HD inline bool isNeighbour(float ax, float ay, float az, float bx, float by, float bz, float r2)
{
  float result = 0.f;
  float d = ax - bx;
  result += d * d;
  d = ay - by;
  result += d * d;
  d = az - bz;
  result += d * d;
  return result < r2;
}

/** C1: point p joins the chain of its cell (the cell's slot is claimed on first use: returns true for the claimant) */
template <typename Cas, typename Xchg>
NB2_HD inline bool insertPoint(const GridParams& g, unsigned p, float x, float y, float z, unsigned long long* cell_key,
                               unsigned* cell_head, unsigned slot_mask, unsigned* next, unsigned* parent, Cas cas, Xchg xchg)
{
  const unsigned long long key = cellKey(cellCoord(g, 0, x), cellCoord(g, 1, y), cellCoord(g, 2, z));
  unsigned s = hashCell(key) & slot_mask;
  bool claimed = false;
  for (;;)
  {
    const unsigned long long seen = cas(cell_key + s, kEmptyCell, key);
    claimed = seen == kEmptyCell;
    if (claimed || seen == key)
      break;
    s = (s + 1) & slot_mask;
  }
  next[p] = xchg(cell_head + s, p);
  parent[p] = p;
  return claimed;
}

/** root of x, halving the path on the way (a stale read only costs another step: parents only ever move towards roots) */
NB2_HD inline unsigned findRoot(unsigned* parent, unsigned x)
{
  volatile unsigned* vp = parent;
  for (;;)
  {
    const unsigned p = vp[x];
    if (p == x)
      return x;
    const unsigned gp = vp[p];
    if (gp != p)
      vp[x] = gp;
    x = p;
  }
}

/** the larger root goes under the smaller one */
template <typename Cas32>
NB2_HD inline void unite(unsigned* parent, unsigned a, unsigned b, Cas32 cas)
{
  for (;;)
  {
    a = findRoot(parent, a);
    b = findRoot(parent, b);
    if (a == b)
      return;
    if (a < b)
    {
      const unsigned t = a;
      a = b;
      b = t;
    }
    if (cas(parent + a, a, b) == a)
      return;
  }
}

/** slot of the cell that holds (x, y, z), or kNone when no point was inserted there */
NB2_HD inline unsigned findCellSlot(const GridParams& g, long long x, long long y, long long z, const unsigned long long* cell_key,
                                    unsigned slot_mask)
{
  if (x < 0 || y < 0 || z < 0)
    return kNone;
  const unsigned long long key = cellKey(x, y, z);
  unsigned s = hashCell(key) & slot_mask;
  unsigned long long seen;
  while ((seen = cell_key[s]) != key && seen != kEmptyCell)
    s = (s + 1) & slot_mask;
  return seen == kEmptyCell ? kNone : s;
}

/** C2: point p against the points with a smaller index in its 27 surrounding cells */
template <typename Cas32, typename Pos>
NB2_HD inline void linkPoint(const GridParams& g, unsigned p, Pos pos, const unsigned long long* cell_key, const unsigned* cell_head,
                             unsigned slot_mask, const unsigned* next, unsigned* parent, Cas32 cas)
{
  float px, py, pz;
  pos(p, px, py, pz);
  const long long c[3] = { cellCoord(g, 0, px), cellCoord(g, 1, py), cellCoord(g, 2, pz) };
  for (long long dz = -1; dz <= 1; ++dz)
    for (long long dy = -1; dy <= 1; ++dy)
      for (long long dx = -1; dx <= 1; ++dx)
      {
        const long long x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
        if (x < 0 || y < 0 || z < 0)
          continue;
        const unsigned long long key = cellKey(x, y, z);
        unsigned s = hashCell(key) & slot_mask;
        unsigned long long seen;
        while ((seen = cell_key[s]) != key && seen != kEmptyCell)
          s = (s + 1) & slot_mask;
        if (seen == kEmptyCell)
          continue;
        for (unsigned q = cell_head[s]; q != kNone; q = next[q])
        {
          if (q >= p)
            continue;
          float qx, qy, qz;
          pos(q, qx, qy, qz);
          if (isNeighbour(px, py, pz, qx, qy, qz, g.r2))
            unite(parent, p, q, cas);
        }
      }
}
}  // namespace clu
}  // namespace nb2
