// Face subdivision mesh modifiers on the device-resident cloud: the per-element work, shared by the
// CUDA kernels (kernels_subdivide.cu) and by the host emulation of the launch sequence the CPU tests run
// (tests/cpp/subdivide_emulation.cpp).
//
//   FaceMidpointSubdivisionMeshModifier::modify   face_subdivision_modifier.cpp:202-248
//   FaceSubdivisionByAreaMeshModifier::modify     face_subdivision_modifier.cpp:250-339
//   copyPointToEnd / updatePointData              :58-114
//
// The reference walks the faces depth first on one core and numbers new vertices in creation order.  Here the trees are
// expanded level by level and the depth-first numbers are recovered from closed forms / prefix sums:
//
//  Midpoint: every root triangle carries a complete 4-ary tree of depth n.  An internal node asks for its three edge
//  midpoints when the walk enters it, so request k of the node with pre-order number p in root t has the sequence number
//  3 (t S_1 + p) + k, S_i = (4^(n-i+1) - 1) / 3 nodes in a subtree rooted at level i.  An edge's midpoint is created by
//  the request with the smallest sequence number (hash insert + atomicMin per level: an edge of level l joins a vertex
//  made at level l - 1, so levels never share edges); until the final numbers exist a new vertex is known as
//  V + that sequence number.  Counting the creating requests per node in pre-order and scanning gives the final numbers.
//
//  By area: the stack pops the LAST face first and the three children of a split triangle in the reverse of their push
//  order.  Level by level, triangles whose area exceeds the limit get a centroid record (provisional number: order of
//  creation in this sweep) and three children; subtree sizes folded bottom-up and offsets pushed top-down give every
//  leaf its place in the output and every centroid its final number.
//
// Arithmetic as the reference's Eigen 3.4 expressions (x86-64, no FMA: the library is built with -fmad=false /
// -ffp-contract=off): mean = ((0 + p0) + p1 [+ p2]) / float(count); rgba by int division;
// area = 0.5f * sqrt(cx^2 + (cy^2 + cz^2)).
#pragma once
#include "host_math.h"

#include <cstdint>
#include <cstring>

namespace nb2
{
namespace sub
{
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kMaxIterations = 15;

/** record layout of the cloud (pcl::PCLPointCloud2: point_step and the offsets of x, normal_x, rgba; -1 = absent) */
struct Layout
{
  uint32_t step;
  uint32_t off_xyz;
  int32_t off_nrm;
  int32_t off_rgba;
};

/** subtree sizes of the midpoint recursion: S[i] = internal nodes below (and including) a node of level i, 1-based */
struct TreeShape
{
  int n;                                   // n_iterations
  unsigned long long S[kMaxIterations + 2];  // S[n + 1] = 0
};
inline TreeShape makeTreeShape(int n)
{
  TreeShape s{};
  s.n = n;
  s.S[n + 1] = 0;
  for (int i = n; i >= 1; --i)
    s.S[i] = 1 + 4 * s.S[i + 1];
  return s;
}

/** pre-order number, among all internal nodes, of node j of level `level` (level-local index: root * 4^(level-1) + path) */
NB2_HD inline unsigned long long midpointNode(const TreeShape& s, int level, unsigned long long j)
{
  const int shift = 2 * (level - 1);
  const unsigned long long root = j >> shift;
  unsigned long long pre = 0;
  for (int i = 1; i < level; ++i)
  {
    const unsigned d = static_cast<unsigned>(j >> (2 * (level - 1 - i))) & 3u;
    pre += 1 + d * s.S[i + 1];
  }
  return root * s.S[1] + pre;
}

NB2_HD inline unsigned long long edgeKey(uint32_t a, uint32_t b)
{
  return a < b ? (static_cast<unsigned long long>(a) << 32) | b : (static_cast<unsigned long long>(b) << 32) | a;
}
NB2_HD inline unsigned hashKey(unsigned long long k)
{
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return static_cast<unsigned>(k);
}
NB2_HD inline unsigned popc3(unsigned m) { return (m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u); }

/** final vertex number of a provisional midpoint id (V + sequence number of the creating request) */
NB2_HD inline uint32_t midpointFinal(uint32_t id, uint32_t V, const unsigned* node_base, const uint8_t* node_mask)
{
  if (id < V)
    return id;
  const uint32_t seq = id - V;
  const uint32_t node = seq / 3u, k = seq - 3u * node;
  return V + node_base[node] + popc3(node_mask[node] & ((1u << k) - 1u));
}

// ---- records -----------------------------------------------------------------------------------------------------
struct F3
{
  float x, y, z;
};
NB2_HD inline F3 loadF3(const uint8_t* rec, uint32_t off)
{
  const float* f = reinterpret_cast<const float*>(rec + off);
  return { f[0], f[1], f[2] };
}
NB2_HD inline void storeF3(uint8_t* rec, uint32_t off, const F3& v)
{
  float* f = reinterpret_cast<float*>(rec + off);
  f[0] = v.x;
  f[1] = v.y;
  f[2] = v.z;
}

/** 16 bytes moved at once (records whose step is a multiple of 16 are copied with a quarter of the memory instructions) */
struct alignas(16) W4
{
  uint32_t w[4];
};
NB2_HD inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
NB2_HD inline uint32_t floatBits(float f)
{
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

/** ((0 + p0) + p1 [+ p2]) / float(n) of a float triple at `off` */
NB2_HD inline F3 meanF3(const uint8_t* const* src, int n, uint32_t off)
{
  const float count = static_cast<float>(n);
  F3 mean{ 0.f, 0.f, 0.f };
  for (int k = 0; k < n; ++k)
  {
    const F3 p = loadF3(src[k], off);
    mean.x = mean.x + p.x;
    mean.y = mean.y + p.y;
    mean.z = mean.z + p.z;
  }
  return { mean.x / count, mean.y / count, mean.z / count };
}

/** copyPointToEnd + updatePointData: dst = byte copy of src[0] whose x y z / normal / rgba are the mean of the n sources.
 *  The means are formed first and patched into the words of the copy as they pass, so that every byte of the new record
 *  is written once. */
NB2_HD inline void makeMeanRecord(uint8_t* dst, const uint8_t* const* src, int n, const Layout& L)
{
  const F3 xyz = meanF3(src, n, L.off_xyz);
  const bool has_nrm = L.off_nrm >= 0;
  F3 nrm{ 0.f, 0.f, 0.f };
  if (has_nrm)
    nrm = meanF3(src, n, static_cast<uint32_t>(L.off_nrm));
  uint8_t rgba[4] = { 0, 0, 0, 0 };
  if (L.off_rgba >= 0)
    for (int ch = 0; ch < 4; ++ch)
    {
      int mean = 0;
      for (int k = 0; k < n; ++k)
        mean += src[k][L.off_rgba + ch];
      rgba[ch] = static_cast<uint8_t>(mean / n);
    }
  const bool rgba_word = L.off_rgba >= 0 && (L.off_rgba & 3) == 0;
  const uint32_t rgba_bits = rgba[0] | (static_cast<uint32_t>(rgba[1]) << 8) | (static_cast<uint32_t>(rgba[2]) << 16) |
                             (static_cast<uint32_t>(rgba[3]) << 24);
  const uint32_t wx = L.off_xyz / 4, wn = has_nrm ? static_cast<uint32_t>(L.off_nrm) / 4 : 0u,
                 wc = rgba_word ? static_cast<uint32_t>(L.off_rgba) / 4 : 0u;
  auto patched = [&](uint32_t i, uint32_t v) {
    if (i >= wx && i < wx + 3)
      v = floatBits(i == wx ? xyz.x : (i == wx + 1 ? xyz.y : xyz.z));
    if (has_nrm && i >= wn && i < wn + 3)
      v = floatBits(i == wn ? nrm.x : (i == wn + 1 ? nrm.y : nrm.z));
    if (rgba_word && i == wc)
      v = rgba_bits;
    return v;
  };
  if ((L.step & 15u) == 0 && aligned16(src[0]) && aligned16(dst))
  {
    const W4* s = reinterpret_cast<const W4*>(src[0]);
    W4* d = reinterpret_cast<W4*>(dst);
    for (uint32_t q = 0; q < L.step / 16; ++q)
    {
      W4 v = s[q];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (uint32_t c = 0; c < 4; ++c)
        v.w[c] = patched(4 * q + c, v.w[c]);
      d[q] = v;
    }
  }
  else
  {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src[0]);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (uint32_t i = 0; i < L.step / 4; ++i)
      d[i] = patched(i, s[i]);
  }
  if (L.off_rgba >= 0 && !rgba_word)
    for (int ch = 0; ch < 4; ++ch)
      dst[L.off_rgba + ch] = rgba[ch];
}

/** 0.5f * ((v1 - v0).cross(v2 - v0)).norm()   (face_subdivision_modifier.cpp:330) */
NB2_HD inline float triangleArea(const F3& v0, const F3& v1, const F3& v2)
{
  const float ax = v1.x - v0.x, ay = v1.y - v0.y, az = v1.z - v0.z;
  const float bx = v2.x - v0.x, by = v2.y - v0.y, bz = v2.z - v0.z;
  const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
  return 0.5f * sqrtf(cx * cx + (cy * cy + cz * cz));
}

// ---- midpoint subdivision, one internal node (= one triangle of the level) per call ----------------------------------
struct MidpointLevel
{
  TreeShape shape;
  int level;                   // 1-based
  unsigned long long n_nodes;  // triangles of this level
  uint32_t V;                  // points of the input cloud
  const uint32_t* tri;         // uint32[3 n_nodes] (provisional ids)
};

/** A slot of the level's hash table is 16 bytes — the edge key and, next to it, the smallest sequence number seen — so
 *  that the insert's compare-and-swap and its atomicMin touch one 32-byte sector of HBM instead of two. */
NB2_HD inline unsigned* slotValue(unsigned long long* table, unsigned long long s) { return reinterpret_cast<unsigned*>(table + 2 * s + 1); }
NB2_HD inline const unsigned* slotValue(const unsigned long long* table, unsigned long long s)
{
  return reinterpret_cast<const unsigned*>(table + 2 * s + 1);
}

/** pass A1: the three edges of the node go into the level's hash table; every slot keeps its smallest sequence number.
 *  `cas` / `amin` are the atomics of the caller's world (device atomics, or plain host code). */
template <typename Cas, typename Min>
NB2_HD inline void midpointInsert(const MidpointLevel& lv, unsigned long long j, unsigned long long* table, unsigned slot_mask,
                                  unsigned* req_slot, Cas cas, Min amin)
{
  const unsigned long long node = midpointNode(lv.shape, lv.level, j);
  const uint32_t v[3] = { lv.tri[3 * j], lv.tri[3 * j + 1], lv.tri[3 * j + 2] };
  for (int k = 0; k < 3; ++k)
  {
    const unsigned long long key = edgeKey(v[k], v[k == 2 ? 0 : k + 1]);
    const unsigned seq = static_cast<unsigned>(3 * node + k);
    unsigned s = hashKey(key) & slot_mask;
    for (;;)
    {
      const unsigned long long seen = cas(table + 2ull * s, kEmptyKey, key);
      if (seen == kEmptyKey || seen == key)
        break;
      s = (s + 1) & slot_mask;
    }
    amin(slotValue(table, s), seq);
    req_slot[3 * j + k] = s;
  }
}

/** pass A2: midpoint ids of the node (provisional), which of its requests create a vertex, its four children */
NB2_HD inline void midpointResolve(const MidpointLevel& lv, unsigned long long j, const unsigned long long* table,
                                   const unsigned* req_slot, uint8_t* node_mask, unsigned* node_cnt, uint32_t* children)
{
  const unsigned long long node = midpointNode(lv.shape, lv.level, j);
  const uint32_t t[3] = { lv.tri[3 * j], lv.tri[3 * j + 1], lv.tri[3 * j + 2] };
  uint32_t m[3];
  unsigned mask = 0;
  for (int k = 0; k < 3; ++k)
  {
    const unsigned first = *slotValue(table, req_slot[3 * j + k]);
    m[k] = lv.V + first;
    if (first == static_cast<unsigned>(3 * node + k))
      mask |= 1u << k;
  }
  node_mask[node] = static_cast<uint8_t>(mask);
  node_cnt[node] = popc3(mask);
  // { t0 m0 m2 } { m0 t1 m1 } { m2 m1 t2 } { m0 m1 m2 }   (face_subdivision_modifier.cpp:170-173)
  uint32_t* c = children + 12 * j;
  if (aligned16(c))
  {
    W4* c4 = reinterpret_cast<W4*>(c);
    c4[0] = W4{ { t[0], m[0], m[2], m[0] } };
    c4[1] = W4{ { t[1], m[1], m[2], m[1] } };
    c4[2] = W4{ { t[2], m[0], m[1], m[2] } };
  }
  else
  {
    const uint32_t out[12] = { t[0], m[0], m[2], m[0], t[1], m[1], m[2], m[1], t[2], m[0], m[1], m[2] };
    for (int i = 0; i < 12; ++i)
      c[i] = out[i];
  }
}

/** pass C: the records of the vertices this node creates (sources were made by earlier levels) */
NB2_HD inline void midpointCreate(const MidpointLevel& lv, unsigned long long j, const unsigned* node_base, const uint8_t* node_mask,
                                  uint8_t* cloud, const Layout& L)
{
  const unsigned long long node = midpointNode(lv.shape, lv.level, j);
  const unsigned mask = node_mask[node];
  if (!mask)
    return;
  uint32_t v[3];
  for (int k = 0; k < 3; ++k)
    v[k] = midpointFinal(lv.tri[3 * j + k], lv.V, node_base, node_mask);
  unsigned made = 0;
  for (int k = 0; k < 3; ++k)
    if (mask & (1u << k))
    {
      const uint32_t target = lv.V + node_base[node] + made++;
      const uint8_t* src[2] = { cloud + static_cast<size_t>(v[k]) * L.step, cloud + static_cast<size_t>(v[k == 2 ? 0 : k + 1]) * L.step };
      makeMeanRecord(cloud + static_cast<size_t>(target) * L.step, src, 2, L);
    }
}

// ---- subdivision by area, one triangle of the level per call -----------------------------------------------------------
/** record of vertex `id`: the input cloud below V, the sweep's own records (creation order) above */
NB2_HD inline const uint8_t* areaRecord(uint32_t id, uint32_t V, const uint8_t* cloud, const uint8_t* fresh, uint32_t step)
{
  return id < V ? cloud + static_cast<size_t>(id) * step : fresh + static_cast<size_t>(id - V) * step;
}

/** returns 1 when the triangle must be split, 0 when it is kept; 2 when its area is not finite (:332-333) */
NB2_HD inline unsigned areaTest(const uint32_t* tri, unsigned long long j, uint32_t V, const uint8_t* cloud, const uint8_t* fresh,
                                const Layout& L, float max_area)
{
  F3 p[3];
  for (int k = 0; k < 3; ++k)
    p[k] = loadF3(areaRecord(tri[3 * j + k], V, cloud, fresh, L.step), L.off_xyz);
  const float tri_area = triangleArea(p[0], p[1], p[2]);
  if (!(fabsf(tri_area) <= 3.402823466e+38f))
    return 2u;
  float area = 0.0f;
  area += tri_area;
  return area > max_area ? 1u : 0u;
}

/** a split triangle: centroid record number V + base + rank_in_level, children in pop order
 *  { t0 m t2 } { m t1 t2 } { t0 t1 m }   (pushed in the opposite order, :297-299) */
NB2_HD inline void areaSplit(const uint32_t* tri, unsigned long long j, unsigned rank_in_level, uint32_t base, uint32_t V,
                             const uint8_t* cloud, uint8_t* fresh, const Layout& L, uint32_t* children)
{
  const uint32_t t[3] = { tri[3 * j], tri[3 * j + 1], tri[3 * j + 2] };
  const uint32_t serial = base + rank_in_level;
  const uint8_t* src[3];
  for (int k = 0; k < 3; ++k)
    src[k] = areaRecord(t[k], V, cloud, fresh, L.step);
  makeMeanRecord(fresh + static_cast<size_t>(serial) * L.step, src, 3, L);
  const uint32_t m = V + serial;
  const uint32_t out[9] = { t[0], m, t[2], m, t[1], t[2], t[0], t[1], m };
  uint32_t* c = children + 9ull * rank_in_level;
  for (int i = 0; i < 9; ++i)
    c[i] = out[i];
}

/** one level of the by-area sweep: n triangles; split[j] (0 / 1), pre = exclusive scan of split; leaves / inner = sizes
 *  of the subtree below each triangle; off_leaf / off_inner = leaves / centroids that precede it in the depth-first walk */
struct AreaLevel
{
  unsigned long long n;
  uint32_t base;  // centroids made by the levels above
  const uint32_t* tri;
  const unsigned* split;
  const unsigned* pre;
  unsigned* leaves;
  unsigned* inner;
  unsigned* off_leaf;
  unsigned* off_inner;
};

/** bottom-up: subtree sizes from those of the children (the level below) */
NB2_HD inline void areaUp(const AreaLevel& lv, const AreaLevel* below, unsigned long long j)
{
  if (!lv.split[j])
  {
    lv.leaves[j] = 1;
    lv.inner[j] = 0;
    return;
  }
  const unsigned long long c = 3ull * lv.pre[j];
  lv.leaves[j] = below->leaves[c] + below->leaves[c + 1] + below->leaves[c + 2];
  lv.inner[j] = 1 + below->inner[c] + below->inner[c + 1] + below->inner[c + 2];
}

/** top-down: a split triangle numbers its centroid and places its children */
NB2_HD inline void areaDown(const AreaLevel& lv, const AreaLevel* below, unsigned long long j, unsigned* rank)
{
  if (!lv.split[j])
    return;
  rank[lv.base + lv.pre[j]] = lv.off_inner[j];
  const unsigned long long c = 3ull * lv.pre[j];
  unsigned lo = lv.off_leaf[j], io = lv.off_inner[j] + 1;
  for (int k = 0; k < 3; ++k)
  {
    below->off_leaf[c + k] = lo;
    below->off_inner[c + k] = io;
    lo += below->leaves[c + k];
    io += below->inner[c + k];
  }
}

/** a kept triangle goes to its place in the output with the final vertex numbers */
NB2_HD inline void areaEmit(const AreaLevel& lv, unsigned long long j, uint32_t V, const unsigned* rank, uint32_t* out)
{
  if (lv.split[j])
    return;
  for (int k = 0; k < 3; ++k)
  {
    const uint32_t id = lv.tri[3 * j + k];
    out[3ull * lv.off_leaf[j] + k] = id < V ? id : V + rank[id - V];
  }
}
}  // namespace sub
}  // namespace nb2
