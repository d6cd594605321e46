// TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product path.
//
// Sequential restatement of EuclideanClusteringMeshModifier::modify (euclidean_clustering_modifier.cpp:40-64) and of the
// third-party calls it makes (source not under /root/reference; written from PCL 1.12 / FLANN 1.9 behaviour):
//
//   pcl::EuclideanClusterExtraction<PointXYZ>::extract -> pcl::extractEuclideanClusters (segmentation/extract_clusters.hpp):
//     for every unprocessed point in index order: breadth-first growth through tree->radiusSearch(point, tolerance);
//     a queue whose size lies in [min, max] becomes a cluster (indices sorted ascending); afterwards
//     std::sort(clusters.rbegin(), clusters.rend(), comparePointClusters)  (a.size() < b.size(): largest first, order of
//     equal sizes whatever libstdc++'s introsort leaves — restated by making the very same call)
//   pcl::KdTreeFLANN::radiusSearch: FLANN L2_Simple distance in float, ((dx^2) + dy^2) + dz^2, a neighbour when it is
//     strictly smaller than float(radius * radius); the query point is its own neighbour; radius <= 0 finds nothing
//   extractSubMeshFromInlierVertices (subset_extraction/subset_extractor.cpp:22-56): faces whose first three vertices are
//     all in the cluster, in face order, then
//   pcl::surface::SimplificationRemoveUnusedVertices::simplify: no faces -> the output stays a default-constructed mesh
//     (no fields, no data); all points used -> the input cloud as it is; otherwise the used points in order of first use
//     by the kept faces, height 1, is_dense false
//
// The neighbour search itself is a brute-force scan for small clouds and a uniform grid otherwise (the oracle's own; it
// only has to find the same neighbour SETS as the k-d tree).
//
// PARITY STATUS: parity unpinned.  The reference's test (mesh_modifier_utest.cpp:15-19,130-183: tolerance 1.0 on the test
// mesh) only checks that every returned mesh has polygons, points and the input's fields.
#include "oracle.h"
#include "oracle_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>

namespace
{
struct SubMesh
{
  std::vector<uint32_t> indices;  // the cluster (ascending)
  std::vector<uint8_t> cloud;
  uint64_t n_points = 0;
  std::vector<uint32_t> tri;
  bool whole_input = false;  // all points used: the input cloud unchanged (organised dimensions and all)
  bool empty = false;        // no face survives: default-constructed output mesh
};
struct Clusters
{
  std::vector<SubMesh> meshes;
  std::vector<uint32_t> labels;  // smallest index of the connected component of every point (before the size filter)
};

inline bool near(const float* a, const float* b, float r2)
{
  float result = 0.f;
  for (int i = 0; i < 3; ++i)
  {
    const float diff = a[i] - b[i];
    result += diff * diff;
  }
  return result < r2;
}

struct Grid
{
  double h = 0, mn[3] = { 0, 0, 0 };
  std::unordered_map<uint64_t, std::vector<uint32_t>> cells;
  static uint64_t key(int64_t x, int64_t y, int64_t z) { return (uint64_t(x) * 0x9e3779b97f4a7c15ull) ^ (uint64_t(y) * 0xc2b2ae3d27d4eb4full) ^ (uint64_t(z) * 0x165667b19e3779f9ull); }
  void cell(const float* p, int64_t c[3]) const
  {
    for (int i = 0; i < 3; ++i)
      c[i] = static_cast<int64_t>(std::floor((double(p[i]) - mn[i]) / h));
  }
};
}  // namespace

extern "C" {

int orc_euclidean_clustering(const uint8_t* cloud, uint64_t n_points, uint32_t point_step, uint32_t off_xyz, const uint32_t* tri,
                             uint64_t n_triangles, double tolerance, int min_cluster_size, int max_cluster_size, int force_brute,
                             void** out)
{
  if (!cloud || !out || (n_triangles && !tri) || n_points == 0)
  {
    orc::setError("invalid argument");
    return ORC_ERR_INVALID;
  }
  for (uint64_t i = 0; i < 3 * n_triangles; ++i)
    if (tri[i] >= n_points)
    {
      orc::setError("triangle index out of range");
      return ORC_ERR_INVALID;
    }
  const uint64_t V = n_points;
  std::vector<float> xyz(3 * V);
  for (uint64_t v = 0; v < V; ++v)
    std::memcpy(&xyz[3 * v], cloud + v * point_step + off_xyz, 12);
  const float r2 = static_cast<float>(tolerance * tolerance);
  const bool searchable = tolerance > 0.0;

  // neighbour search
  Grid grid;
  const bool brute = force_brute || V <= 2048 || !searchable;
  if (!brute)
  {
    grid.h = tolerance * 1.001;
    for (int i = 0; i < 3; ++i)
    {
      float m = xyz[i];
      for (uint64_t v = 1; v < V; ++v)
        m = std::min(m, xyz[3 * v + i]);
      grid.mn[i] = m;
    }
    for (uint64_t v = 0; v < V; ++v)
    {
      int64_t c[3];
      grid.cell(&xyz[3 * v], c);
      grid.cells[Grid::key(c[0], c[1], c[2])].push_back(static_cast<uint32_t>(v));
    }
  }
  auto neighbours = [&](uint32_t p, std::vector<uint32_t>& nn) {
    nn.clear();
    if (!searchable)
      return;
    if (brute)
    {
      for (uint64_t q = 0; q < V; ++q)
        if (near(&xyz[3 * p], &xyz[3 * q], r2))
          nn.push_back(static_cast<uint32_t>(q));
      return;
    }
    int64_t c[3];
    grid.cell(&xyz[3 * p], c);
    for (int64_t dz = -1; dz <= 1; ++dz)
      for (int64_t dy = -1; dy <= 1; ++dy)
        for (int64_t dx = -1; dx <= 1; ++dx)
        {
          auto it = grid.cells.find(Grid::key(c[0] + dx, c[1] + dy, c[2] + dz));
          if (it == grid.cells.end())
            continue;
          for (uint32_t q : it->second)
            if (near(&xyz[3 * p], &xyz[3 * q], r2))
              nn.push_back(q);
        }
    std::sort(nn.begin(), nn.end());
    nn.erase(std::unique(nn.begin(), nn.end()), nn.end());  // hash collisions may list a cell twice
  };

  // pcl::extractEuclideanClusters
  auto* result = new Clusters();
  result->labels.assign(V, 0);
  std::vector<std::vector<uint32_t>> clusters;
  {
    const uint32_t min_pts = static_cast<uint32_t>(min_cluster_size);  // setMinClusterSize takes pcl::uindex_t
    const uint32_t max_pts = max_cluster_size <= 0 ? static_cast<uint32_t>(V) : static_cast<uint32_t>(max_cluster_size);  // :53
    std::vector<bool> processed(V, false);
    std::vector<uint32_t> nn;
    for (uint64_t i = 0; i < V; ++i)
    {
      if (processed[i])
        continue;
      std::vector<uint32_t> seed_queue{ static_cast<uint32_t>(i) };
      processed[i] = true;
      for (size_t sq = 0; sq < seed_queue.size(); ++sq)
      {
        neighbours(seed_queue[sq], nn);
        for (uint32_t q : nn)
          if (!processed[q])
          {
            seed_queue.push_back(q);
            processed[q] = true;
          }
      }
      for (uint32_t q : seed_queue)
        result->labels[q] = static_cast<uint32_t>(i);  // i is the smallest index of its component
      if (seed_queue.size() >= min_pts && seed_queue.size() <= max_pts)
      {
        std::sort(seed_queue.begin(), seed_queue.end());
        clusters.push_back(std::move(seed_queue));
      }
    }
    std::sort(clusters.rbegin(), clusters.rend(),
              [](const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) { return a.size() < b.size(); });
  }

  // extractSubMeshFromInlierVertices + SimplificationRemoveUnusedVertices
  for (auto& indices : clusters)
  {
    SubMesh m;
    std::vector<bool> whitelist(V, false);
    for (uint32_t v : indices)
      whitelist[v] = true;
    std::vector<uint32_t> kept;
    for (uint64_t f = 0; f < n_triangles; ++f)
      if (whitelist[tri[3 * f]] && whitelist[tri[3 * f + 1]] && whitelist[tri[3 * f + 2]])
        kept.insert(kept.end(), tri + 3 * f, tri + 3 * f + 3);
    m.indices = std::move(indices);
    if (kept.empty())
    {
      m.empty = true;
      result->meshes.push_back(std::move(m));
      continue;
    }
    std::vector<int64_t> new_index(V, -1);
    std::vector<uint32_t> used;
    for (uint32_t v : kept)
      if (new_index[v] < 0)
      {
        new_index[v] = static_cast<int64_t>(used.size());
        used.push_back(v);
      }
    if (used.size() == V)
    {
      m.whole_input = true;
      m.n_points = V;
      m.cloud.assign(cloud, cloud + V * point_step);
      m.tri = std::move(kept);
    }
    else
    {
      m.n_points = used.size();
      m.cloud.resize(used.size() * point_step);
      for (size_t i = 0; i < used.size(); ++i)
        std::memcpy(m.cloud.data() + i * point_step, cloud + static_cast<size_t>(used[i]) * point_step, point_step);
      m.tri.resize(kept.size());
      for (size_t i = 0; i < kept.size(); ++i)
        m.tri[i] = static_cast<uint32_t>(new_index[kept[i]]);
    }
    result->meshes.push_back(std::move(m));
  }
  *out = result;
  return ORC_OK;
}

uint64_t orc_clusters_count(const void* h) { return static_cast<const Clusters*>(h)->meshes.size(); }
void orc_clusters_labels(const void* h, uint32_t* labels)
{
  const auto* c = static_cast<const Clusters*>(h);
  std::memcpy(labels, c->labels.data(), c->labels.size() * 4);
}
/* out: cluster size, points of the sub-mesh, triangles of the sub-mesh, whole_input, empty */
void orc_cluster_info(const void* h, uint64_t k, uint64_t out[5])
{
  const SubMesh& m = static_cast<const Clusters*>(h)->meshes[k];
  out[0] = m.indices.size();
  out[1] = m.n_points;
  out[2] = m.tri.size() / 3;
  out[3] = m.whole_input;
  out[4] = m.empty;
}
void orc_cluster_export(const void* h, uint64_t k, uint32_t* indices, uint8_t* cloud, uint32_t* tri)
{
  const SubMesh& m = static_cast<const Clusters*>(h)->meshes[k];
  if (indices)
    std::memcpy(indices, m.indices.data(), m.indices.size() * 4);
  if (cloud)
    std::memcpy(cloud, m.cloud.data(), m.cloud.size());
  if (tri)
    std::memcpy(tri, m.tri.data(), m.tri.size() * 4);
}
void orc_clusters_free(void* h) { delete static_cast<Clusters*>(h); }
}
