// Minimal stand-in (see plugin/compat/README.md): same members as pcl::Vertices / pcl::PolygonMesh.
#pragma once
#include <pcl/PCLPointCloud2.h>

#include <vector>

namespace pcl
{
struct Vertices
{
  std::vector<index_t> vertices;
};
struct PolygonMesh
{
  PCLHeader header;
  PCLPointCloud2 cloud;
  std::vector<Vertices> polygons;
};
}  // namespace pcl
