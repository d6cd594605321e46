// Stand-in declaring noether::MeshModifier (noether_tpp/core/mesh_modifier.h:38); see plugin/compat/README.md.
#pragma once
#include <pcl/PolygonMesh.h>

#include <memory>
#include <vector>

namespace noether
{
struct MeshModifier
{
  using Ptr = std::unique_ptr<MeshModifier>;
  using ConstPtr = std::unique_ptr<const MeshModifier>;
  virtual ~MeshModifier() = default;
  virtual std::vector<pcl::PolygonMesh> modify(const pcl::PolygonMesh& mesh) const = 0;
};
}  // namespace noether
