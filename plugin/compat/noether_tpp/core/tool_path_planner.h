// Stand-in declaring noether::ToolPathPlanner (noether_tpp/core/tool_path_planner.h:37); see plugin/compat/README.md.
#pragma once
#include <noether_tpp/core/types.h>
#include <pcl/PolygonMesh.h>

#include <memory>

namespace noether
{
struct ToolPathPlanner
{
  using Ptr = std::unique_ptr<ToolPathPlanner>;
  using ConstPtr = std::unique_ptr<const ToolPathPlanner>;
  virtual ~ToolPathPlanner() = default;
  virtual ToolPaths plan(const pcl::PolygonMesh& mesh) const = 0;
};
}  // namespace noether
