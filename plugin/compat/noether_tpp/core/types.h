// Stand-in declaring the reference's tool path types (noether_tpp/core/types.h:31-50); see plugin/compat/README.md.
#pragma once
#include <Eigen/Geometry>

#include <vector>

namespace noether
{
using ToolPathWaypoint = Eigen::Isometry3d;
using ToolPathSegment = std::vector<Eigen::Isometry3d, Eigen::aligned_allocator<Eigen::Isometry3d>>;
using ToolPath = std::vector<ToolPathSegment>;
using ToolPaths = std::vector<ToolPath>;
}  // namespace noether
