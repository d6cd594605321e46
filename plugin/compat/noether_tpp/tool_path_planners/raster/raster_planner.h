// Stand-in declaring the reference's raster planner interfaces (raster_planner.h:25-88); see plugin/compat/README.md.
#pragma once
#include <noether_tpp/core/tool_path_planner.h>

#include <Eigen/Geometry>

namespace noether
{
struct DirectionGenerator
{
  using Ptr = std::unique_ptr<DirectionGenerator>;
  using ConstPtr = std::unique_ptr<const DirectionGenerator>;
  virtual ~DirectionGenerator() = default;
  virtual Eigen::Vector3d generate(const pcl::PolygonMesh& mesh) const = 0;
};

struct OriginGenerator
{
  using Ptr = std::unique_ptr<OriginGenerator>;
  using ConstPtr = std::unique_ptr<const OriginGenerator>;
  virtual ~OriginGenerator() = default;
  virtual Eigen::Vector3d generate(const pcl::PolygonMesh& mesh) const = 0;
};

class RasterPlanner : public ToolPathPlanner
{
public:
  RasterPlanner(DirectionGenerator::ConstPtr dir_gen, OriginGenerator::ConstPtr origin_gen)
    : dir_gen_(std::move(dir_gen)), origin_gen_(std::move(origin_gen))
  {
  }
  /** The real implementation (raster_planner.cpp:16-35) post-processes planImpl's result with
   *  RasterOrganizationModifier and DirectionOfTravelOrientationModifier; this stand-in returns it unchanged. */
  ToolPaths plan(const pcl::PolygonMesh& mesh) const override final { return planImpl(mesh); }
  void setLineSpacing(const double line_spacing) { line_spacing_ = line_spacing; }

protected:
  virtual ToolPaths planImpl(const pcl::PolygonMesh& mesh) const = 0;
  DirectionGenerator::ConstPtr dir_gen_;
  OriginGenerator::ConstPtr origin_gen_;
  double line_spacing_;
};
}  // namespace noether
