// Minimal stand-in (see plugin/compat/README.md): same members as pcl::PCLPointCloud2.
#pragma once
#include <pcl/PCLPointField.h>

#include <vector>

namespace pcl
{
struct PCLPointCloud2
{
  PCLHeader header;
  uindex_t height = 0;
  uindex_t width = 0;
  std::vector<PCLPointField> fields;
  std::uint8_t is_bigendian = 0;
  uindex_t point_step = 0;
  uindex_t row_step = 0;
  std::vector<std::uint8_t> data;
  std::uint8_t is_dense = 0;
};
}  // namespace pcl
