// Minimal stand-in (see plugin/compat/README.md): same members as pcl::PCLPointField.
#pragma once
#include <cstdint>
#include <string>

namespace pcl
{
using uindex_t = std::uint32_t;
using index_t = std::uint32_t;
struct PCLHeader
{
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;
  std::string frame_id;
};
struct PCLPointField
{
  std::string name;
  uindex_t offset = 0;
  std::uint8_t datatype = 0;
  uindex_t count = 0;
  enum PointFieldTypes
  {
    INT8 = 1,
    UINT8 = 2,
    INT16 = 3,
    UINT16 = 4,
    INT32 = 5,
    UINT32 = 6,
    FLOAT32 = 7,
    FLOAT64 = 8
  };
};
}  // namespace pcl
