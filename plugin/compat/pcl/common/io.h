// Minimal stand-in (see plugin/compat/README.md) for pcl::concatenateFields (pcl/common/io.h).
#pragma once
#include <pcl/PCLPointCloud2.h>

#include <cstring>

namespace pcl
{
/** cloud_out = all fields of cloud2, then the fields of cloud1 that cloud2 does not have (padding fields named "_"
 *  dropped), offsets of the latter shifted by cloud2.point_step; false when the clouds differ in size. */
inline bool concatenateFields(const PCLPointCloud2& cloud1, const PCLPointCloud2& cloud2, PCLPointCloud2& cloud_out)
{
  if (cloud1.width != cloud2.width || cloud1.height != cloud2.height || cloud1.is_bigendian != cloud2.is_bigendian)
    return false;
  std::vector<PCLPointField> extra;
  for (const auto& f : cloud1.fields)
  {
    if (f.name == "_")
      continue;
    bool clash = false;
    for (const auto& g : cloud2.fields)
      clash = clash || g.name == f.name;
    if (!clash)
      extra.push_back(f);
  }
  PCLPointCloud2 out;
  out.header = cloud2.header;
  out.width = cloud2.width;
  out.height = cloud2.height;
  out.is_bigendian = cloud2.is_bigendian;
  out.is_dense = cloud1.is_dense && cloud2.is_dense;
  out.fields = cloud2.fields;
  for (auto f : extra)
  {
    f.offset += cloud2.point_step;
    out.fields.push_back(f);
  }
  out.point_step = cloud2.point_step + cloud1.point_step;
  out.row_step = out.point_step * out.width;
  const std::size_t n = static_cast<std::size_t>(out.width) * out.height;
  out.data.assign(n * out.point_step, 0);
  for (std::size_t i = 0; i < n; ++i)
  {
    std::memcpy(out.data.data() + i * out.point_step, cloud2.data.data() + i * cloud2.point_step, cloud2.point_step);
    std::memcpy(out.data.data() + i * out.point_step + cloud2.point_step, cloud1.data.data() + i * cloud1.point_step,
                cloud1.point_step);
  }
  cloud_out = std::move(out);
  return true;
}
}  // namespace pcl
