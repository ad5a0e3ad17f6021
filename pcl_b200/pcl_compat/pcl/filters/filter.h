// pcl/filters/filter.h — removeNaNFromPointCloud (filters/include/pcl/filters/impl/filter.hpp:45-92): the usual step before
// a registration on sensor data.  Host code.
#pragma once
#include <cmath>

#include "../point_cloud.h"
#include "../types.h"

namespace pcl {
// cloud_out = the points of cloud_in with finite x, y, z (in place allowed); index[j] = position in cloud_in of output
// point j.  A dense input is copied as is.  The result is unorganised and dense.
template <typename PointT>
inline void removeNaNFromPointCloud(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out, Indices& index)
{
  if (&cloud_in != &cloud_out) {
    cloud_out.header = cloud_in.header;
    cloud_out.points.resize(cloud_in.size());
    cloud_out.sensor_origin_ = cloud_in.sensor_origin_;
    cloud_out.sensor_orientation_ = cloud_in.sensor_orientation_;
  }
  index.resize(cloud_in.size());
  if (cloud_in.is_dense) {
    if (&cloud_in != &cloud_out) {
      cloud_out.points = cloud_in.points;
      cloud_out.width = cloud_in.width;
      cloud_out.height = cloud_in.height;
      cloud_out.is_dense = true;
    }
    for (std::size_t j = 0; j < index.size(); ++j) index[j] = static_cast<index_t>(j);
    return;
  }
  std::size_t j = 0;
  for (std::size_t i = 0; i < cloud_in.size(); ++i) {
    const PointT p = cloud_in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    cloud_out.points[j] = p;
    index[j] = static_cast<index_t>(i);
    ++j;
  }
  if (j != cloud_in.size()) {
    cloud_out.points.resize(j);
    index.resize(j);
  }
  cloud_out.height = 1;
  cloud_out.width = static_cast<std::uint32_t>(j);
  cloud_out.is_dense = true;
}
}  // namespace pcl
