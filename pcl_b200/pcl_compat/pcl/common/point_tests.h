// pcl/common/point_tests.h — pcl::isFinite / isXYZFinite / isNormalFinite (common/include/pcl/common/point_tests.h:55-150);
// the first two live with the point types here, this header adds the normal test and is what PCL programs include
#pragma once
#include <cmath>

#include "../point_types.h"

namespace pcl {
template <typename PointT>
inline bool isNormalFinite(const PointT& pt)
{
  return std::isfinite(pt.normal_x) && std::isfinite(pt.normal_y) && std::isfinite(pt.normal_z);
}
}  // namespace pcl
