// pcl/PointIndices.h — common/include/pcl/PointIndices.h:12-25
#pragma once
#include <memory>

#include "point_cloud.h"
#include "types.h"

namespace pcl {
struct PointIndices {
  using Ptr = std::shared_ptr<PointIndices>;
  using ConstPtr = std::shared_ptr<const PointIndices>;
  PCLHeader header;
  Indices indices;
};
using PointIndicesPtr = PointIndices::Ptr;
using PointIndicesConstPtr = PointIndices::ConstPtr;
}  // namespace pcl
