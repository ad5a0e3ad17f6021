// pcl/common/common.h — getMinMax3D (common/include/pcl/common/impl/common.hpp:295-400): axis-aligned bounds of a cloud,
// non-finite points of a non-dense cloud skipped.  Host code (VoxelGrid takes the same bounds on the device).
#pragma once
#include <cmath>
#include <limits>

#include "../eigen_lite.h"
#include "../point_cloud.h"
#include "../types.h"

namespace pcl {
namespace detail {
template <typename PointT, typename It>
inline void minMax3D(const pcl::PointCloud<PointT>& cloud, It first, It last, bool by_index, Eigen::Vector4f& min_pt, Eigen::Vector4f& max_pt)
{
  const float big = std::numeric_limits<float>::max();
  float mn[3] = {big, big, big}, mx[3] = {-big, -big, -big};
  for (It it = first; it != last; ++it) {
    const PointT& p = by_index ? cloud[static_cast<std::size_t>(*it)] : cloud[static_cast<std::size_t>(it - first)];
    if (!cloud.is_dense && !(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
    mn[0] = p.x < mn[0] ? p.x : mn[0]; mn[1] = p.y < mn[1] ? p.y : mn[1]; mn[2] = p.z < mn[2] ? p.z : mn[2];
    mx[0] = p.x > mx[0] ? p.x : mx[0]; mx[1] = p.y > mx[1] ? p.y : mx[1]; mx[2] = p.z > mx[2] ? p.z : mx[2];
  }
  for (int d = 0; d < 3; ++d) { min_pt[d] = mn[d]; max_pt[d] = mx[d]; }
  min_pt[3] = 1.f;   // the reference writes the homogeneous coordinate of its Array4f bounds
  max_pt[3] = 1.f;
}
}  // namespace detail

template <typename PointT>
inline void getMinMax3D(const pcl::PointCloud<PointT>& cloud, Eigen::Vector4f& min_pt, Eigen::Vector4f& max_pt)
{
  Indices all(cloud.size());
  for (std::size_t i = 0; i < all.size(); ++i) all[i] = static_cast<index_t>(i);
  detail::minMax3D(cloud, all.begin(), all.end(), true, min_pt, max_pt);
}
template <typename PointT>
inline void getMinMax3D(const pcl::PointCloud<PointT>& cloud, const Indices& indices, Eigen::Vector4f& min_pt, Eigen::Vector4f& max_pt)
{
  detail::minMax3D(cloud, indices.begin(), indices.end(), true, min_pt, max_pt);
}
template <typename PointT>
inline void getMinMax3D(const pcl::PointCloud<PointT>& cloud, PointT& min_pt, PointT& max_pt)
{
  Eigen::Vector4f mn, mx;
  getMinMax3D(cloud, mn, mx);
  min_pt.x = mn[0]; min_pt.y = mn[1]; min_pt.z = mn[2];
  max_pt.x = mx[0]; max_pt.y = mx[1]; max_pt.z = mx[2];
}
}  // namespace pcl
