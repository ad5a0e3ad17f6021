// pcl/common/transforms.h — pcl::transformPointCloud / transformPointCloudWithNormals / transformPoint
// (common/include/pcl/common/transforms.h:60-330, impl/transforms.hpp:60-360): apply a rigid 4x4 to a cloud on the host —
// what a caller does with IterativeClosestPoint::getFinalTransformation().  The per-point arithmetic follows the
// reference's Transformer: products first, then the sum nested from the right, x' = m00 x + (m01 y + (m02 z + m03)),
// in the transform's Scalar, rounded to float once.  Non-finite points of a non-dense cloud are left untouched.
#pragma once
#include <cmath>
#include <cstddef>

#include "../eigen_lite.h"
#include "../point_cloud.h"
#include "../point_types.h"
#include "../types.h"

namespace pcl {
namespace detail {
template <typename Scalar>
struct Transformer {
  const Eigen::Matrix<Scalar, 4, 4>& tf;
  explicit Transformer(const Eigen::Matrix<Scalar, 4, 4>& transform) : tf(transform) {}
  // rotation only (normals): dst[3] = 0
  void so3(const float* src, float* dst) const
  {
    const Scalar p[3] = {static_cast<Scalar>(src[0]), static_cast<Scalar>(src[1]), static_cast<Scalar>(src[2])};
    for (int r = 0; r < 3; ++r) dst[r] = static_cast<float>(tf(r, 0) * p[0] + (tf(r, 1) * p[1] + tf(r, 2) * p[2]));
    dst[3] = 0.f;
  }
  // rotation + translation (coordinates): dst[3] = 1
  void se3(const float* src, float* dst) const
  {
    const Scalar p[3] = {static_cast<Scalar>(src[0]), static_cast<Scalar>(src[1]), static_cast<Scalar>(src[2])};
    for (int r = 0; r < 3; ++r) dst[r] = static_cast<float>(tf(r, 0) * p[0] + (tf(r, 1) * p[1] + (tf(r, 2) * p[2] + tf(r, 3))));
    dst[3] = 1.f;
  }
};
template <typename PointT>
inline bool finiteXYZ(const PointT& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }
template <typename PointT>
inline void prepareOutput(const pcl::PointCloud<PointT>& in, pcl::PointCloud<PointT>& out, bool copy_all_fields)
{
  if (&in == &out) return;
  out.header = in.header;
  out.is_dense = in.is_dense;
  if (copy_all_fields) out.points = in.points;   // every field, then x, y, z (and the normal) are overwritten
  else out.points.assign(in.size(), PointT());
  out.width = in.width;
  out.height = in.height;
  out.sensor_orientation_ = in.sensor_orientation_;
  out.sensor_origin_ = in.sensor_origin_;
}
}  // namespace detail

template <typename PointT, typename Scalar>
inline void transformPointCloud(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out,
                                const Eigen::Matrix<Scalar, 4, 4>& transform, bool copy_all_fields = true)
{
  detail::prepareOutput(cloud_in, cloud_out, copy_all_fields);
  const detail::Transformer<Scalar> tf(transform);
  for (std::size_t i = 0; i < cloud_in.size(); ++i) {
    if (!cloud_in.is_dense && !detail::finiteXYZ(cloud_in[i])) continue;
    tf.se3(cloud_in[i].data, cloud_out[i].data);
  }
}
template <typename PointT, typename Scalar>
inline void transformPointCloud(const pcl::PointCloud<PointT>& cloud_in, const Indices& indices, pcl::PointCloud<PointT>& cloud_out,
                                const Eigen::Matrix<Scalar, 4, 4>& transform, bool copy_all_fields = true)
{
  const std::size_t n = indices.size();
  pcl::PointCloud<PointT> out;
  out.header = cloud_in.header;
  out.is_dense = cloud_in.is_dense;
  out.sensor_orientation_ = cloud_in.sensor_orientation_;
  out.sensor_origin_ = cloud_in.sensor_origin_;
  out.points.assign(n, PointT());
  const detail::Transformer<Scalar> tf(transform);
  for (std::size_t i = 0; i < n; ++i) {
    const PointT& src = cloud_in[static_cast<std::size_t>(indices[i])];
    if (copy_all_fields) out.points[i] = src;
    if (!cloud_in.is_dense && !detail::finiteXYZ(src)) continue;
    tf.se3(src.data, out.points[i].data);
  }
  out.width = static_cast<std::uint32_t>(n);
  out.height = 1;
  cloud_out = std::move(out);
}
// coordinates and normals (point types with normal_x/y/z: the rotation is applied to the normal)
template <typename PointT, typename Scalar>
inline void transformPointCloudWithNormals(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out,
                                           const Eigen::Matrix<Scalar, 4, 4>& transform, bool copy_all_fields = true)
{
  detail::prepareOutput(cloud_in, cloud_out, copy_all_fields);
  const detail::Transformer<Scalar> tf(transform);
  for (std::size_t i = 0; i < cloud_in.size(); ++i) {
    if (!cloud_in.is_dense && !detail::finiteXYZ(cloud_in[i])) continue;
    tf.se3(cloud_in[i].data, cloud_out[i].data);
    tf.so3(cloud_in[i].data_n, cloud_out[i].data_n);
  }
}
// transforms.h:260-300: the indexed form — output of indices.size() points, unorganized
template <typename PointT, typename Scalar>
inline void transformPointCloudWithNormals(const pcl::PointCloud<PointT>& cloud_in, const Indices& indices, pcl::PointCloud<PointT>& cloud_out,
                                           const Eigen::Matrix<Scalar, 4, 4>& transform, bool copy_all_fields = true)
{
  const std::size_t n = indices.size();
  pcl::PointCloud<PointT> out;
  out.header = cloud_in.header;
  out.is_dense = cloud_in.is_dense;
  out.sensor_orientation_ = cloud_in.sensor_orientation_;
  out.sensor_origin_ = cloud_in.sensor_origin_;
  out.points.assign(n, PointT());
  const detail::Transformer<Scalar> tf(transform);
  for (std::size_t i = 0; i < n; ++i) {
    const PointT& src = cloud_in[static_cast<std::size_t>(indices[i])];
    if (copy_all_fields) out.points[i] = src;
    if (!cloud_in.is_dense && !detail::finiteXYZ(src)) continue;
    tf.se3(src.data, out.points[i].data);
    tf.so3(src.data_n, out.points[i].data_n);
  }
  out.width = static_cast<std::uint32_t>(n);
  out.height = 1;
  cloud_out = std::move(out);
}
template <typename PointT, typename Scalar>
inline PointT transformPoint(const PointT& point, const Eigen::Matrix<Scalar, 4, 4>& transform)
{
  PointT ret = point;
  detail::Transformer<Scalar>(transform).se3(point.data, ret.data);
  return ret;
}
template <typename PointT, typename Scalar>
inline PointT transformPointWithNormal(const PointT& point, const Eigen::Matrix<Scalar, 4, 4>& transform)
{
  PointT ret = point;
  const detail::Transformer<Scalar> tf(transform);
  tf.se3(point.data, ret.data);
  tf.so3(point.data_n, ret.data_n);
  return ret;
}
}  // namespace pcl
