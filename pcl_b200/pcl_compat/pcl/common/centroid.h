// pcl/common/centroid.h — host side of the moment helpers on the path (common/include/pcl/common/impl/centroid.hpp):
// compute3DCentroid (:55-137), computeMeanAndCovarianceMatrix (:578-652, the single-pass form shifted by the first
// finite point) and demeanPointCloud (:933-1023).  The device evaluates the same sums inside its kernels
// (search.cu: moments_add; icp.cu: the correlation estimator); these are the public functions a PCL caller can reach.
#pragma once
#include <cmath>
#include <cstddef>

#include "../eigen_lite.h"
#include "../point_cloud.h"
#include "../types.h"

namespace pcl {
namespace detail {
template <typename PointT>
inline bool xyzFinite(const PointT& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }
// visits cloud[indices[j]] (or cloud[j] when indices == nullptr)
template <typename PointT, typename F>
inline void forEachPoint(const pcl::PointCloud<PointT>& cloud, const Indices* indices, F&& f)
{
  if (indices)
    for (index_t i : *indices) f(cloud[static_cast<std::size_t>(i)]);
  else
    for (const PointT& p : cloud.points) f(p);
}
template <typename PointT, typename Scalar>
inline unsigned int centroid3D(const pcl::PointCloud<PointT>& cloud, const Indices* indices, Eigen::Matrix<Scalar, 4, 1>& centroid)
{
  Eigen::Matrix<Scalar, 4, 1> accumulator = Eigen::Matrix<Scalar, 4, 1>::Zero();
  unsigned int cp = 0;
  forEachPoint(cloud, indices, [&](const PointT& p) {
    if (!cloud.is_dense && !xyzFinite(p)) return;
    accumulator[0] += p.x; accumulator[1] += p.y; accumulator[2] += p.z;
    ++cp;
  });
  if (cp == 0) return 0;  // impl/centroid.hpp:77-82: the caller's centroid is written only when a point was used
  for (int d = 0; d < 3; ++d) centroid[d] = accumulator[d] / static_cast<Scalar>(cp);
  centroid[3] = Scalar(1);
  return cp;
}
template <typename PointT, typename Scalar>
inline unsigned int meanAndCovariance(const pcl::PointCloud<PointT>& cloud, const Indices* indices,
                                      Eigen::Matrix<Scalar, 3, 3>& covariance_matrix, Eigen::Matrix<Scalar, 4, 1>& centroid)
{
  // shift by the first finite point K: the sums stay small and the single pass loses little to cancellation
  Scalar K[3] = {0, 0, 0};
  bool have_k = false;
  forEachPoint(cloud, indices, [&](const PointT& p) {
    if (!have_k && xyzFinite(p)) {
      K[0] = p.x; K[1] = p.y; K[2] = p.z;
      have_k = true;
    }
  });
  Scalar accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned int point_count = 0;
  forEachPoint(cloud, indices, [&](const PointT& p) {
    if (!cloud.is_dense && !xyzFinite(p)) return;
    const Scalar x = p.x - K[0], y = p.y - K[1], z = p.z - K[2];
    accu[0] += x * x; accu[1] += x * y; accu[2] += x * z;
    accu[3] += y * y; accu[4] += y * z; accu[5] += z * z;
    accu[6] += x; accu[7] += y; accu[8] += z;
    ++point_count;
  });
  if (point_count == 0) return 0;
  for (Scalar& a : accu) a /= static_cast<Scalar>(point_count);
  centroid[0] = accu[6] + K[0]; centroid[1] = accu[7] + K[1]; centroid[2] = accu[8] + K[2];
  centroid[3] = Scalar(1);
  covariance_matrix(0, 0) = accu[0] - accu[6] * accu[6];
  covariance_matrix(0, 1) = accu[1] - accu[6] * accu[7];
  covariance_matrix(0, 2) = accu[2] - accu[6] * accu[8];
  covariance_matrix(1, 1) = accu[3] - accu[7] * accu[7];
  covariance_matrix(1, 2) = accu[4] - accu[7] * accu[8];
  covariance_matrix(2, 2) = accu[5] - accu[8] * accu[8];
  covariance_matrix(1, 0) = covariance_matrix(0, 1);
  covariance_matrix(2, 0) = covariance_matrix(0, 2);
  covariance_matrix(2, 1) = covariance_matrix(1, 2);
  return point_count;
}
}  // namespace detail

// centroid of the finite points; returns how many were used (0: the caller's centroid is left as it was)
template <typename PointT, typename Scalar>
inline unsigned int compute3DCentroid(const pcl::PointCloud<PointT>& cloud, Eigen::Matrix<Scalar, 4, 1>& centroid)
{
  return detail::centroid3D(cloud, nullptr, centroid);
}
template <typename PointT, typename Scalar>
inline unsigned int compute3DCentroid(const pcl::PointCloud<PointT>& cloud, const Indices& indices, Eigen::Matrix<Scalar, 4, 1>& centroid)
{
  return detail::centroid3D(cloud, &indices, centroid);
}

// normalised (divided by n) covariance and the centroid in one pass; returns the number of points used
template <typename PointT, typename Scalar>
inline unsigned int computeMeanAndCovarianceMatrix(const pcl::PointCloud<PointT>& cloud, Eigen::Matrix<Scalar, 3, 3>& covariance_matrix,
                                                   Eigen::Matrix<Scalar, 4, 1>& centroid)
{
  return detail::meanAndCovariance(cloud, nullptr, covariance_matrix, centroid);
}
template <typename PointT, typename Scalar>
inline unsigned int computeMeanAndCovarianceMatrix(const pcl::PointCloud<PointT>& cloud, const Indices& indices,
                                                   Eigen::Matrix<Scalar, 3, 3>& covariance_matrix, Eigen::Matrix<Scalar, 4, 1>& centroid)
{
  return detail::meanAndCovariance(cloud, &indices, covariance_matrix, centroid);
}

// cloud_out = cloud_in (all fields) with the centroid subtracted from x, y, z; the index form writes x, y, z only
template <typename PointT, typename Scalar>
inline void demeanPointCloud(const pcl::PointCloud<PointT>& cloud_in, const Eigen::Matrix<Scalar, 4, 1>& centroid,
                             pcl::PointCloud<PointT>& cloud_out)
{
  cloud_out = cloud_in;
  for (PointT& p : cloud_out.points) {
    p.x -= static_cast<float>(centroid[0]); p.y -= static_cast<float>(centroid[1]); p.z -= static_cast<float>(centroid[2]);
  }
}
template <typename PointT, typename Scalar>
inline void demeanPointCloud(const pcl::PointCloud<PointT>& cloud_in, const Indices& indices, const Eigen::Matrix<Scalar, 4, 1>& centroid,
                             pcl::PointCloud<PointT>& cloud_out)
{
  cloud_out.header = cloud_in.header;
  cloud_out.is_dense = cloud_in.is_dense;
  cloud_out.points.resize(indices.size());
  if (indices.size() == cloud_in.size()) { cloud_out.width = cloud_in.width; cloud_out.height = cloud_in.height; }
  else { cloud_out.width = static_cast<std::uint32_t>(indices.size()); cloud_out.height = 1; }
  for (std::size_t i = 0; i < indices.size(); ++i) {  // only x, y, z are written; the difference is taken in Scalar (:1003-1008)
    const PointT& q = cloud_in[static_cast<std::size_t>(indices[i])];
    cloud_out.points[i].x = static_cast<float>(q.x - centroid[0]);
    cloud_out.points[i].y = static_cast<float>(q.y - centroid[1]);
    cloud_out.points[i].z = static_cast<float>(q.z - centroid[2]);
  }
}
}  // namespace pcl
