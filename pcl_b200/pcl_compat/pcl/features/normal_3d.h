// pcl/features/normal_3d.h — pcl::NormalEstimation[OMP]<PointInT, PointOutT> (setKSearch or setRadiusSearch) on the device
// (features/include/pcl/features/normal_3d.h:242-415, impl/normal_3d.hpp:47-96, impl/feature.hpp:95-230).
#pragma once
#include <cstdio>
#include <limits>

#include "../common/centroid.h"
#include "../point_types.h"
#include "../search/kdtree.h"
#include "feature.h"

namespace pcl {
// normal_3d.h:60-111: plane (Hessian form) and curvature of a set of points, on the host.  Fewer than three points or no
// finite one: NaNs and false.
template <typename PointT>
inline bool computePointNormal(const pcl::PointCloud<PointT>& cloud, Eigen::Vector4f& plane_parameters, float& curvature)
{
  Eigen::Matrix3f covariance_matrix;
  Eigen::Vector4f xyz_centroid;
  if (cloud.size() < 3 || computeMeanAndCovarianceMatrix(cloud, covariance_matrix, xyz_centroid) == 0) {
    for (int d = 0; d < 4; ++d) plane_parameters[d] = std::numeric_limits<float>::quiet_NaN();
    curvature = std::numeric_limits<float>::quiet_NaN();
    return false;
  }
  solvePlaneParameters(covariance_matrix, xyz_centroid, plane_parameters, curvature);
  return true;
}
template <typename PointT>
inline bool computePointNormal(const pcl::PointCloud<PointT>& cloud, const Indices& indices, Eigen::Vector4f& plane_parameters,
                               float& curvature)
{
  Eigen::Matrix3f covariance_matrix;
  Eigen::Vector4f xyz_centroid;
  if (indices.size() < 3 || computeMeanAndCovarianceMatrix(cloud, indices, covariance_matrix, xyz_centroid) == 0) {
    for (int d = 0; d < 4; ++d) plane_parameters[d] = std::numeric_limits<float>::quiet_NaN();
    curvature = std::numeric_limits<float>::quiet_NaN();
    return false;
  }
  solvePlaneParameters(covariance_matrix, xyz_centroid, plane_parameters, curvature);
  return true;
}
// normal_3d.h:113-188: turn a normal towards the viewpoint; the 4-vector form also re-derives d through the point
template <typename PointT, typename Scalar>
inline void flipNormalTowardsViewpoint(const PointT& point, float vp_x, float vp_y, float vp_z, Eigen::Matrix<Scalar, 4, 1>& normal)
{
  const Scalar vx = vp_x - point.x, vy = vp_y - point.y, vz = vp_z - point.z;
  const float cos_theta = static_cast<float>(vx * normal[0] + vy * normal[1] + vz * normal[2] + Scalar(0) * normal[3]);
  if (cos_theta < 0) {
    for (int d = 0; d < 4; ++d) normal[d] *= -1;
    normal[3] = 0.0f;
    normal[3] = -1 * (normal[0] * point.x + normal[1] * point.y + normal[2] * point.z + normal[3] * Scalar(1));
  }
}
template <typename PointT, typename Scalar>
inline void flipNormalTowardsViewpoint(const PointT& point, float vp_x, float vp_y, float vp_z, Eigen::Matrix<Scalar, 3, 1>& normal)
{
  const Scalar vx = vp_x - point.x, vy = vp_y - point.y, vz = vp_z - point.z;
  if (vx * normal[0] + vy * normal[1] + vz * normal[2] < 0)
    for (int d = 0; d < 3; ++d) normal[d] *= -1;
}
template <typename PointT>
inline void flipNormalTowardsViewpoint(const PointT& point, float vp_x, float vp_y, float vp_z, float& nx, float& ny, float& nz)
{
  vp_x -= point.x;
  vp_y -= point.y;
  vp_z -= point.z;
  const float cos_theta = (vp_x * nx + vp_y * ny + vp_z * nz);
  if (cos_theta < 0) {
    nx *= -1;
    ny *= -1;
    nz *= -1;
  }
}

template <typename PointInT, typename PointOutT>
class NormalEstimation : public PCLBase<PointInT> {
public:
  using PointCloudOut = pcl::PointCloud<PointOutT>;
  using KdTreePtr = typename pcl::search::KdTree<PointInT>::Ptr;
  using PointCloudInConstPtr = typename pcl::PointCloud<PointInT>::ConstPtr;

  void setInputCloud(const PointCloudInConstPtr& cloud) override
  {
    this->input_ = cloud;
    if (use_sensor_origin_ && cloud) {  // normal_3d.h:328-337
      vp_[0] = cloud->sensor_origin_[0];
      vp_[1] = cloud->sensor_origin_[1];
      vp_[2] = cloud->sensor_origin_[2];
    }
  }
  void setSearchSurface(const PointCloudInConstPtr& cloud) { surface_ = cloud; fake_surface_ = false; }
  PointCloudInConstPtr getSearchSurface() const { return surface_; }   // feature.h:146-150
  double getSearchParameter() const { return search_radius_ != 0.0 ? search_radius_ : static_cast<double>(k_); }  // feature.h:170-174 (after compute)
  // normal_3d.h:279-322: one neighbourhood on the host (the batch form is compute())
  bool computePointNormal(const pcl::PointCloud<PointInT>& cloud, const Indices& indices, Eigen::Vector4f& plane_parameters,
                          float& curvature)
  {
    return pcl::computePointNormal(cloud, indices, plane_parameters, curvature);
  }
  bool computePointNormal(const pcl::PointCloud<PointInT>& cloud, const Indices& indices, float& nx, float& ny, float& nz,
                          float& curvature)
  {
    Eigen::Matrix3f covariance_matrix;
    Eigen::Vector4f xyz_centroid;
    if (indices.size() < 3 || computeMeanAndCovarianceMatrix(cloud, indices, covariance_matrix, xyz_centroid) == 0) {
      nx = ny = nz = curvature = std::numeric_limits<float>::quiet_NaN();
      return false;
    }
    solvePlaneParameters(covariance_matrix, nx, ny, nz, curvature);
    return true;
  }
  void setSearchMethod(const KdTreePtr& tree) { tree_ = tree; }
  // feature.h:114,164-168: the reference's parameter type is pcl::search::Search<PointInT>::Ptr
  void setSearchMethod(const typename pcl::search::Search<PointInT>::Ptr& tree)
  {
    tree_ = pcl::search::deviceSearcher<PointInT>(tree, "pcl::NormalEstimation");
  }
  KdTreePtr getSearchMethod() const { return tree_; }
  void setKSearch(int k) { k_ = k; }
  int getKSearch() const { return k_; }
  void setRadiusSearch(double r) { search_radius_ = r; }
  double getRadiusSearch() const { return search_radius_; }
  void setViewPoint(float x, float y, float z) { vp_[0] = x; vp_[1] = y; vp_[2] = z; use_sensor_origin_ = false; }
  void getViewPoint(float& x, float& y, float& z) { x = vp_[0]; y = vp_[1]; z = vp_[2]; }
  void useSensorOriginAsViewPoint()
  {
    use_sensor_origin_ = true;
    if (this->input_) { vp_[0] = this->input_->sensor_origin_[0]; vp_[1] = this->input_->sensor_origin_[1]; vp_[2] = this->input_->sensor_origin_[2]; }
  }
  void setNumberOfThreads(unsigned int) {}

  // Feature::compute — impl/feature.hpp:195-230
  void compute(PointCloudOut& output)
  {
    output.clear();
    if (!this->input_ || this->input_->empty()) {
      std::fprintf(stderr, "[pcl::NormalEstimation::compute] input cloud is empty!\n");
      return;
    }
    if (search_radius_ != 0.0 && k_ != 0) {  // impl/feature.hpp:135-141
      std::fprintf(stderr, "[pcl::NormalEstimation::compute] Both radius and K defined! Set one of them to zero first.\n");
      return;
    }
    if (k_ == 0 && search_radius_ == 0.0) {  // impl/feature.hpp:168-173
      std::fprintf(stderr, "[pcl::NormalEstimation::compute] Neither radius nor K defined! Set one of them to a positive number first.\n");
      return;
    }
    PCLBase<PointInT>::initCompute();
    if (!surface_) { surface_ = this->input_; fake_surface_ = true; }
    if (!tree_) tree_.reset(new pcl::search::KdTree<PointInT>());
    if (tree_->getInputCloud() != surface_ || !tree_->deviceIndex()) tree_->setInputCloud(surface_);
    if (!tree_->deviceIndex()) return;
    const std::size_t n = this->indices_->size();
    output.header = this->input_->header;
    output.points.assign(n, PointOutT());
    std::vector<float> buf(4 * (n ? n : 1));
    int dense = 1;
    int rc = search_radius_ != 0.0
                 ? pclb200_normals_radius(b200::Context::get(), tree_->deviceIndex(), this->input_->points.data(),
                                          this->input_->size(), sizeof(PointInT), this->abiIndices(), this->abiIndexCount(),
                                          this->input_->is_dense ? 1 : 0, search_radius_, vp_, buf.data(), &dense)
                 : pclb200_normals_knn(b200::Context::get(), tree_->deviceIndex(), this->input_->points.data(),
                                       this->input_->size(), sizeof(PointInT), this->abiIndices(), this->abiIndexCount(),
                                       this->input_->is_dense ? 1 : 0, k_, vp_, buf.data(), &dense);
    if (rc != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::NormalEstimation::compute] %s\n", pclb200_last_error());
      output.clear();
      return;
    }
    for (std::size_t i = 0; i < n; ++i) {
      output.points[i].normal_x = buf[4 * i];
      output.points[i].normal_y = buf[4 * i + 1];
      output.points[i].normal_z = buf[4 * i + 2];
      output.points[i].curvature = buf[4 * i + 3];
    }
    if (n != this->input_->size()) { output.width = static_cast<std::uint32_t>(n); output.height = 1; }
    else { output.width = this->input_->width; output.height = this->input_->height; }
    output.is_dense = dense != 0;
    if (fake_surface_) surface_.reset();
  }

protected:
  PointCloudInConstPtr surface_;
  bool fake_surface_ = false;
  KdTreePtr tree_;
  int k_ = 0;
  double search_radius_ = 0.0;
  float vp_[3] = {0.f, 0.f, 0.f};
  bool use_sensor_origin_ = true;
};
// features/include/pcl/features/normal_3d_omp.h:52-107: the same estimator with a thread count; on the device the batch is one
// launch, so the count is accepted and kept, nothing more
template <typename PointInT, typename PointOutT>
class NormalEstimationOMP : public NormalEstimation<PointInT, PointOutT> {
public:
  using Ptr = std::shared_ptr<NormalEstimationOMP<PointInT, PointOutT>>;
  using ConstPtr = std::shared_ptr<const NormalEstimationOMP<PointInT, PointOutT>>;
  explicit NormalEstimationOMP(unsigned int nr_threads = 0, int /*chunk_size*/ = 256) { setNumberOfThreads(nr_threads); }
  void setNumberOfThreads(unsigned int nr_threads = 0) { threads_ = nr_threads; }
  unsigned int getNumberOfThreads() const { return threads_; }

protected:
  unsigned int threads_ = 0;
};
}  // namespace pcl
