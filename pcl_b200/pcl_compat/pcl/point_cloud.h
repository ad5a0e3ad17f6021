// pcl/point_cloud.h — pcl::PointCloud<PointT> (common/include/pcl/point_cloud.h:393-409): header, points, width,
// height, is_dense, sensor_origin_; plus pcl::PCLBase (common/include/pcl/impl/pcl_base.hpp:138-171).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "eigen_lite.h"
#include "types.h"

namespace pcl {
struct PCLHeader {
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;
  std::string frame_id;
};

template <typename PointT>
class PointCloud {
public:
  using PointType = PointT;
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  using iterator = typename std::vector<PointT>::iterator;
  using const_iterator = typename std::vector<PointT>::const_iterator;

  PCLHeader header;
  std::vector<PointT> points;
  std::uint32_t width = 0;
  std::uint32_t height = 0;
  bool is_dense = true;
  Eigen::Vector4f sensor_origin_;

  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  void resize(std::size_t n)
  {
    points.resize(n);
    if (width * height != n) { width = static_cast<std::uint32_t>(n); height = 1; }
  }
  void push_back(const PointT& p) { points.push_back(p); width = static_cast<std::uint32_t>(points.size()); height = 1; }
  template <typename... A> void emplace_back(A&&... a) { points.emplace_back(std::forward<A>(a)...); width = static_cast<std::uint32_t>(points.size()); height = 1; }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  PointT* data() { return points.data(); }
  const PointT* data() const { return points.data(); }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};

template <typename PointT>
class PCLBase {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  using PointCloudPtr = typename PointCloud::Ptr;
  using PointCloudConstPtr = typename PointCloud::ConstPtr;
  virtual ~PCLBase() = default;
  virtual void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  PointCloudConstPtr const getInputCloud() const { return input_; }
  virtual void setIndices(const IndicesPtr& indices) { indices_ = indices; use_indices_ = true; fake_indices_ = false; }
  virtual void setIndices(const IndicesConstPtr& indices) { indices_.reset(new Indices(*indices)); use_indices_ = true; fake_indices_ = false; }
  IndicesPtr getIndices() { return indices_; }

protected:
  PointCloudConstPtr input_;
  IndicesPtr indices_;
  bool use_indices_ = false;
  bool fake_indices_ = false;
  // pcl_base.hpp:138-171: identity indices when none were given
  bool initCompute()
  {
    if (!input_) return false;
    if (!indices_) { fake_indices_ = true; indices_.reset(new Indices); }
    if (fake_indices_ && indices_->size() != input_->size()) {
      indices_->resize(input_->size());
      for (std::size_t i = 0; i < indices_->size(); ++i) (*indices_)[i] = static_cast<index_t>(i);
    }
    return true;
  }
  bool deinitCompute() { return true; }
  // indices to hand to the C-ABI: NULL when they are the identity
  const index_t* abiIndices() const { return (fake_indices_ || !indices_) ? nullptr : indices_->data(); }
  std::size_t abiIndexCount() const { return (fake_indices_ || !indices_) ? 0 : indices_->size(); }
};
}  // namespace pcl
