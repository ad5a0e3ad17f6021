// pcl/point_cloud.h — pcl::PointCloud<PointT> (common/include/pcl/point_cloud.h:393-409): header, points, width,
// height, is_dense, sensor_origin_; plus pcl::PCLBase (common/include/pcl/impl/pcl_base.hpp:138-171).
#pragma once
#include <algorithm>
#include <cstdint>
#include <initializer_list>
#include <iterator>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "console/print.h"
#include "eigen_lite.h"
#include "memory.h"
#include "types.h"

namespace pcl {
struct PCLHeader {
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;
  std::string frame_id;
};

// PCLHeader.h:29-35
inline std::ostream& operator<<(std::ostream& out, const PCLHeader& h)
{
  out << "seq: " << h.seq;
  out << " stamp: " << h.stamp;
  out << " frame_id: " << h.frame_id << std::endl;
  return out;
}

// 2-D indexing of an unorganised cloud (common/include/pcl/exceptions.h: UnorganizedPointCloudException)
struct UnorganizedPointCloudException : std::runtime_error {
  explicit UnorganizedPointCloudException(const std::string& what) : std::runtime_error(what) {}
};

// The container keeps the reference's bookkeeping (point_cloud.h:173-760): every size-changing member that is not
// "transient_" leaves an unorganised cloud behind (width = size, height = 1); the (width, height) forms keep the grid.
template <typename PointT>
class PointCloud {
public:
  using PointType = PointT;
  using VectorType = std::vector<PointT>;
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  using value_type = PointT;
  using reference = PointT&;
  using const_reference = const PointT&;
  using difference_type = typename VectorType::difference_type;
  using size_type = typename VectorType::size_type;
  using iterator = typename VectorType::iterator;
  using const_iterator = typename VectorType::const_iterator;
  using reverse_iterator = typename VectorType::reverse_iterator;
  using const_reverse_iterator = typename VectorType::const_reverse_iterator;

  PCLHeader header;
  VectorType points;
  std::uint32_t width = 0;
  std::uint32_t height = 0;
  bool is_dense = true;
  Eigen::Vector4f sensor_origin_;
  Eigen::Quaternionf sensor_orientation_;

  PointCloud() = default;
  // subset copy (point_cloud.h:186-197)
  PointCloud(const PointCloud<PointT>& pc, const Indices& indices)
  : header(pc.header), points(indices.size()), width(static_cast<std::uint32_t>(indices.size())), height(1),
    is_dense(pc.is_dense), sensor_origin_(pc.sensor_origin_), sensor_orientation_(pc.sensor_orientation_)
  {
    for (std::size_t i = 0; i < indices.size(); ++i) points[i] = pc[static_cast<std::size_t>(indices[i])];
  }
  PointCloud(std::uint32_t width_, std::uint32_t height_, const PointT& value_ = PointT())
  : points(static_cast<std::size_t>(width_) * height_, value_), width(width_), height(height_) {}

  // concatenation (point_cloud.h:209-254): newest stamp, unorganised result, dense only if both were
  PointCloud& operator+=(const PointCloud& rhs)
  {
    concatenate(*this, rhs);
    return *this;
  }
  PointCloud operator+(const PointCloud& rhs) const { return PointCloud(*this) += rhs; }
  static bool concatenate(PointCloud<PointT>& cloud1, const PointCloud<PointT>& cloud2)
  {
    cloud1.header.stamp = std::max(cloud1.header.stamp, cloud2.header.stamp);
    cloud1.points.insert(cloud1.points.end(), cloud2.points.begin(), cloud2.points.end());
    cloud1.width = static_cast<std::uint32_t>(cloud1.size());
    cloud1.height = 1;
    cloud1.is_dense = cloud1.is_dense && cloud2.is_dense;
    return true;
  }
  static bool concatenate(const PointCloud<PointT>& cloud1, const PointCloud<PointT>& cloud2, PointCloud<PointT>& cloud_out)
  {
    cloud_out = cloud1;
    return concatenate(cloud_out, cloud2);
  }

  // organised access (point_cloud.h:261-312)
  const PointT& at(int column, int row) const
  {
    if (height <= 1) throw UnorganizedPointCloudException("Can't use 2D indexing with an unorganized point cloud");
    return points.at(static_cast<std::size_t>(row) * width + static_cast<std::size_t>(column));
  }
  PointT& at(int column, int row)
  {
    if (height <= 1) throw UnorganizedPointCloudException("Can't use 2D indexing with an unorganized point cloud");
    return points.at(static_cast<std::size_t>(row) * width + static_cast<std::size_t>(column));
  }
  const PointT& operator()(std::size_t column, std::size_t row) const { return points[row * width + column]; }
  PointT& operator()(std::size_t column, std::size_t row) { return points[row * width + column]; }
  bool isOrganized() const { return height > 1; }

  // iterators
  iterator begin() noexcept { return points.begin(); }
  iterator end() noexcept { return points.end(); }
  const_iterator begin() const noexcept { return points.begin(); }
  const_iterator end() const noexcept { return points.end(); }
  const_iterator cbegin() const noexcept { return points.cbegin(); }
  const_iterator cend() const noexcept { return points.cend(); }
  reverse_iterator rbegin() noexcept { return points.rbegin(); }
  reverse_iterator rend() noexcept { return points.rend(); }
  const_reverse_iterator rbegin() const noexcept { return points.rbegin(); }
  const_reverse_iterator rend() const noexcept { return points.rend(); }
  const_reverse_iterator crbegin() const noexcept { return points.crbegin(); }
  const_reverse_iterator crend() const noexcept { return points.crend(); }

  // capacity
  std::size_t size() const { return points.size(); }
  index_t max_size() const noexcept { return static_cast<index_t>(std::min<std::size_t>(points.max_size(), 0x7fffffff)); }
  void reserve(std::size_t n) { points.reserve(n); }
  bool empty() const { return points.empty(); }
  PointT* data() noexcept { return points.data(); }
  const PointT* data() const noexcept { return points.data(); }
  void resize(std::size_t count)
  {
    points.resize(count);
    if (static_cast<std::size_t>(width) * height != count) unorganised();
  }
  void resize(uindex_t new_width, uindex_t new_height)
  {
    points.resize(static_cast<std::size_t>(new_width) * new_height);
    width = new_width;
    height = new_height;
  }
  void resize(index_t count, const PointT& value)
  {
    points.resize(static_cast<std::size_t>(count), value);
    if (static_cast<std::size_t>(width) * height != static_cast<std::size_t>(count)) unorganised();
  }
  void resize(index_t new_width, index_t new_height, const PointT& value)
  {
    points.resize(static_cast<std::size_t>(new_width) * static_cast<std::size_t>(new_height), value);
    width = static_cast<std::uint32_t>(new_width);
    height = static_cast<std::uint32_t>(new_height);
  }

  // element access
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  PointT& front() { return points.front(); }
  const PointT& front() const { return points.front(); }
  PointT& back() { return points.back(); }
  const PointT& back() const { return points.back(); }

  // assign (point_cloud.h:545-640): a width that does not divide the size falls back to one row
  void assign(index_t count, const PointT& value)
  {
    points.assign(static_cast<std::size_t>(count), value);
    unorganised();
  }
  void assign(index_t new_width, index_t new_height, const PointT& value)
  {
    points.assign(static_cast<std::size_t>(new_width) * static_cast<std::size_t>(new_height), value);
    width = static_cast<std::uint32_t>(new_width);
    height = static_cast<std::uint32_t>(new_height);
  }
  template <class InputIterator, typename = typename std::iterator_traits<InputIterator>::iterator_category>
  void assign(InputIterator first, InputIterator last)
  {
    points.assign(first, last);
    unorganised();
  }
  template <class InputIterator, typename = typename std::iterator_traits<InputIterator>::iterator_category>
  void assign(InputIterator first, InputIterator last, index_t new_width)
  {
    points.assign(first, last);
    organise(new_width);
  }
  void assign(std::initializer_list<PointT> ilist)
  {
    points.assign(ilist);
    unorganised();
  }
  void assign(std::initializer_list<PointT> ilist, index_t new_width)
  {
    points.assign(ilist);
    organise(new_width);
  }

  // growth: the plain members reset the grid, the transient_ ones leave width / height to the caller
  void push_back(const PointT& p)
  {
    points.push_back(p);
    unorganised();
  }
  void transient_push_back(const PointT& p) { points.push_back(p); }
  template <typename... A>
  reference emplace_back(A&&... a)
  {
    points.emplace_back(std::forward<A>(a)...);
    unorganised();
    return points.back();
  }
  template <typename... A>
  reference transient_emplace_back(A&&... a)
  {
    points.emplace_back(std::forward<A>(a)...);
    return points.back();
  }
  iterator insert(iterator position, const PointT& p)
  {
    iterator it = points.insert(position, p);
    unorganised();
    return it;
  }
  iterator transient_insert(iterator position, const PointT& p) { return points.insert(position, p); }
  void insert(iterator position, std::size_t n, const PointT& p)
  {
    points.insert(position, n, p);
    unorganised();
  }
  void transient_insert(iterator position, std::size_t n, const PointT& p) { points.insert(position, n, p); }
  template <class InputIterator, typename = typename std::iterator_traits<InputIterator>::iterator_category>
  void insert(iterator position, InputIterator first, InputIterator last)
  {
    points.insert(position, first, last);
    unorganised();
  }
  template <class InputIterator, typename = typename std::iterator_traits<InputIterator>::iterator_category>
  void transient_insert(iterator position, InputIterator first, InputIterator last)
  {
    points.insert(position, first, last);
  }
  template <typename... A>
  iterator emplace(iterator position, A&&... a)
  {
    iterator it = points.emplace(position, std::forward<A>(a)...);
    unorganised();
    return it;
  }
  template <typename... A>
  iterator transient_emplace(iterator position, A&&... a)
  {
    return points.emplace(position, std::forward<A>(a)...);
  }
  iterator erase(iterator position)
  {
    iterator it = points.erase(position);
    unorganised();
    return it;
  }
  iterator transient_erase(iterator position) { return points.erase(position); }
  iterator erase(iterator first, iterator last)
  {
    iterator it = points.erase(first, last);
    unorganised();
    return it;
  }
  iterator transient_erase(iterator first, iterator last) { return points.erase(first, last); }
  void swap(PointCloud<PointT>& rhs)
  {
    std::swap(header, rhs.header);
    points.swap(rhs.points);
    std::swap(width, rhs.width);
    std::swap(height, rhs.height);
    std::swap(is_dense, rhs.is_dense);
    std::swap(sensor_origin_, rhs.sensor_origin_);
    std::swap(sensor_orientation_, rhs.sensor_orientation_);
  }
  void clear()
  {
    points.clear();
    width = 0;
    height = 0;
  }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }

private:
  void unorganised()
  {
    width = static_cast<std::uint32_t>(points.size());
    height = 1;
  }
  void organise(index_t new_width)
  {
    if (new_width <= 0) {
      unorganised();
      return;
    }
    width = static_cast<std::uint32_t>(new_width);
    height = static_cast<std::uint32_t>(points.size() / width);
    if (static_cast<std::size_t>(width) * height != points.size()) unorganised();
  }
};

// point_cloud.h:904-923
template <typename PointT>
std::ostream& operator<<(std::ostream& s, const pcl::PointCloud<PointT>& p)
{
  s << "header: " << p.header << std::endl;
  s << "points[]: " << p.size() << std::endl;
  s << "width: " << p.width << std::endl;
  s << "height: " << p.height << std::endl;
  s << "is_dense: " << p.is_dense << std::endl;
  s << "sensor origin (xyz): [" << p.sensor_origin_[0] << ", " << p.sensor_origin_[1] << ", " << p.sensor_origin_[2] << "] / orientation (xyzw): ["
    << p.sensor_orientation_.x() << ", " << p.sensor_orientation_.y() << ", " << p.sensor_orientation_.z() << ", " << p.sensor_orientation_.w() << "]"
    << std::endl;
  return s;
}

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
