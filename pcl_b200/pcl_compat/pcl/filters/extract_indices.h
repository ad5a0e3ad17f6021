// pcl/filters/extract_indices.h — pcl::ExtractIndices<PointT> and ExtractIndices<pcl::PCLPointCloud2>
// (filters/include/pcl/filters/extract_indices.h:60-200, impl/extract_indices.hpp:49-160, filters/src/extract_indices.cpp:48-160):
// the points named by an index list, or with setNegative(true) all the others, in cloud order.  Host code — a gather of
// records; the PCL tools of the path use it to map what a filter removed back onto the blob they loaded.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "../PCLPointCloud2.h"
#include "../PointIndices.h"
#include "../point_cloud.h"

namespace pcl {
namespace detail {
// the kept positions, in ascending cloud order for the negative case and in list order otherwise (extract_indices.hpp:100-160)
inline Indices extractSelection(std::size_t n_points, const Indices* indices, bool negative)
{
  Indices out;
  if (!negative) {
    if (!indices) { out.resize(n_points); for (std::size_t i = 0; i < n_points; ++i) out[i] = static_cast<index_t>(i); return out; }
    for (index_t i : *indices)
      if (i >= 0 && static_cast<std::size_t>(i) < n_points) out.push_back(i);
    return out;
  }
  if (!indices) return out;   // everything was selected: the complement is empty
  std::vector<char> named(n_points, 0);
  for (index_t i : *indices)
    if (i >= 0 && static_cast<std::size_t>(i) < n_points) named[static_cast<std::size_t>(i)] = 1;
  for (std::size_t i = 0; i < n_points; ++i)
    if (!named[i]) out.push_back(static_cast<index_t>(i));
  return out;
}
}  // namespace detail

template <typename PointT>
class ExtractIndices : public PCLBase<PointT> {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  using PCLBase<PointT>::setIndices;
  explicit ExtractIndices(bool extract_removed_indices = false) : extract_removed_indices_(extract_removed_indices), removed_indices_(new Indices) {}
  void setIndices(const PointIndicesConstPtr& indices) { this->setIndices(IndicesPtr(new Indices(indices->indices))); }
  void setNegative(bool negative) { negative_ = negative; }
  bool getNegative() const { return negative_; }
  void setKeepOrganized(bool keep_organized) { keep_organized_ = keep_organized; }
  bool getKeepOrganized() const { return keep_organized_; }
  void setUserFilterValue(float value) { user_filter_value_ = value; }
  IndicesConstPtr getRemovedIndices() const { return removed_indices_; }
  void getRemovedIndices(PointIndices& pi) const { pi.indices = *removed_indices_; }

  void filter(Indices& indices)
  {
    indices.clear();
    removed_indices_->clear();
    if (!this->input_) return;
    const bool have_list = this->use_indices_ && this->indices_;
    indices = detail::extractSelection(this->input_->size(), have_list ? this->indices_.get() : nullptr, negative_);
    if (extract_removed_indices_ || keep_organized_) *removed_indices_ = detail::extractSelection(this->input_->size(), have_list ? this->indices_.get() : nullptr, !negative_);
  }
  void filter(PointCloud& output)
  {
    if (!this->input_) { output.clear(); return; }
    Indices kept;
    filter(kept);
    if (keep_organized_) {
      PointCloud out = *this->input_;
      for (index_t r : *removed_indices_) { PointT& p = out.points[static_cast<std::size_t>(r)]; p.x = p.y = p.z = user_filter_value_; }
      if (!removed_indices_->empty() && !std::isfinite(user_filter_value_)) out.is_dense = false;
      output = std::move(out);
      return;
    }
    PointCloud out;
    out.header = this->input_->header;
    out.is_dense = this->input_->is_dense;
    out.sensor_origin_ = this->input_->sensor_origin_;
    out.sensor_orientation_ = this->input_->sensor_orientation_;
    out.points.reserve(kept.size());
    for (index_t i : kept) out.points.push_back((*this->input_)[static_cast<std::size_t>(i)]);
    out.width = static_cast<std::uint32_t>(out.points.size());
    out.height = 1;
    output = std::move(out);
  }

protected:
  bool negative_ = false, keep_organized_ = false, extract_removed_indices_;
  float user_filter_value_ = std::numeric_limits<float>::quiet_NaN();
  IndicesPtr removed_indices_;
};

template <>
class ExtractIndices<pcl::PCLPointCloud2> {
public:
  using PCLPointCloud2 = pcl::PCLPointCloud2;
  void setInputCloud(const PCLPointCloud2::ConstPtr& cloud) { input_ = cloud; }
  PCLPointCloud2::ConstPtr const getInputCloud() const { return input_; }
  void setIndices(const IndicesPtr& indices) { indices_ = indices; }
  void setIndices(const IndicesConstPtr& indices) { indices_.reset(new Indices(*indices)); }
  void setIndices(const PointIndicesConstPtr& indices) { indices_.reset(new Indices(indices->indices)); }
  IndicesPtr getIndices() { return indices_; }
  void setNegative(bool negative) { negative_ = negative; }
  bool getNegative() const { return negative_; }
  void setKeepOrganized(bool keep_organized) { keep_organized_ = keep_organized; }
  bool getKeepOrganized() const { return keep_organized_; }
  void setUserFilterValue(float value) { user_filter_value_ = value; }

  void filter(Indices& indices)
  {
    indices.clear();
    if (!input_) return;
    indices = detail::extractSelection(static_cast<std::size_t>(input_->width) * input_->height, indices_.get(), negative_);
  }
  void filter(PCLPointCloud2& output)
  {
    if (!input_) { output = PCLPointCloud2(); return; }
    const std::size_t n = static_cast<std::size_t>(input_->width) * input_->height;
    if (keep_organized_) {  // src/extract_indices.cpp:52-97: the removed records get the user filter value in x, y, z
      PCLPointCloud2 out = *input_;
      const Indices removed = detail::extractSelection(n, indices_.get(), !negative_);
      for (const char* name : {"x", "y", "z"}) {
        int f = -1;
        for (std::size_t k = 0; k < out.fields.size(); ++k)
          if (out.fields[k].name == name) f = static_cast<int>(k);
        if (f < 0 || out.fields[static_cast<std::size_t>(f)].datatype != PCLPointField::FLOAT32) continue;
        for (index_t r : removed) std::memcpy(&out.data[static_cast<std::size_t>(r) * out.point_step + out.fields[static_cast<std::size_t>(f)].offset], &user_filter_value_, 4);
      }
      if (!removed.empty() && !std::isfinite(user_filter_value_)) out.is_dense = 0;
      output = std::move(out);
      return;
    }
    const Indices kept = detail::extractSelection(n, indices_.get(), negative_);
    PCLPointCloud2 out;
    out.header = input_->header;
    out.fields = input_->fields;
    out.is_bigendian = input_->is_bigendian;
    out.is_dense = input_->is_dense;
    out.point_step = input_->point_step;
    out.height = 1;
    out.width = static_cast<std::uint32_t>(kept.size());
    out.row_step = out.point_step * out.width;
    out.data.resize(static_cast<std::size_t>(out.row_step));
    for (std::size_t k = 0; k < kept.size(); ++k)
      std::memcpy(&out.data[k * out.point_step], &input_->data[static_cast<std::size_t>(kept[k]) * input_->point_step], out.point_step);
    output = std::move(out);
  }

protected:
  PCLPointCloud2::ConstPtr input_;
  IndicesPtr indices_;
  bool negative_ = false, keep_organized_ = false;
  float user_filter_value_ = std::numeric_limits<float>::quiet_NaN();
};
}  // namespace pcl
