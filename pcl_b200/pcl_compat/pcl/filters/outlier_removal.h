// pcl/filters/statistical_outlier_removal.h + radius_outlier_removal.h — the two outlier filters that sit in front of
// ICP in most PCL pipelines, as batch calls on the same device index.
// Reference: filters/include/pcl/filters/impl/statistical_outlier_removal.hpp:47-135 and
// radius_outlier_removal.hpp:49-179.  The k-NN part (the whole cost) runs on the GPU (pclb200_knn_stats, 4 bytes per
// point come back); the global mean / variance and the classification are the reference's own sequential double
// arithmetic, evaluated on the host so the threshold is bit-identical.
#pragma once
#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

#include "../point_cloud.h"
#include "../point_types.h"
#include "../PointIndices.h"
#include "../search/kdtree.h"
#include "filter.h"

namespace pcl {

template <typename PointT>
class OutlierFilterBase : public PCLBase<PointT> {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  explicit OutlierFilterBase(bool extract_removed_indices = false)
  : removed_indices_(new Indices), extract_removed_indices_(extract_removed_indices) {}
  void setNegative(bool negative) { negative_ = negative; }
  bool getNegative() const { return negative_; }
  // filter_indices.h:120-170: the output keeps the input's size and structure, removed points get the user filter value (NaN)
  void setKeepOrganized(bool keep_organized) { keep_organized_ = keep_organized; }
  bool getKeepOrganized() const { return keep_organized_; }
  void setUserFilterValue(float value) { user_filter_value_ = value; }
  void setNumberOfThreads(unsigned int) {}
  IndicesConstPtr getRemovedIndices() const { return removed_indices_; }
  void getRemovedIndices(PointIndices& pi) const { pi.indices = *removed_indices_; }   // filter.h:100-104
  void setSearchMethod(const typename pcl::search::KdTree<PointT>::Ptr& tree) { searcher_ = tree; }
  // statistical_outlier_removal.h:88,150 / radius_outlier_removal.h:79,146: SearcherPtr = pcl::search::Search<PointT>::Ptr
  void setSearchMethod(const typename pcl::search::Search<PointT>::Ptr& searcher)
  {
    searcher_ = pcl::search::deviceSearcher<PointT>(searcher, "pcl::OutlierRemoval");
  }

  void filter(Indices& indices)
  {
    indices.clear();
    removed_indices_->clear();
    if (!this->input_ || this->input_->empty()) return;
    PCLBase<PointT>::initCompute();
    if (!searcher_) searcher_.reset(new pcl::search::KdTree<PointT>());
    // like the reference (statistical_outlier_removal.hpp:63, radius_outlier_removal.hpp:66): the searcher is handed the
    // input on EVERY call — a cloud mutated in place, or a tree that was rebuilt elsewhere, must not meet a stale index
    if (!searcher_->setInputCloud(this->input_)) {
      std::fprintf(stderr, "[pcl::%s::applyFilter] Error when initializing search method!\n", name_);
      return;
    }
    std::vector<std::uint8_t> keep;
    if (!classify(keep)) return;
    for (std::size_t i = 0; i < keep.size(); ++i) {
      if (keep[i]) indices.push_back((*this->indices_)[i]);
      else if (extract_removed_indices_) removed_indices_->push_back((*this->indices_)[i]);
    }
  }
  void filter(PointCloud& output)
  {
    if (keep_organized_ && this->input_) {  // impl/filter_indices.hpp:47-63
      const bool temp = extract_removed_indices_;
      extract_removed_indices_ = true;
      Indices ind;
      filter(ind);
      extract_removed_indices_ = temp;
      output = *this->input_;
      for (index_t rii : *removed_indices_) {
        PointT& p = output.points[static_cast<std::size_t>(rii)];
        p.x = p.y = p.z = user_filter_value_;
      }
      if (!std::isfinite(user_filter_value_)) output.is_dense = false;
      return;
    }
    Indices ind;
    filter(ind);
    output.header = this->input_ ? this->input_->header : PCLHeader();
    output.points.clear();
    if (this->input_)
      for (index_t i : ind) output.points.push_back((*this->input_)[i]);
    output.width = static_cast<std::uint32_t>(output.points.size());
    output.height = 1;
    output.is_dense = true;  // both filters drop non-finite points
  }

protected:
  virtual bool classify(std::vector<std::uint8_t>& keep) = 0;
  bool stats(int k, float* mean, float* kth)
  {
    int rc = pclb200_knn_stats(b200::Context::get(), searcher_->deviceIndex(), this->input_->points.data(), this->input_->size(),
                               sizeof(PointT), this->abiIndices(), this->abiIndexCount(), k, mean, kth);
    if (rc != PCLB200_OK) std::fprintf(stderr, "[pcl::%s::applyFilter] %s\n", name_, pclb200_last_error());
    return rc == PCLB200_OK;
  }
  typename pcl::search::KdTree<PointT>::Ptr searcher_;
  IndicesPtr removed_indices_;
  bool extract_removed_indices_;
  bool negative_ = false;
  bool keep_organized_ = false;
  float user_filter_value_ = std::numeric_limits<float>::quiet_NaN();
  const char* name_ = "Filter";
};

template <typename PointT>
class StatisticalOutlierRemoval : public OutlierFilterBase<PointT> {
public:
  explicit StatisticalOutlierRemoval(bool extract_removed_indices = false) : OutlierFilterBase<PointT>(extract_removed_indices)
  {
    this->name_ = "StatisticalOutlierRemoval";
  }
  void setMeanK(int k) { mean_k_ = k; }
  int getMeanK() const { return mean_k_; }
  void setStddevMulThresh(double m) { std_mul_ = m; }
  double getStddevMulThresh() const { return std_mul_; }

protected:
  bool classify(std::vector<std::uint8_t>& keep) override
  {
    const std::size_t n = this->indices_->size();
    std::vector<float> distances(n);
    if (!this->stats(mean_k_ + 1, distances.data(), nullptr)) return false;
    long long valid = 0;
    for (std::size_t i = 0; i < n; ++i)
      if (isXYZFinite((*this->input_)[(*this->indices_)[i]])) ++valid;
    // statistical_outlier_removal.hpp:100-112
    double sum = 0, sq_sum = 0;
    for (const float& d : distances) {
      sum += d;
      sq_sum += d * d;
    }
    const double mean = sum / static_cast<double>(valid);
    const double variance = (sq_sum - sum * sum / static_cast<double>(valid)) / (static_cast<double>(valid) - 1);
    const double thr = mean + std_mul_ * std::sqrt(variance);
    keep.assign(n, 1);
    for (std::size_t i = 0; i < n; ++i)
      if ((!this->negative_ && distances[i] > thr) || (this->negative_ && distances[i] <= thr)) keep[i] = 0;
    return true;
  }
  int mean_k_ = 1;
  double std_mul_ = 0.0;
};

template <typename PointT>
class RadiusOutlierRemoval : public OutlierFilterBase<PointT> {
public:
  explicit RadiusOutlierRemoval(bool extract_removed_indices = false) : OutlierFilterBase<PointT>(extract_removed_indices)
  {
    this->name_ = "RadiusOutlierRemoval";
  }
  void setRadiusSearch(double r) { search_radius_ = r; }
  double getRadiusSearch() const { return search_radius_; }
  void setMinNeighborsInRadius(int n) { min_pts_radius_ = n; }
  int getMinNeighborsInRadius() const { return min_pts_radius_; }

protected:
  bool classify(std::vector<std::uint8_t>& keep) override
  {
    if (search_radius_ == 0.0) {
      std::fprintf(stderr, "[pcl::RadiusOutlierRemoval::applyFilter] No radius defined!\n");
      return false;
    }
    const std::size_t n = this->indices_->size();
    std::vector<float> kth(n);
    if (!this->stats(min_pts_radius_ + 1, nullptr, kth.data())) return false;
    const double nn_dists_max = search_radius_ * search_radius_;
    const float r2f = static_cast<float>(search_radius_ * search_radius_);
    keep.assign(n, 1);
    for (std::size_t i = 0; i < n; ++i) {
      if (this->input_->is_dense) {  // radius_outlier_removal.hpp:86-118: k-NN rule
        if (std::isfinite(kth[i])) {
          if ((!this->negative_ && nn_dists_max < kth[i]) || (this->negative_ && nn_dists_max >= kth[i])) keep[i] = 0;
        }
        else if (!this->negative_) keep[i] = 0;
      }
      else {  // :121-150: radius rule (FLANN's strict d2 < r2), non-finite points removed
        if (!isXYZFinite((*this->input_)[(*this->indices_)[i]])) { keep[i] = 0; continue; }
        const bool enough = kth[i] < r2f;
        if ((!this->negative_ && !enough) || (this->negative_ && enough)) keep[i] = 0;
      }
    }
    return true;
  }
  double search_radius_ = 0.0;
  int min_pts_radius_ = 1;
};

}  // namespace pcl
