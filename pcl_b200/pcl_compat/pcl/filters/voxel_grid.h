// pcl/filters/voxel_grid.h — pcl::VoxelGrid<PointT> (filters/include/pcl/filters/voxel_grid.h:220-530,
// impl/voxel_grid.hpp:596-814) on the device.  Supported: setLeafSize, setMinimumPointsNumberPerVoxel, indices,
// filter().  downsample_all_data_ (default true, voxel_grid.hpp:796-806): for point types that carry a normal and a
// curvature (PointNormal) those fields are averaged on the device too, like CentroidPoint does (normal = normalised
// 4-vector sum, curvature = mean); with setDownsampleAllData(false) only xyz is produced and the other fields are
// default-initialised (voxel_grid.hpp:784-794 copies just the 4-float centroid).
#pragma once
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../b200/context.h"
#include "../point_cloud.h"

namespace pcl {
template <typename PointT>
class VoxelGrid : public PCLBase<PointT> {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void setMinimumPointsNumberPerVoxel(unsigned int n) { min_points_per_voxel_ = n; }
  unsigned int getMinimumPointsNumberPerVoxel() const { return min_points_per_voxel_; }
  void setDownsampleAllData(bool v) { downsample_all_data_ = v; }
  void setLeafSize(const Eigen::Vector4f& l) { leaf_[0] = l[0]; leaf_[1] = l[1]; leaf_[2] = l[2]; }
  Eigen::Vector3f getLeafSize() const
  {
    Eigen::Vector3f l;
    l[0] = leaf_[0]; l[1] = leaf_[1]; l[2] = leaf_[2];
    return l;
  }
  // voxel_grid.h:393-470: pass-through limits on one coordinate field before the grid.  Only "x", "y", "z" exist on
  // the point types of this path; min/max (and therefore the grid origin) are taken over the selected points, exactly
  // like getMinMax3D with a filter field (voxel_grid.hpp:614-617, 663-690).
  void setFilterFieldName(const std::string& name) { filter_field_name_ = name; }
  const std::string& getFilterFieldName() const { return filter_field_name_; }
  void setFilterLimits(const double& lo, const double& hi) { filter_limit_min_ = lo; filter_limit_max_ = hi; }
  void getFilterLimits(double& lo, double& hi) const { lo = filter_limit_min_; hi = filter_limit_max_; }
  void setFilterLimitsNegative(bool v) { filter_limit_negative_ = v; }
  void filter(PointCloud& output)
  {
    if (!this->input_) {
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] No input dataset given!\n");
      output.clear();
      return;
    }
    PCLBase<PointT>::initCompute();
    output.header = this->input_->header;
    Indices selected;
    const index_t* abi_idx = this->abiIndices();
    std::size_t abi_cnt = this->abiIndexCount();
    if (!filter_field_name_.empty()) {
      const int f = filter_field_name_ == "x" ? 0 : filter_field_name_ == "y" ? 1 : filter_field_name_ == "z" ? 2 : -1;
      if (f < 0) {
        std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] Invalid filter field name (%s).\n", filter_field_name_.c_str());
        output.clear();
        return;
      }
      for (index_t i : *this->indices_) {
        const PointT& p = (*this->input_)[i];
        if (!this->input_->is_dense && !isXYZFinite(p)) continue;
        const float v = p.data[f];
        if (filter_limit_negative_) { if (v < filter_limit_max_ && v > filter_limit_min_) continue; }
        else if (v > filter_limit_max_ || v < filter_limit_min_) continue;
        selected.push_back(i);
      }
      if (selected.empty()) { output.clear(); return; }
      abi_idx = selected.data();
      abi_cnt = selected.size();
    }
    const std::size_t n = abi_idx ? abi_cnt : this->indices_->size();
    std::vector<float> xyz1(4 * (n ? n : 1));
    std::vector<float> ncurv;
    std::size_t m = 0;
    int rc;
    constexpr bool has_normal = pcl::has_normal<PointT>::value && sizeof(PointT) == 48;  // pcl::PointNormal: {xyz1 | normal4 | curvature, pad}
    if (has_normal && downsample_all_data_) {
      ncurv.resize(8 * (n ? n : 1));
      const unsigned char* base = reinterpret_cast<const unsigned char*>(this->input_->points.data());
      rc = pclb200_voxelgrid_normals(b200::Context::get(), base, this->input_->size(), sizeof(PointT), base + 16,
                                     sizeof(PointT), abi_idx, abi_cnt, this->input_->is_dense ? 1 : 0, leaf_,
                                     min_points_per_voxel_, xyz1.data(), ncurv.data(), &m);
    }
    else
      rc = pclb200_voxelgrid(b200::Context::get(), this->input_->points.data(), this->input_->size(), sizeof(PointT),
                             abi_idx, abi_cnt, this->input_->is_dense ? 1 : 0, leaf_, min_points_per_voxel_,
                             xyz1.data(), &m);
    if (rc == PCLB200_ERR_LEAF_TOO_SMALL) {  // voxel_grid.hpp:620-629: warn and return the input unfiltered
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] Leaf size is too small for the input dataset. Integer indices would overflow.\n");
      output = *this->input_;
      return;
    }
    if (rc != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] %s\n", pclb200_last_error());
      output.clear();
      return;
    }
    output.points.assign(m, PointT());
    for (std::size_t i = 0; i < m; ++i) {
      output.points[i].x = xyz1[4 * i];
      output.points[i].y = xyz1[4 * i + 1];
      output.points[i].z = xyz1[4 * i + 2];
      if (!ncurv.empty())  // bytes 16..47 of a PointNormal: normal4 | curvature
        std::memcpy(reinterpret_cast<unsigned char*>(&output.points[i]) + 16, &ncurv[8 * i], 32);
    }
    output.width = static_cast<std::uint32_t>(m);
    output.height = 1;       // downsampling breaks the organized structure (:609)
    output.is_dense = true;  // (:610)
  }

protected:
  float leaf_[3] = {0.f, 0.f, 0.f};
  unsigned int min_points_per_voxel_ = 0;
  bool downsample_all_data_ = true;
  std::string filter_field_name_;
  double filter_limit_min_ = std::numeric_limits<float>::lowest(), filter_limit_max_ = std::numeric_limits<float>::max();
  bool filter_limit_negative_ = false;
};
}  // namespace pcl
