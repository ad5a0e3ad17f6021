// pcl/filters/voxel_grid.h — pcl::VoxelGrid<PointT> (filters/include/pcl/filters/voxel_grid.h:220-530,
// impl/voxel_grid.hpp:596-814) on the device.  Supported: setLeafSize, setMinimumPointsNumberPerVoxel, indices,
// filter().  downsample_all_data only matters for fields beyond xyz: the centroid of xyz is produced, other fields
// of the output are default-initialised.
#pragma once
#include <cstdio>

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
  void filter(PointCloud& output)
  {
    if (!this->input_) {
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] No input dataset given!\n");
      output.clear();
      return;
    }
    PCLBase<PointT>::initCompute();
    output.header = this->input_->header;
    const std::size_t n = this->indices_->size();
    std::vector<float> xyz1(4 * (n ? n : 1));
    std::size_t m = 0;
    int rc = pclb200_voxelgrid(b200::Context::get(), this->input_->points.data(), this->input_->size(), sizeof(PointT),
                               this->abiIndices(), this->abiIndexCount(), this->input_->is_dense ? 1 : 0, leaf_, min_points_per_voxel_,
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
    }
    output.width = static_cast<std::uint32_t>(m);
    output.height = 1;       // downsampling breaks the organized structure (:609)
    output.is_dense = true;  // (:610)
  }

protected:
  float leaf_[3] = {0.f, 0.f, 0.f};
  unsigned int min_points_per_voxel_ = 0;
  bool downsample_all_data_ = true;
};
}  // namespace pcl
