// pcl/registration/transformation_validation_euclidean.h — pcl::registration::TransformationValidationEuclidean
// (registration/include/pcl/registration/transformation_validation_euclidean.h:77-263,
// impl/transformation_validation_euclidean.hpp:50-109) on the device: transform the source by the candidate pose, one
// batch 1-NN on the target's index, mean squared distance of the pairs inside max_range.  The transform keeps the
// reference's Scalar arithmetic (products and sums left to right, cast to float at the end, :62-75).
#pragma once
#include <cmath>
#include <cstdio>
#include <limits>
#include <memory>

#include "../b200/context.h"
#include "../eigen_lite.h"
#include "../point_cloud.h"
#include "../search/kdtree.h"

namespace pcl {
namespace registration {

template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationValidationEuclidean {
public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using Ptr = std::shared_ptr<TransformationValidationEuclidean>;
  using ConstPtr = std::shared_ptr<const TransformationValidationEuclidean>;
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  using PointCloudSourceConstPtr = typename pcl::PointCloud<PointSource>::ConstPtr;
  using PointCloudTargetConstPtr = typename pcl::PointCloud<PointTarget>::ConstPtr;

  TransformationValidationEuclidean()
  : max_range_(std::numeric_limits<double>::max()), threshold_(std::numeric_limits<double>::quiet_NaN()), tree_(new KdTree) {}
  virtual ~TransformationValidationEuclidean() = default;

  // transformation_validation_euclidean.h:112-122: a tree the caller has already filled is used as is
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false)
  {
    tree_ = tree;
    force_no_recompute_ = force_no_recompute;
  }
  void setMaxRange(double max_range) { max_range_ = max_range; }  // compared with SQUARED distances, as the reference does
  double getMaxRange() { return max_range_; }
  void setThreshold(double threshold) { threshold_ = threshold; }
  double getThreshold() { return threshold_; }

  double validateTransformation(const PointCloudSourceConstPtr& cloud_src, const PointCloudTargetConstPtr& cloud_tgt,
                                const Matrix4& transformation_matrix) const
  {
    if (!cloud_src || cloud_src->empty())
      return std::numeric_limits<double>::max();
    if (!force_no_recompute_ || !tree_->deviceIndex()) {
      if (!tree_->setInputCloud(cloud_tgt))
        return std::numeric_limits<double>::max();
    }
    double T[16];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) T[4 * r + c] = static_cast<double>(transformation_matrix(r, c));
    double score = std::numeric_limits<double>::max();
    int rc = pclb200_validate_transformation(b200::Context::get(), tree_->deviceIndex(), cloud_src->points.data(),
                                             cloud_src->size(), sizeof(PointSource), T, sizeof(Scalar) == 8 ? 1 : 0,
                                             max_range_, &score);
    if (rc != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::TransformationValidationEuclidean::validateTransformation] %s\n", pclb200_last_error());
      return std::numeric_limits<double>::max();
    }
    return score;
  }

  virtual bool operator()(const double& score1, const double& score2) const { return score1 < score2; }

  virtual bool isValid(const PointCloudSourceConstPtr& cloud_src, const PointCloudTargetConstPtr& cloud_tgt,
                       const Matrix4& transformation_matrix) const
  {
    if (std::isnan(threshold_)) {
      std::fprintf(stderr, "[pcl::TransformationValidationEuclidean::isValid] Threshold not set! Please use setThreshold () "
                           "before continuing.\n");
      return false;
    }
    return validateTransformation(cloud_src, cloud_tgt, transformation_matrix) < threshold_;
  }

protected:
  double max_range_;
  double threshold_;
  KdTreePtr tree_;
  bool force_no_recompute_ = false;
};

}  // namespace registration
}  // namespace pcl
