// pcl/registration/correspondence_estimation_normal_shooting.h (+ _backprojection.h) — the two estimators that pick
// one of the k nearest target points with the help of normals, on the device.
// Reference: registration/include/pcl/registration/correspondence_estimation_normal_shooting.h:90-254 and
// impl/correspondence_estimation_normal_shooting.hpp:47-131; correspondence_estimation_backprojection.h:58-254 and
// impl/correspondence_estimation_backprojection.hpp:47-118.  The reciprocal variants (…hpp:133-240) are not built.
#pragma once
#include "correspondence_estimation.h"

namespace pcl {
namespace registration {

namespace detail {
template <typename PointSource, typename PointTarget, typename NormalT, typename Scalar, int Kind>
class CorrespondenceEstimationByNormals : public CorrespondenceEstimationBase<PointSource, PointTarget, Scalar> {
public:
  using Base = CorrespondenceEstimationBase<PointSource, PointTarget, Scalar>;
  using NormalsConstPtr = typename pcl::PointCloud<NormalT>::ConstPtr;
  using NormalsPtr = typename pcl::PointCloud<NormalT>::Ptr;

  void setSourceNormals(const NormalsConstPtr& normals) { source_normals_ = normals; }
  NormalsConstPtr getSourceNormals() const { return source_normals_; }
  // blob form (impl/correspondence_estimation_normal_shooting.hpp / …backprojection.hpp: fromPCLPointCloud2 + the typed setter)
  void setSourceNormals(pcl::PCLPointCloud2::ConstPtr cloud2) override
  {
    NormalsPtr cloud(new pcl::PointCloud<NormalT>);
    fromPCLPointCloud2(*cloud2, *cloud);
    setSourceNormals(NormalsConstPtr(cloud));
  }
  void setTargetNormals(pcl::PCLPointCloud2::ConstPtr cloud2) override
  {
    if (Kind != PCLB200_CORR_BACK_PROJECTION)
      return;
    NormalsPtr cloud(new pcl::PointCloud<NormalT>);
    fromPCLPointCloud2(*cloud2, *cloud);
    target_normals_ = cloud;
  }
  void setKSearch(unsigned int k) { k_ = k; }
  unsigned int getKSearch() const { return k_; }
  bool requiresSourceNormals() const override { return true; }
  int abiKind() const override { return Kind; }
  int abiK() const override { return static_cast<int>(k_); }

  void determineCorrespondences(pcl::Correspondences& out, double max_distance = std::numeric_limits<double>::max()) override
  {
    out.clear();
    if (!source_normals_ || (Kind == PCLB200_CORR_BACK_PROJECTION && !target_normals_)) {  // …hpp:51-57
      std::fprintf(stderr, "[pcl::registration::%s::initCompute] Datasets containing normals for source/target have not been given!\n", name());
      return;
    }
    if (!this->initCompute() || !this->tree_->deviceIndex()) return;
    if (source_normals_->size() != this->input_->size() ||
        (Kind == PCLB200_CORR_BACK_PROJECTION && target_normals_->size() != this->target_->size())) {
      std::fprintf(stderr, "[pcl::registration::%s] the normal clouds must have one normal per point\n", name());
      return;
    }
    out.resize(this->indices_->size());
    std::size_t n_out = 0;
    const unsigned char* sn = reinterpret_cast<const unsigned char*>(source_normals_->points.data()) + normal_offset<NormalT>::value;
    const unsigned char* tn = (Kind == PCLB200_CORR_BACK_PROJECTION)
                                  ? reinterpret_cast<const unsigned char*>(target_normals_->points.data()) + normal_offset<NormalT>::value
                                  : nullptr;
    int rc = pclb200_correspondences_normals(b200::Context::get(), this->tree_->deviceIndex(), Kind, this->input_->points.data(),
                                             this->input_->size(), sizeof(PointSource), sn, sizeof(NormalT), tn, sizeof(NormalT),
                                             this->abiIndices(), this->abiIndexCount(), static_cast<int>(k_), max_distance,
                                             reinterpret_cast<pclb200_corr*>(out.data()), &n_out);
    if (rc != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::registration::%s::determineCorrespondences] %s\n", name(), pclb200_last_error());
      n_out = 0;
    }
    out.resize(n_out);
  }
  void determineReciprocalCorrespondences(pcl::Correspondences& out, double = std::numeric_limits<double>::max()) override
  {
    std::fprintf(stderr, "[pcl::registration::%s::determineReciprocalCorrespondences] not built on the accelerated path\n", name());
    out.clear();
  }

protected:
  static const char* name()
  {
    return Kind == PCLB200_CORR_NORMAL_SHOOTING ? "CorrespondenceEstimationNormalShooting" : "CorrespondenceEstimationBackProjection";
  }
  NormalsConstPtr source_normals_, target_normals_;
  unsigned int k_ = 10;  // correspondence_estimation_normal_shooting.h:251 / correspondence_estimation_backprojection.h:251
};
}  // namespace detail

template <typename PointSource, typename PointTarget, typename NormalT, typename Scalar = float>
class CorrespondenceEstimationNormalShooting
: public detail::CorrespondenceEstimationByNormals<PointSource, PointTarget, NormalT, Scalar, PCLB200_CORR_NORMAL_SHOOTING> {
public:
  using Ptr = std::shared_ptr<CorrespondenceEstimationNormalShooting>;
  using ConstPtr = std::shared_ptr<const CorrespondenceEstimationNormalShooting>;
  typename CorrespondenceEstimationBase<PointSource, PointTarget, Scalar>::Ptr clone() const override
  {
    return typename CorrespondenceEstimationBase<PointSource, PointTarget, Scalar>::Ptr(new CorrespondenceEstimationNormalShooting(*this));
  }
};

template <typename PointSource, typename PointTarget, typename NormalT, typename Scalar = float>
class CorrespondenceEstimationBackProjection
: public detail::CorrespondenceEstimationByNormals<PointSource, PointTarget, NormalT, Scalar, PCLB200_CORR_BACK_PROJECTION> {
public:
  using Ptr = std::shared_ptr<CorrespondenceEstimationBackProjection>;
  using ConstPtr = std::shared_ptr<const CorrespondenceEstimationBackProjection>;
  using NormalsConstPtr = typename pcl::PointCloud<NormalT>::ConstPtr;
  using detail::CorrespondenceEstimationByNormals<PointSource, PointTarget, NormalT, Scalar, PCLB200_CORR_BACK_PROJECTION>::setTargetNormals;
  void setTargetNormals(const NormalsConstPtr& normals) { this->target_normals_ = normals; }
  NormalsConstPtr getTargetNormals() const { return this->target_normals_; }
  bool requiresTargetNormals() const override { return true; }
  typename CorrespondenceEstimationBase<PointSource, PointTarget, Scalar>::Ptr clone() const override
  {
    return typename CorrespondenceEstimationBase<PointSource, PointTarget, Scalar>::Ptr(new CorrespondenceEstimationBackProjection(*this));
  }
};

}  // namespace registration
}  // namespace pcl
