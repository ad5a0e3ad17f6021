// pcl/point_representation.h — pcl::PointRepresentation / pcl::DefaultPointRepresentation
// (common/include/pcl/point_representation.h:59-250): how a point becomes the float vector the searcher indexes.
// The device index is three-dimensional, so representations of up to three dimensions are honoured (their vectors are
// built on the host with the representation's own copyToFloatArray + rescale values, exactly like
// KdTreeFLANN::convertCloudToArray, kdtree_flann.hpp:429-498); a representation with more dimensions is refused.
#pragma once
#include <cmath>
#include <memory>
#include <vector>

#include "point_types.h"

namespace pcl {

template <typename PointT>
class PointRepresentation {
public:
  using Ptr = std::shared_ptr<PointRepresentation<PointT>>;
  using ConstPtr = std::shared_ptr<const PointRepresentation<PointT>>;
  virtual ~PointRepresentation() = default;

  virtual void copyToFloatArray(const PointT& p, float* out) const = 0;

  // point_representation.h:104-140
  inline bool isTrivial() const { return trivial_ && alpha_.empty(); }
  virtual bool isValid(const PointT& p) const
  {
    std::vector<float> tmp(static_cast<std::size_t>(nr_dimensions_));
    copyToFloatArray(p, tmp.data());
    for (int i = 0; i < nr_dimensions_; ++i)
      if (!std::isfinite(tmp[static_cast<std::size_t>(i)])) return false;
    return true;
  }
  // point_representation.h:150-183: copy, then rescale when alpha_ is set
  template <typename OutputType>
  void vectorize(const PointT& p, OutputType& out) const
  {
    std::vector<float> tmp(static_cast<std::size_t>(nr_dimensions_));
    copyToFloatArray(p, tmp.data());
    for (int i = 0; i < nr_dimensions_; ++i)
      out[i] = alpha_.empty() ? tmp[static_cast<std::size_t>(i)] : tmp[static_cast<std::size_t>(i)] * alpha_[static_cast<std::size_t>(i)];
  }
  void vectorize(const PointT& p, float* out) const
  {
    copyToFloatArray(p, out);
    if (!alpha_.empty())
      for (int i = 0; i < nr_dimensions_; ++i) out[i] *= alpha_[static_cast<std::size_t>(i)];
  }
  // point_representation.h:185-192
  void setRescaleValues(const float* rescale_array)
  {
    alpha_.assign(rescale_array, rescale_array + nr_dimensions_);
  }
  inline int getNumberOfDimensions() const { return nr_dimensions_; }

protected:
  int nr_dimensions_ = 0;
  std::vector<float> alpha_;
  bool trivial_ = false;
};

// point_representation.h:197-235 (and the PointXYZ / PointNormal specialisations :253-330 for the xyz part)
template <typename PointDefault>
class DefaultPointRepresentation : public PointRepresentation<PointDefault> {
public:
  using Ptr = std::shared_ptr<DefaultPointRepresentation<PointDefault>>;
  using ConstPtr = std::shared_ptr<const DefaultPointRepresentation<PointDefault>>;
  DefaultPointRepresentation()
  {
    this->nr_dimensions_ = 3;
    this->trivial_ = true;
  }
  inline Ptr makeShared() const { return Ptr(new DefaultPointRepresentation<PointDefault>(*this)); }
  void copyToFloatArray(const PointDefault& p, float* out) const override
  {
    out[0] = p.x;
    out[1] = p.y;
    out[2] = p.z;
  }
};

// point_representation.h:560-604: a representation with a chosen number of leading dimensions
template <typename PointDefault>
class CustomPointRepresentation : public PointRepresentation<PointDefault> {
public:
  using Ptr = std::shared_ptr<CustomPointRepresentation<PointDefault>>;
  using ConstPtr = std::shared_ptr<const CustomPointRepresentation<PointDefault>>;
  explicit CustomPointRepresentation(const int max_dim = 3, const int start_dim = 0) : start_dim_(start_dim)
  {
    this->nr_dimensions_ = max_dim < 3 - start_dim ? max_dim : 3 - start_dim;
    if (this->nr_dimensions_ < 0) this->nr_dimensions_ = 0;
  }
  void copyToFloatArray(const PointDefault& p, float* out) const override
  {
    const float v[3] = {p.x, p.y, p.z};
    for (int i = 0; i < this->nr_dimensions_; ++i) out[i] = v[start_dim_ + i];
  }

protected:
  int start_dim_;
};

}  // namespace pcl
