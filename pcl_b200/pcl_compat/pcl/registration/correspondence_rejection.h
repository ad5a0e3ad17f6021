// pcl/registration/correspondence_rejection*.h — the four correspondence rejectors of SURVEY.md §8f #1 on the device.
// Reference: registration/include/pcl/registration/correspondence_rejection.h:52-200 (base class),
// correspondence_rejection_distance.h, …_median_distance.h, …_one_to_one.h, …_trimmed.h and their src/*.cpp.
#pragma once
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <functional>
#include <vector>
#include <limits>
#include <memory>
#include <string>

#include "../PCLPointCloud2.h"
#include "../b200/context.h"
#include "../correspondence.h"
#include "../point_cloud.h"

namespace pcl {
namespace registration {

class CorrespondenceRejector {
public:
  using Ptr = std::shared_ptr<CorrespondenceRejector>;
  using ConstPtr = std::shared_ptr<const CorrespondenceRejector>;
  virtual ~CorrespondenceRejector() = default;
  virtual void setInputCorrespondences(const CorrespondencesConstPtr& c) { input_correspondences_ = c; }
  CorrespondencesConstPtr getInputCorrespondences() { return input_correspondences_; }
  void getCorrespondences(pcl::Correspondences& out)
  {
    if (!input_correspondences_ || input_correspondences_->empty()) { out.clear(); return; }
    getRemainingCorrespondences(*input_correspondences_, out);
  }
  virtual void getRemainingCorrespondences(const pcl::Correspondences& in, pcl::Correspondences& out)
  {
    out.resize(in.size());
    std::size_t n = 0;
    const pclb200_rejector r = abiRejector();
    double med = 0.0;
    if (pclb200_reject(b200::Context::get(), &r, reinterpret_cast<const pclb200_corr*>(in.data()), in.size(),
                       reinterpret_cast<pclb200_corr*>(out.data()), &n, &med) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::registration::%s::getRemainingCorrespondences] %s\n", getClassName().c_str(), pclb200_last_error());
      n = 0;
    }
    out.resize(n);
    last_median_ = med;
  }
  // correspondence_rejection.h:103-121: which query points the step dropped, relative to the INPUT correspondences
  void getRejectedQueryIndices(const pcl::Correspondences& correspondences, pcl::Indices& indices)
  {
    if (!input_correspondences_ || input_correspondences_->empty()) {
      std::fprintf(stderr, "[pcl::registration::%s::getRejectedQueryIndices] Input correspondences not set (lookup of rejected "
                           "correspondences _not_ possible).\n", getClassName().c_str());
      return;
    }
    pcl::getRejectedQueryIndices(*input_correspondences_, correspondences, indices);
  }
  const std::string& getClassName() const { return rejection_name_; }
  // correspondence_rejection.h:130-196: the type-erased route by which IterativeClosestPoint hands a rejector the clouds
  // or normals it asks for (impl/icp.hpp:142-155); a rejector that needs none says so
  virtual bool requiresSourcePoints() const { return false; }
  virtual void setSourcePoints(pcl::PCLPointCloud2::ConstPtr /*cloud2*/) { notRequired("setSourcePoints", "an input source cloud"); }
  virtual bool requiresSourceNormals() const { return false; }
  virtual void setSourceNormals(pcl::PCLPointCloud2::ConstPtr /*cloud2*/) { notRequired("setSourceNormals", "input source normals"); }
  virtual bool requiresTargetPoints() const { return false; }
  virtual void setTargetPoints(pcl::PCLPointCloud2::ConstPtr /*cloud2*/) { notRequired("setTargetPoints", "an input target cloud"); }
  virtual bool requiresTargetNormals() const { return false; }
  virtual void setTargetNormals(pcl::PCLPointCloud2::ConstPtr /*cloud2*/) { notRequired("setTargetNormals", "input target normals"); }
  virtual pclb200_rejector abiRejector() const = 0;  // lets ICP run the chain inside the device loop
  // false for a rejector that only exists on the host (CorrespondenceRejectorSampleConsensus): an ICP holding one runs the
  // stage-by-stage loop and calls getRemainingCorrespondences between the device stages
  virtual bool runsOnDevice() const { return true; }

protected:
  void notRequired(const char* method, const char* what) const
  {
    std::fprintf(stderr, "[pcl::registration::%s::%s] This class does not require %s\n", getClassName().c_str(), method, what);
  }
  std::string rejection_name_;
  CorrespondencesConstPtr input_correspondences_;
  double last_median_ = 0.0;
};

class CorrespondenceRejectorDistance : public CorrespondenceRejector {
public:
  using Ptr = std::shared_ptr<CorrespondenceRejectorDistance>;
  CorrespondenceRejectorDistance() { rejection_name_ = "CorrespondenceRejectorDistance"; }
  void setMaximumDistance(float d) { max_distance_ = d; }
  float getMaximumDistance() const { return max_distance_; }
  pclb200_rejector abiRejector() const override { return pclb200_rejector{PCLB200_REJ_DISTANCE, 0, max_distance_}; }

protected:
  float max_distance_ = std::sqrt(std::numeric_limits<float>::max());
};

class CorrespondenceRejectorMedianDistance : public CorrespondenceRejector {
public:
  using Ptr = std::shared_ptr<CorrespondenceRejectorMedianDistance>;
  CorrespondenceRejectorMedianDistance() { rejection_name_ = "CorrespondenceRejectorMedianDistance"; }
  void setMedianFactor(double f) { factor_ = f; }
  double getMedianFactor() const { return factor_; }
  double getMedianDistance() const { return last_median_; }
  pclb200_rejector abiRejector() const override { return pclb200_rejector{PCLB200_REJ_MEDIAN, 0, factor_}; }
  // correspondence_rejection_median_distance.h:100-130: with a source AND a target cloud the score of a pair is the squared
  // distance of the two POINTS (DataContainer::getCorrespondenceScore), not Correspondence::distance; the stored distances
  // are what an ICP loop keeps current, so inside ICP the two agree and the device form is used; stand-alone the scored form
  // runs on the host (src/correspondence_rejection_median_distance.cpp:44-70)
  template <typename PointT>
  void setInputSource(const typename pcl::PointCloud<PointT>::ConstPtr& cloud)
  {
    source_ = cloud;
    rescore<PointT>();
  }
  template <typename PointT>
  void setInputTarget(const typename pcl::PointCloud<PointT>::ConstPtr& cloud)
  {
    target_ = cloud;
    rescore<PointT>();
  }
  void getRemainingCorrespondences(const pcl::Correspondences& in, pcl::Correspondences& out) override
  {
    if (!score_) { CorrespondenceRejector::getRemainingCorrespondences(in, out); return; }
    out.clear();
    if (in.empty()) return;
    std::vector<double> dists(in.size());
    for (std::size_t i = 0; i < in.size(); ++i) dists[i] = score_(in[i]);
    std::vector<double> nth(dists);
    std::nth_element(nth.begin(), nth.begin() + static_cast<std::ptrdiff_t>(nth.size() / 2), nth.end());
    last_median_ = nth[nth.size() / 2];
    for (std::size_t i = 0; i < in.size(); ++i)
      if (dists[i] <= last_median_ * factor_) out.push_back(in[i]);
  }

protected:
  template <typename PointT>
  void rescore()
  {
    auto s = std::static_pointer_cast<const pcl::PointCloud<PointT>>(source_);
    auto t = std::static_pointer_cast<const pcl::PointCloud<PointT>>(target_);
    if (!s || !t) { score_ = nullptr; return; }
    score_ = [s, t](const pcl::Correspondence& c) {
      const PointT& a = (*s)[static_cast<std::size_t>(c.index_query)];
      const PointT& b = (*t)[static_cast<std::size_t>(c.index_match)];
      const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
      return static_cast<double>(dx * dx + dy * dy + dz * dz);
    };
  }
  double factor_ = 1.0;
  std::shared_ptr<const void> source_, target_;
  std::function<double(const pcl::Correspondence&)> score_;
};

class CorrespondenceRejectorOneToOne : public CorrespondenceRejector {
public:
  using Ptr = std::shared_ptr<CorrespondenceRejectorOneToOne>;
  CorrespondenceRejectorOneToOne() { rejection_name_ = "CorrespondenceRejectorOneToOne"; }
  pclb200_rejector abiRejector() const override { return pclb200_rejector{PCLB200_REJ_ONE_TO_ONE, 0, 0.0}; }
};

class CorrespondenceRejectorTrimmed : public CorrespondenceRejector {
public:
  using Ptr = std::shared_ptr<CorrespondenceRejectorTrimmed>;
  CorrespondenceRejectorTrimmed() { rejection_name_ = "CorrespondenceRejectorTrimmed"; }
  void setOverlapRatio(float r) { overlap_ratio_ = std::min(1.0f, std::max(0.0f, r)); }
  float getOverlapRatio() const { return overlap_ratio_; }
  void setMinCorrespondences(unsigned n) { nr_min_correspondences_ = n; }
  unsigned getMinCorrespondences() const { return nr_min_correspondences_; }
  pclb200_rejector abiRejector() const override
  {
    return pclb200_rejector{PCLB200_REJ_TRIMMED, static_cast<int32_t>(nr_min_correspondences_), overlap_ratio_};
  }

protected:
  float overlap_ratio_ = 0.5f;
  unsigned nr_min_correspondences_ = 0;
};

}  // namespace registration
}  // namespace pcl
