// pcl/registration/correspondence_rejection*.h — the four correspondence rejectors of SURVEY.md §8f #1 on the device.
// Reference: registration/include/pcl/registration/correspondence_rejection.h:52-200 (base class),
// correspondence_rejection_distance.h, …_median_distance.h, …_one_to_one.h, …_trimmed.h and their src/*.cpp.
#pragma once
#include <cmath>
#include <cstdio>
#include <limits>
#include <memory>
#include <string>

#include "../b200/context.h"
#include "../correspondence.h"

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
  const std::string& getClassName() const { return rejection_name_; }
  virtual bool requiresSourcePoints() const { return false; }
  virtual bool requiresSourceNormals() const { return false; }
  virtual bool requiresTargetPoints() const { return false; }
  virtual bool requiresTargetNormals() const { return false; }
  virtual pclb200_rejector abiRejector() const = 0;  // lets ICP run the chain inside the device loop

protected:
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

protected:
  double factor_ = 1.0;
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
