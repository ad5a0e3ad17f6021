// pcl/registration/correspondence_rejection_var_trimmed.h — CorrespondenceRejectorVarTrimmed: the trimming ratio is chosen by
// minimising the FRMS of "Outlier Robust ICP for Minimizing Fractional RMSD" over [min_ratio, max_ratio].
// Reference: registration/include/pcl/registration/correspondence_rejection_var_trimmed.h:60-250,
// registration/src/correspondence_rejection_var_trimmed.cpp:42-110.  Host code (a sort and two short passes; like
// CorrespondenceRejectorSampleConsensus an ICP that holds one runs the stage-by-stage loop).
// Restated as the reference BEHAVES, including two things a reader of its header would not guess: FRMS(j) uses the sum of the
// distances below min_ratio plus the ONE distance at j (not the running sum), and the final pass compares the SORTED distances
// with the threshold while walking the correspondences in INPUT order — so the first m input correspondences are kept, m being
// the number of distances below the threshold.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <typeindex>
#include <typeinfo>
#include <vector>

#include "../conversions.h"
#include "../point_cloud.h"
#include "../point_types.h"
#include "correspondence_rejection.h"

namespace pcl {
namespace registration {
namespace detail {
// DataContainer::getCorrespondenceScore (correspondence_rejection.h:300-320): the squared distance of the two points, in float
template <typename PointT>
inline std::function<double(const pcl::Correspondence&)> pairScore(const typename pcl::PointCloud<PointT>::ConstPtr& source,
                                                                   const typename pcl::PointCloud<PointT>::ConstPtr& target)
{
  return [source, target](const pcl::Correspondence& c) {
    const PointT& s = (*source)[static_cast<std::size_t>(c.index_query)];
    const PointT& t = (*target)[static_cast<std::size_t>(c.index_match)];
    const float dx = s.x - t.x, dy = s.y - t.y, dz = s.z - t.z;
    return static_cast<double>(dx * dx + dy * dy + dz * dz);
  };
}
// keeps the (source, target) pair a rejector was given through setInputSource<PointT> / setInputTarget<PointT> and scores with it
struct ScoreContainer {
  std::function<double(const pcl::Correspondence&)> score;   // empty: use Correspondence::distance
  std::function<void()> rebuild;
  template <typename PointT>
  void setSource(const typename pcl::PointCloud<PointT>::ConstPtr& cloud)
  {
    auto& slot = holder<PointT>();
    slot.first = cloud;
    refresh<PointT>();
  }
  template <typename PointT>
  void setTarget(const typename pcl::PointCloud<PointT>::ConstPtr& cloud)
  {
    auto& slot = holder<PointT>();
    slot.second = cloud;
    refresh<PointT>();
  }
  double operator()(const pcl::Correspondence& c) const { return score ? score(c) : static_cast<double>(c.distance); }

private:
  template <typename PointT>
  std::pair<typename pcl::PointCloud<PointT>::ConstPtr, typename pcl::PointCloud<PointT>::ConstPtr>& holder()
  {
    using Pair = std::pair<typename pcl::PointCloud<PointT>::ConstPtr, typename pcl::PointCloud<PointT>::ConstPtr>;
    if (!store_ || tag_ != std::type_index(typeid(PointT))) {
      store_ = std::make_shared<Pair>();
      tag_ = std::type_index(typeid(PointT));
    }
    return *std::static_pointer_cast<Pair>(store_);
  }
  template <typename PointT>
  void refresh()
  {
    auto& slot = holder<PointT>();
    if (slot.first && slot.second) score = pairScore<PointT>(slot.first, slot.second);
    else score = nullptr;
  }
  std::shared_ptr<void> store_;
  std::type_index tag_ = std::type_index(typeid(void));
};
}  // namespace detail

class CorrespondenceRejectorVarTrimmed : public CorrespondenceRejector {
public:
  using Ptr = std::shared_ptr<CorrespondenceRejectorVarTrimmed>;
  using ConstPtr = std::shared_ptr<const CorrespondenceRejectorVarTrimmed>;
  CorrespondenceRejectorVarTrimmed() { rejection_name_ = "CorrespondenceRejectorVarTrimmed"; }

  void getRemainingCorrespondences(const pcl::Correspondences& original, pcl::Correspondences& remaining) override
  {
    remaining.clear();
    if (original.empty()) return;
    std::vector<double> dists(original.size());
    for (std::size_t i = 0; i < original.size(); ++i) dists[i] = container_(original[i]);
    factor_ = optimizeInlierRatio(dists);   // sorts dists
    const std::size_t at = std::min(dists.size() - 1, static_cast<std::size_t>(static_cast<int>(static_cast<double>(dists.size()) * factor_)));
    trimmed_distance_ = dists[at];
    for (std::size_t i = 0; i < original.size(); ++i)
      if (dists[i] < trimmed_distance_) remaining.push_back(original[i]);   // dists is sorted here: the first m inputs survive
  }
  double getTrimmedDistance() const { return trimmed_distance_; }
  template <typename PointT> void setInputSource(const typename pcl::PointCloud<PointT>::ConstPtr& cloud) { container_.template setSource<PointT>(cloud); }
  template <typename PointT> void setInputTarget(const typename pcl::PointCloud<PointT>::ConstPtr& cloud) { container_.template setTarget<PointT>(cloud); }
  bool requiresSourcePoints() const override { return true; }
  void setSourcePoints(pcl::PCLPointCloud2::ConstPtr cloud2) override
  {
    pcl::PointCloud<PointXYZ>::Ptr cloud(new pcl::PointCloud<PointXYZ>);
    fromPCLPointCloud2(*cloud2, *cloud);
    setInputSource<PointXYZ>(cloud);
  }
  bool requiresTargetPoints() const override { return true; }
  void setTargetPoints(pcl::PCLPointCloud2::ConstPtr cloud2) override
  {
    pcl::PointCloud<PointXYZ>::Ptr cloud(new pcl::PointCloud<PointXYZ>);
    fromPCLPointCloud2(*cloud2, *cloud);
    setInputTarget<PointXYZ>(cloud);
  }
  double getTrimFactor() const { return factor_; }
  void setMinRatio(double ratio) { min_ratio_ = ratio; }
  double getMinRatio() const { return min_ratio_; }
  void setMaxRatio(double ratio) { max_ratio_ = ratio; }
  double getMaxRatio() const { return max_ratio_; }

  bool runsOnDevice() const override { return false; }
  pclb200_rejector abiRejector() const override { return pclb200_rejector{-1, 0, 0.0}; }   // no device form

protected:
  // src/correspondence_rejection_var_trimmed.cpp:82-110
  float optimizeInlierRatio(std::vector<double>& dists) const
  {
    const unsigned int points_nbr = static_cast<unsigned int>(dists.size());
    std::sort(dists.begin(), dists.end());
    const int min_el = static_cast<int>(std::floor(min_ratio_ * points_nbr));
    const int max_el = static_cast<int>(std::floor(max_ratio_ * points_nbr));
    double lower_sum = 0.0;
    for (int i = 0; i < min_el; ++i) lower_sum += dists[static_cast<std::size_t>(i)];
    int min_index = 0;
    double best = 0.0;
    for (int j = 0; j < max_el - min_el; ++j) {
      const double id = static_cast<double>(min_el + 1 + j);   // LinSpaced(max_el - min_el, min_el + 1, max_el): step 1
      const double deno = std::pow(id / points_nbr, lambda_);
      const double frms = (1.0 / deno) * (1.0 / deno) * (1.0 / id) * (lower_sum + dists[static_cast<std::size_t>(min_el + j)]);
      if (j == 0 || frms < best) { best = frms; min_index = j; }
    }
    return static_cast<float>(min_index + min_el) / static_cast<float>(points_nbr);
  }

  double trimmed_distance_ = 0.0;
  double factor_ = 0.0;
  double min_ratio_ = 0.05, max_ratio_ = 0.95, lambda_ = 0.95;
  detail::ScoreContainer container_;
};
}  // namespace registration
}  // namespace pcl
