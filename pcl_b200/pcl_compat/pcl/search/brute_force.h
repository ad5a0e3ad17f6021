// pcl/search/brute_force.h — pcl::search::BruteForce<PointT>: the linear-scan searcher the reference's tests use as ground
// truth for every other searcher (search/include/pcl/search/brute_force.h:51-139, impl/brute_force.hpp:45-351).  Host code:
// it exists so that programs (and the reference's test_search.cpp scenarios) that cross-check a searcher against it run
// unchanged; the device searcher is pcl::search::KdTree.
//   nearestKSearch: the k smallest squared distances, ascending (a bounded max-heap, like the reference's priority queue);
//   radiusSearch:   every point with d2 <= radius^2 (NOT strict, unlike FLANN's), in index order and cut after max_nn hits,
//                   sorted afterwards only if setSortedResults(true);
//   non-finite input points are skipped when the cloud is not dense.
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>
#include <queue>
#include <vector>

#include "search.h"

namespace pcl {
namespace search {

template <typename PointT>
class BruteForce : public Search<PointT> {
public:
  using PointCloud = typename Search<PointT>::PointCloud;
  using PointCloudConstPtr = typename Search<PointT>::PointCloudConstPtr;
  using Ptr = std::shared_ptr<BruteForce<PointT>>;
  using ConstPtr = std::shared_ptr<const BruteForce<PointT>>;
  using Search<PointT>::nearestKSearch;
  using Search<PointT>::radiusSearch;

  explicit BruteForce(bool sorted_results = false) : Search<PointT>("BruteForce", sorted_results) {}

  int nearestKSearch(const PointT& point, int k, Indices& k_indices, std::vector<float>& k_distances) const override
  {
    k_indices.clear();
    k_distances.clear();
    const PointCloudConstPtr input = this->getInputCloud();
    if (k < 1 || !input) return 0;
    std::priority_queue<Entry> queue;
    forEachCandidate(*input, point, [&](index_t idx, float d2) {
      if (queue.size() < static_cast<std::size_t>(k)) queue.push(Entry{idx, d2});
      else if (queue.top().distance > d2) { queue.pop(); queue.push(Entry{idx, d2}); }
      return true;
    });
    k_indices.resize(queue.size());
    k_distances.resize(queue.size());
    for (std::size_t i = queue.size(); i-- > 0;) {
      k_indices[i] = queue.top().index;
      k_distances[i] = queue.top().distance;
      queue.pop();
    }
    return static_cast<int>(k_indices.size());
  }

  int radiusSearch(const PointT& point, double radius, Indices& k_indices, std::vector<float>& k_sqr_distances,
                   unsigned int max_nn = 0) const override
  {
    k_indices.clear();
    k_sqr_distances.clear();
    const PointCloudConstPtr input = this->getInputCloud();
    if (!input) return 0;
    radius *= radius;
    forEachCandidate(*input, point, [&](index_t idx, float d2) {
      if (d2 <= radius) {
        k_indices.push_back(idx);
        k_sqr_distances.push_back(d2);
        if (k_indices.size() == max_nn) return false;  // never true for max_nn = 0
      }
      return true;
    });
    if (this->getSortedResults()) {  // Search::sortResults: ascending distance
      std::vector<std::size_t> order(k_indices.size());
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return k_sqr_distances[a] < k_sqr_distances[b]; });
      Indices si(order.size());
      std::vector<float> sd(order.size());
      for (std::size_t i = 0; i < order.size(); ++i) { si[i] = k_indices[order[i]]; sd[i] = k_sqr_distances[order[i]]; }
      k_indices.swap(si);
      k_sqr_distances.swap(sd);
    }
    return static_cast<int>(k_indices.size());
  }

private:
  struct Entry {
    index_t index;
    float distance;
    bool operator<(const Entry& o) const { return distance < o.distance; }
  };
  // visits (index, squared distance to the query) over the view or the whole cloud, in index-list order; fn returns false to stop
  template <typename Fn>
  void forEachCandidate(const PointCloud& cloud, const PointT& query, Fn fn) const
  {
    const IndicesConstPtr view = this->getIndices();
    const bool sparse = !cloud.is_dense;
    const std::size_t n = view ? view->size() : cloud.size();
    for (std::size_t i = 0; i < n; ++i) {
      const index_t idx = view ? (*view)[i] : static_cast<index_t>(i);
      const PointT& p = cloud[static_cast<std::size_t>(idx)];
      if (sparse && !std::isfinite(p.x)) continue;
      const float dx = p.x - query.x, dy = p.y - query.y, dz = p.z - query.z;
      if (!fn(idx, dx * dx + dy * dy + dz * dz)) return;
    }
  }
};

}  // namespace search
}  // namespace pcl
