// pcl/search/search.h — pcl::search::Search<PointT> (search/include/pcl/search/search.h:73-437, impl/search.hpp:44-214):
// the abstract nearest-neighbour interface PCL code holds searchers through.  Two per-point virtuals are pure; every
// other form is expressed through them here, so any subclass is complete once it implements those two — and
// pcl::search::KdTree (the device searcher) overrides the batch forms with ONE launch per call.
#pragma once
#include <algorithm>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../point_cloud.h"
#include "../types.h"

namespace pcl {
namespace search {

template <typename PointT>
class Search {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  using PointCloudPtr = typename PointCloud::Ptr;
  using PointCloudConstPtr = typename PointCloud::ConstPtr;
  using Ptr = std::shared_ptr<pcl::search::Search<PointT>>;
  using ConstPtr = std::shared_ptr<const pcl::search::Search<PointT>>;

  explicit Search(const std::string& name = "", bool sorted = false) : search_sorted_(sorted), search_name_(name) {}
  virtual ~Search() = default;

  virtual const std::string& getName() const { return search_name_; }
  virtual void setSortedResults(bool sorted) { search_sorted_ = sorted; }
  virtual bool getSortedResults() const { return search_sorted_; }

  virtual bool setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& indices = IndicesConstPtr())
  {
    search_input_ = cloud;
    search_indices_ = indices;
    return true;
  }
  virtual PointCloudConstPtr getInputCloud() const { return search_input_; }
  virtual IndicesConstPtr getIndices() const { return search_indices_; }

  // ---- k nearest neighbours ------------------------------------------------------------------------------
  virtual int nearestKSearch(const PointT& point, int k, Indices& k_indices, std::vector<float>& k_sqr_distances) const = 0;
  template <typename PointTDiff>
  int nearestKSearchT(const PointTDiff& point, int k, Indices& k_indices, std::vector<float>& k_sqr_distances) const
  {
    PointT p;
    p.x = point.x; p.y = point.y; p.z = point.z;   // copyPoint: the coordinates are what a searcher compares
    return nearestKSearch(p, k, k_indices, k_sqr_distances);
  }
  virtual int nearestKSearch(const PointCloud& cloud, index_t index, int k, Indices& k_indices,
                             std::vector<float>& k_sqr_distances) const
  {
    return nearestKSearch(cloud[static_cast<std::size_t>(index)], k, k_indices, k_sqr_distances);
  }
  // index into the input cloud, or into the index list the searcher was given (impl/search.hpp:93-108)
  virtual int nearestKSearch(index_t index, int k, Indices& k_indices, std::vector<float>& k_sqr_distances) const
  {
    const PointCloudConstPtr in = getInputCloud();
    const IndicesConstPtr idx = getIndices();
    return nearestKSearch((*in)[static_cast<std::size_t>(idx ? (*idx)[static_cast<std::size_t>(index)] : index)], k, k_indices,
                          k_sqr_distances);
  }
  virtual void nearestKSearch(const PointCloud& cloud, const Indices& indices, int k, std::vector<Indices>& k_indices,
                              std::vector<std::vector<float>>& k_sqr_distances) const
  {
    const std::size_t n = indices.empty() ? cloud.size() : indices.size();
    k_indices.assign(n, Indices());
    k_sqr_distances.assign(n, std::vector<float>());
    for (std::size_t i = 0; i < n; ++i)
      nearestKSearch(cloud, indices.empty() ? static_cast<index_t>(i) : indices[i], k, k_indices[i], k_sqr_distances[i]);
  }

  // ---- radius ----------------------------------------------------------------------------------------------
  virtual int radiusSearch(const PointT& point, double radius, Indices& k_indices, std::vector<float>& k_sqr_distances,
                           unsigned int max_nn = 0) const = 0;
  template <typename PointTDiff>
  int radiusSearchT(const PointTDiff& point, double radius, Indices& k_indices, std::vector<float>& k_sqr_distances,
                    unsigned int max_nn = 0) const
  {
    PointT p;
    p.x = point.x; p.y = point.y; p.z = point.z;
    return radiusSearch(p, radius, k_indices, k_sqr_distances, max_nn);
  }
  virtual int radiusSearch(const PointCloud& cloud, index_t index, double radius, Indices& k_indices,
                           std::vector<float>& k_sqr_distances, unsigned int max_nn = 0) const
  {
    return radiusSearch(cloud[static_cast<std::size_t>(index)], radius, k_indices, k_sqr_distances, max_nn);
  }
  virtual int radiusSearch(index_t index, double radius, Indices& k_indices, std::vector<float>& k_sqr_distances,
                           unsigned int max_nn = 0) const
  {
    const PointCloudConstPtr in = getInputCloud();
    const IndicesConstPtr idx = getIndices();
    return radiusSearch((*in)[static_cast<std::size_t>(idx ? (*idx)[static_cast<std::size_t>(index)] : index)], radius, k_indices,
                        k_sqr_distances, max_nn);
  }
  virtual void radiusSearch(const PointCloud& cloud, const Indices& indices, double radius, std::vector<Indices>& k_indices,
                            std::vector<std::vector<float>>& k_sqr_distances, unsigned int max_nn = 0) const
  {
    const std::size_t n = indices.empty() ? cloud.size() : indices.size();
    k_indices.assign(n, Indices());
    k_sqr_distances.assign(n, std::vector<float>());
    for (std::size_t i = 0; i < n; ++i)
      radiusSearch(cloud, indices.empty() ? static_cast<index_t>(i) : indices[i], radius, k_indices[i], k_sqr_distances[i], max_nn);
  }

  // search.h:401-411: thread count of the batch forms (a host-side notion; kept for the subclasses that loop)
  void setNumberOfThreads(unsigned int nr_threads) { num_threads_ = nr_threads ? nr_threads : 1u; }
  unsigned int getNumberOfThreads() const { return num_threads_; }

protected:
  // impl/search.hpp:196-214: order a result pair by distance (for searchers whose raw output is unordered)
  static void sortResults(Indices& indices, std::vector<float>& distances)
  {
    std::vector<std::size_t> order(indices.size());
    for (std::size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return distances[a] < distances[b]; });
    Indices si(indices.size());
    std::vector<float> sd(distances.size());
    for (std::size_t i = 0; i < order.size(); ++i) { si[i] = indices[order[i]]; sd[i] = distances[order[i]]; }
    indices.swap(si);
    distances.swap(sd);
  }
  bool search_sorted_;
  std::string search_name_;
  PointCloudConstPtr search_input_;
  IndicesConstPtr search_indices_;
  unsigned int num_threads_ = 1;
};

}  // namespace search
}  // namespace pcl
