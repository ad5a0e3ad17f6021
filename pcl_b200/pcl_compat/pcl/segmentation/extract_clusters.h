// pcl/segmentation/extract_clusters.h — pcl::EuclideanClusterExtraction / pcl::extractEuclideanClusters on the device
// (another consumer of the searcher).
// Reference: segmentation/include/pcl/segmentation/extract_clusters.h:54-125 (free functions), :305-445 (class),
// impl/extract_clusters.hpp:45-257.  The device returns, for every point, the smallest index of its connected component
// (pclb200_cluster_labels); grouping, the [min, max] size window (:98, :196) and the ordering by size (:249) are the
// reference's own host steps on 4 bytes per point.  Equal-sized clusters come out ordered by their smallest index (the
// reference's std::sort leaves that order unspecified).
#pragma once
#include <algorithm>
#include <cstdio>
#include <limits>
#include <vector>

#include "../PointIndices.h"
#include "../search/kdtree.h"

namespace pcl {
namespace detail {
// labels (smallest index of the component, -1 = not clustered) -> clusters in seed order, indices ascending (:98-113)
inline void clustersFromLabels(const std::vector<index_t>& labels, const PCLHeader& header, std::vector<PointIndices>& clusters,
                               unsigned int min_pts, unsigned int max_pts)
{
  std::vector<std::size_t> count(labels.size(), 0);
  for (index_t l : labels)
    if (l >= 0) ++count[static_cast<std::size_t>(l)];
  std::vector<std::ptrdiff_t> slot(labels.size(), -1);
  for (std::size_t i = 0; i < labels.size(); ++i)
    if (count[i] > 0 && count[i] >= min_pts && count[i] <= max_pts) {  // i is the seed (= smallest index) of a kept cluster
      slot[i] = static_cast<std::ptrdiff_t>(clusters.size());
      clusters.emplace_back();
      clusters.back().header = header;
      clusters.back().indices.reserve(count[i]);
    }
  for (std::size_t i = 0; i < labels.size(); ++i)
    if (labels[i] >= 0 && slot[static_cast<std::size_t>(labels[i])] >= 0)
      clusters[static_cast<std::size_t>(slot[static_cast<std::size_t>(labels[i])])].indices.push_back(static_cast<index_t>(i));
}
}  // namespace detail

// impl/extract_clusters.hpp:45-119 — the tree must have been built over `cloud`
template <typename PointT>
void extractEuclideanClusters(const PointCloud<PointT>& cloud, const typename search::KdTree<PointT>::Ptr& tree, float tolerance,
                              std::vector<PointIndices>& clusters, unsigned int min_pts_per_cluster = 1,
                              unsigned int max_pts_per_cluster = std::numeric_limits<int>::max())
{
  if (!tree || !tree->getInputCloud() || tree->getInputCloud()->size() != cloud.size()) {
    std::fprintf(stderr, "[pcl::extractEuclideanClusters] Tree built for a different point cloud dataset than the input cloud (%zu)!\n",
                 cloud.size());
    return;
  }
  if (!tree->deviceIndex()) return;
  std::vector<index_t> labels(cloud.size(), -1);
  if (pclb200_cluster_labels(b200::Context::get(), tree->deviceIndex(), static_cast<double>(tolerance), labels.data(), labels.size()) !=
      PCLB200_OK) {
    std::fprintf(stderr, "[pcl::extractEuclideanClusters] %s\n", pclb200_last_error());
    return;
  }
  detail::clustersFromLabels(labels, cloud.header, clusters, min_pts_per_cluster, max_pts_per_cluster);
}

// impl/extract_clusters.hpp:124-223 — the tree must have been built over <cloud, indices>
template <typename PointT>
void extractEuclideanClusters(const PointCloud<PointT>& cloud, const Indices& indices, const typename search::KdTree<PointT>::Ptr& tree,
                              float tolerance, std::vector<PointIndices>& clusters, unsigned int min_pts_per_cluster = 1,
                              unsigned int max_pts_per_cluster = std::numeric_limits<int>::max())
{
  (void)indices;  // the labels of points outside the tree's subset are -1 already
  extractEuclideanClusters<PointT>(cloud, tree, tolerance, clusters, min_pts_per_cluster, max_pts_per_cluster);
}

inline bool comparePointClusters(const PointIndices& a, const PointIndices& b) { return a.indices.size() < b.indices.size(); }

template <typename PointT>
class EuclideanClusterExtraction : public PCLBase<PointT> {
public:
  using KdTree = pcl::search::KdTree<PointT>;
  using KdTreePtr = typename KdTree::Ptr;
  void setSearchMethod(const KdTreePtr& tree) { tree_ = tree; }
  // extract_clusters.h:336-353: the reference's parameter type is pcl::search::Search<PointT>::Ptr
  void setSearchMethod(const typename pcl::search::Search<PointT>::Ptr& tree)
  {
    tree_ = pcl::search::deviceSearcher<PointT>(tree, "pcl::EuclideanClusterExtraction");
  }
  KdTreePtr getSearchMethod() const { return tree_; }
  void setClusterTolerance(double tolerance) { cluster_tolerance_ = tolerance; }
  double getClusterTolerance() const { return cluster_tolerance_; }
  void setMinClusterSize(uindex_t n) { min_pts_per_cluster_ = n; }
  uindex_t getMinClusterSize() const { return min_pts_per_cluster_; }
  void setMaxClusterSize(uindex_t n) { max_pts_per_cluster_ = n; }
  uindex_t getMaxClusterSize() const { return max_pts_per_cluster_; }

  // impl/extract_clusters.hpp:225-252
  void extract(std::vector<PointIndices>& clusters)
  {
    clusters.clear();
    if (!PCLBase<PointT>::initCompute() || this->input_->empty() || this->indices_->empty()) return;
    if (!tree_) tree_.reset(new KdTree(false));
    tree_->setInputCloud(this->input_, this->fake_indices_ ? IndicesConstPtr() : IndicesConstPtr(this->indices_));
    extractEuclideanClusters<PointT>(*this->input_, tree_, static_cast<float>(cluster_tolerance_), clusters, min_pts_per_cluster_,
                                     max_pts_per_cluster_);
    // "Sort the clusters based on their size (largest one first)"; stable, so equal sizes keep the seed order
    std::stable_sort(clusters.begin(), clusters.end(),
                     [](const PointIndices& a, const PointIndices& b) { return a.indices.size() > b.indices.size(); });
  }

protected:
  KdTreePtr tree_;
  double cluster_tolerance_ = 0.0;
  uindex_t min_pts_per_cluster_ = 1;
  uindex_t max_pts_per_cluster_ = std::numeric_limits<uindex_t>::max();
};
}  // namespace pcl
