// pcl/search/kdtree.h — pcl::search::KdTree<PointT> backed by the device LBVH.
//
// Mirrors the public surface of search/include/pcl/search/search.h:73-437 + kdtree.h:61-168 that the ICP path
// uses: setInputCloud, nearestKSearch / radiusSearch (per point, per index, and the batch overloads of
// impl/search.hpp:111-194), setSortedResults, setEpsilon.  Conventions preserved: output vectors
// are RESIZED by the callee, return value = number of neighbours (0 on error), indices refer to the ORIGINAL cloud,
// distances are squared, k is clamped to the number of valid points.  With real PCL headers this class derives
// from pcl::search::KdTree<PointT> through its protected (name, sorted) constructor exactly like
// pcl::search::KdTreeNanoflann does (search/include/pcl/search/kdtree_nanoflann.h:186,310-325) — INTEGRATION.md.
#pragma once
#include <cstdio>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../b200/context.h"
#include "../point_cloud.h"
#include "../point_representation.h"
#include "../types.h"
#include "search.h"

namespace pcl {
namespace search {

template <typename PointT>
class KdTree : public Search<PointT> {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  using PointCloudConstPtr = typename PointCloud::ConstPtr;
  using Ptr = std::shared_ptr<KdTree<PointT>>;
  using ConstPtr = std::shared_ptr<const KdTree<PointT>>;
  using PointRepresentationConstPtr = typename pcl::PointRepresentation<PointT>::ConstPtr;

  explicit KdTree(bool sorted = true) : Search<PointT>("KdTree", sorted), sorted_results_(sorted) {}

  // kdtree.h:104-119 / search/kdtree.h:106-119: the representation decides which float vector of a point is indexed.
  // A cloud that is already set is re-indexed at once, like the reference does.  Up to three dimensions.
  void setPointRepresentation(const PointRepresentationConstPtr& rep)
  {
    point_representation_ = rep;
    if (input_)  // KdTreeFLANN::setPointRepresentation re-indexes a cloud that is already set (kdtree_flann.hpp)
      setInputCloud(input_, indices_);
  }
  PointRepresentationConstPtr getPointRepresentation() const { return point_representation_; }
  virtual ~KdTree() = default;

  const std::string& getName() const override { return name_; }
  void setSortedResults(bool sorted) override { sorted_results_ = sorted; }
  bool getSortedResults() const override { return sorted_results_; }
  void setEpsilon(float eps) { epsilon_ = eps; }  // accepted for API parity; the search is always exact (eps = 0)
  float getEpsilon() const { return epsilon_; }
  void setMinPts(int min_pts) { min_pts_ = min_pts; }  // kdtree.h:322-333 (stored, never consulted — as in KdTreeFLANN)
  int getMinPts() const { return min_pts_; }

  // KdTreeFLANN::setInputCloud — kdtree_flann.hpp:100-136: rebuilds from scratch, epsilon reset to 0
  bool setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& indices = IndicesConstPtr()) override
  {
    input_ = cloud;
    indices_ = indices;
    epsilon_ = 0.f;
    index_.reset();
    if (!cloud || cloud->empty()) {
      std::fprintf(stderr, "[pcl::search::KdTree::setInputCloud] Invalid input!\n");
      return false;
    }
    pclb200_index* h = nullptr;
    int rc;
    vectorized_ = point_representation_ && !point_representation_->isTrivial();
    if (vectorized_) {
      // KdTreeFLANN::convertCloudToArray (kdtree_flann.hpp:429-498): valid points only, through vectorize(); an invalid
      // point becomes a NaN row, which the index drops while keeping the original numbering (index_mapping_)
      const int dim = point_representation_->getNumberOfDimensions();
      if (dim < 1 || dim > 3) {
        std::fprintf(stderr, "[pcl::search::KdTree::setInputCloud] the device index is 3-D: a %d-dimensional point representation is not supported\n", dim);
        return false;
      }
      std::vector<float> v(3 * cloud->size(), 0.f);
      for (std::size_t i = 0; i < cloud->size(); ++i) {
        if (point_representation_->isValid((*cloud)[i]))
          point_representation_->vectorize((*cloud)[i], &v[3 * i]);
        else
          v[3 * i] = std::numeric_limits<float>::quiet_NaN();
      }
      rc = pclb200_index_build(b200::Context::get(), v.data(), cloud->size(), 12, indices ? indices->data() : nullptr,
                               indices ? indices->size() : 0, &h);
    }
    else
      rc = pclb200_index_build(b200::Context::get(), cloud->points.data(), cloud->size(), sizeof(PointT),
                               indices ? indices->data() : nullptr, indices ? indices->size() : 0, &h);
    if (rc != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::search::KdTree::setInputCloud] %s\n", pclb200_last_error());
      return false;
    }
    index_.reset(new b200::IndexHandle(h));
    return true;
  }
  PointCloudConstPtr getInputCloud() const override { return input_; }
  IndicesConstPtr getIndices() const override { return indices_; }
  pclb200_index* deviceIndex() const { return index_ ? index_->h : nullptr; }
  bool usesRepresentationVectors() const { return vectorized_; }

  // ---- k-NN ---------------------------------------------------------------------------------------------
  int nearestKSearch(const PointT& point, int k, Indices& k_indices, std::vector<float>& k_sqr_distances) const override
  {
    return knn(&point, 1, k, &k_indices, &k_sqr_distances);
  }
  int nearestKSearch(const PointCloud& cloud, index_t index, int k, Indices& ki, std::vector<float>& kd) const override
  {
    return nearestKSearch(cloud[index], k, ki, kd);
  }
  int nearestKSearch(index_t index, int k, Indices& ki, std::vector<float>& kd) const override
  {
    return nearestKSearch((*input_)[indices_ ? (*indices_)[index] : index], k, ki, kd);
  }
  template <typename PointTDiff>
  int nearestKSearchT(const PointTDiff& p, int k, Indices& ki, std::vector<float>& kd) const
  {
    PointT q;
    q.x = p.x; q.y = p.y; q.z = p.z;
    return nearestKSearch(q, k, ki, kd);
  }
  // batch overload (search.h:216-219, impl/search.hpp:111-137): ONE device launch for the whole cloud
  void nearestKSearch(const PointCloud& cloud, const Indices& indices, int k, std::vector<Indices>& k_indices,
                      std::vector<std::vector<float>>& k_sqr_distances) const override
  {
    std::vector<PointT> q;
    const PointT* qp = cloud.points.data();
    std::size_t nq = cloud.size();
    if (!indices.empty()) {
      q.reserve(indices.size());
      for (index_t i : indices) q.push_back(cloud[i]);
      qp = q.data();
      nq = q.size();
    }
    k_indices.assign(nq, Indices());
    k_sqr_distances.assign(nq, std::vector<float>());
    if (!index_ || k <= 0 || nq == 0) return;
    std::vector<index_t> oi(nq * static_cast<std::size_t>(k));
    std::vector<float> od(nq * static_cast<std::size_t>(k));
    int keff = 0;
    const Queries qs = queries(qp, nq);
    if (pclb200_knn(b200::Context::get(), index_->h, qs.ptr, nq, qs.stride, k, oi.data(), od.data(), &keff) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::search::KdTree::nearestKSearch] %s\n", pclb200_last_error());
      return;
    }
    for (std::size_t i = 0; i < nq; ++i) {
      k_indices[i].assign(oi.begin() + i * k, oi.begin() + i * k + keff);
      k_sqr_distances[i].assign(od.begin() + i * k, od.begin() + i * k + keff);
    }
  }

  // search.h:229-260: queries of another point type — x, y, z are copied into PointT, then the batch overload
  template <typename PointTDiff>
  void nearestKSearchT(const pcl::PointCloud<PointTDiff>& cloud, const Indices& indices, int k,
                       std::vector<Indices>& k_indices, std::vector<std::vector<float>>& k_sqr_distances) const
  {
    PointCloud pc;
    copyXYZ(cloud, indices, pc);
    nearestKSearch(pc, Indices(), k, k_indices, k_sqr_distances);
  }

  // ---- radius ---------------------------------------------------------------------------------------------
  int radiusSearch(const PointT& point, double radius, Indices& k_indices, std::vector<float>& k_sqr_distances,
                   unsigned int max_nn = 0) const override
  {
    k_indices.clear();
    k_sqr_distances.clear();
    if (!index_) return 0;
    std::int64_t offs[2] = {0, 0};
    index_t* pi = nullptr;
    float* pd = nullptr;
    const Queries qs = queries(&point, 1);
    if (pclb200_radius(b200::Context::get(), index_->h, qs.ptr, 1, qs.stride, radius, max_nn, sorted_results_ ? 1 : 0,
                       offs, &pi, &pd) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::search::KdTree::radiusSearch] %s\n", pclb200_last_error());
      return 0;
    }
    k_indices.assign(pi, pi + offs[1]);
    k_sqr_distances.assign(pd, pd + offs[1]);
    pclb200_free(pi);
    pclb200_free(pd);
    return static_cast<int>(offs[1]);
  }
  int radiusSearch(index_t index, double radius, Indices& ki, std::vector<float>& kd, unsigned int max_nn = 0) const override
  {
    return radiusSearch((*input_)[indices_ ? (*indices_)[index] : index], radius, ki, kd, max_nn);
  }
  // search.h:311-315
  int radiusSearch(const PointCloud& cloud, index_t index, double radius, Indices& ki, std::vector<float>& kd,
                   unsigned int max_nn = 0) const override
  {
    return radiusSearch(cloud[index], radius, ki, kd, max_nn);
  }
  // search.h:285-292, 368-397
  template <typename PointTDiff>
  int radiusSearchT(const PointTDiff& p, double radius, Indices& ki, std::vector<float>& kd, unsigned int max_nn = 0) const
  {
    PointT q;
    q.x = p.x; q.y = p.y; q.z = p.z;
    return radiusSearch(q, radius, ki, kd, max_nn);
  }
  template <typename PointTDiff>
  void radiusSearchT(const pcl::PointCloud<PointTDiff>& cloud, const Indices& indices, double radius,
                     std::vector<Indices>& k_indices, std::vector<std::vector<float>>& k_sqr_distances,
                     unsigned int max_nn = 0) const
  {
    PointCloud pc;
    copyXYZ(cloud, indices, pc);
    radiusSearch(pc, Indices(), radius, k_indices, k_sqr_distances, max_nn);
  }
  // batch overload (search.h:349-355, impl/search.hpp:157-194)
  void radiusSearch(const PointCloud& cloud, const Indices& indices, double radius, std::vector<Indices>& k_indices,
                    std::vector<std::vector<float>>& k_sqr_distances, unsigned int max_nn = 0) const override
  {
    std::vector<PointT> q;
    const PointT* qp = cloud.points.data();
    std::size_t nq = cloud.size();
    if (!indices.empty()) {
      q.reserve(indices.size());
      for (index_t i : indices) q.push_back(cloud[i]);
      qp = q.data();
      nq = q.size();
    }
    k_indices.assign(nq, Indices());
    k_sqr_distances.assign(nq, std::vector<float>());
    if (!index_ || nq == 0) return;
    std::vector<std::int64_t> offs(nq + 1, 0);
    index_t* pi = nullptr;
    float* pd = nullptr;
    const Queries qs = queries(qp, nq);
    if (pclb200_radius(b200::Context::get(), index_->h, qs.ptr, nq, qs.stride, radius, max_nn, sorted_results_ ? 1 : 0,
                       offs.data(), &pi, &pd) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::search::KdTree::radiusSearch] %s\n", pclb200_last_error());
      return;
    }
    for (std::size_t i = 0; i < nq; ++i) {
      k_indices[i].assign(pi + offs[i], pi + offs[i + 1]);
      k_sqr_distances[i].assign(pd + offs[i], pd + offs[i + 1]);
    }
    pclb200_free(pi);
    pclb200_free(pd);
  }

protected:
  template <typename PointTDiff>
  static void copyXYZ(const pcl::PointCloud<PointTDiff>& cloud, const Indices& indices, PointCloud& pc)
  {
    const std::size_t n = indices.empty() ? cloud.size() : indices.size();
    pc.points.resize(n);
    for (std::size_t i = 0; i < n; ++i) {
      const PointTDiff& p = cloud[indices.empty() ? i : static_cast<std::size_t>(indices[i])];
      PointT q;
      q.x = p.x; q.y = p.y; q.z = p.z;
      pc.points[i] = q;
    }
    pc.width = static_cast<std::uint32_t>(n);
    pc.height = 1;
  }
  // queries go through the same representation as the indexed points
  struct Queries {
    const void* ptr;
    std::size_t stride;
    std::vector<float> store;
  };
  Queries queries(const PointT* q, std::size_t nq) const
  {
    Queries r{q, sizeof(PointT), {}};
    if (vectorized_) {
      r.store.assign(3 * nq, 0.f);
      for (std::size_t i = 0; i < nq; ++i) point_representation_->vectorize(q[i], &r.store[3 * i]);
      r.ptr = r.store.data();
      r.stride = 12;
    }
    return r;
  }

  int knn(const PointT* q, std::size_t nq, int k, Indices* ki, std::vector<float>* kd) const
  {
    ki->clear();
    kd->clear();
    if (!index_ || k <= 0) return 0;
    const Queries qs = queries(q, nq);
    std::vector<index_t> oi(nq * static_cast<std::size_t>(k));
    std::vector<float> od(nq * static_cast<std::size_t>(k));
    int keff = 0;
    if (pclb200_knn(b200::Context::get(), index_->h, qs.ptr, nq, qs.stride, k, oi.data(), od.data(), &keff) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::search::KdTree::nearestKSearch] %s\n", pclb200_last_error());
      return 0;
    }
    ki->assign(oi.begin(), oi.begin() + keff);  // resized to the clamped k (kdtree_flann.hpp:241-245)
    kd->assign(od.begin(), od.begin() + keff);
    return keff;
  }

  PointCloudConstPtr input_;
  IndicesConstPtr indices_;
  PointRepresentationConstPtr point_representation_;
  bool vectorized_ = false;  // the index holds representation vectors, not raw xyz
  std::shared_ptr<b200::IndexHandle> index_;
  bool sorted_results_ = true;
  float epsilon_ = 0.f;
  int min_pts_ = 1;
  std::string name_ = "KdTree";
};

// The consumers (NormalEstimation, the outlier filters, EuclideanClusterExtraction) take a pcl::search::Search<PointT>::Ptr
// in the reference.  Here they run on the device index of a pcl::search::KdTree; another Search subclass has no device
// side, and there is no CPU path to fall back to: it is refused with a message (nullptr).
template <typename PointT>
inline typename KdTree<PointT>::Ptr deviceSearcher(const typename Search<PointT>::Ptr& searcher, const char* who)
{
  if (!searcher) return typename KdTree<PointT>::Ptr();
  typename KdTree<PointT>::Ptr tree = std::dynamic_pointer_cast<KdTree<PointT>>(searcher);
  if (!tree)
    std::fprintf(stderr, "[%s::setSearchMethod] '%s' is not a pcl::search::KdTree: only the device searcher is supported\n", who,
                 searcher->getName().c_str());
  return tree;
}

}  // namespace search

// pcl::KdTreeFLANN<PointT> call sites compile unchanged against the same backend
template <typename PointT>
using KdTreeFLANN = search::KdTree<PointT>;
}  // namespace pcl
