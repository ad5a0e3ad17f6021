// pcl/registration/correspondence_estimation.h — pcl::registration::CorrespondenceEstimation on the device.
// Reference: registration/include/pcl/registration/correspondence_estimation.h:59-504 and
// impl/correspondence_estimation.hpp:52-311.
#pragma once
#include "../PCLPointCloud2.h"
#include <cmath>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#include "../correspondence.h"
#include "../search/kdtree.h"

namespace pcl {
namespace registration {

// correspondence_estimation.h:59-384 — state and setters every estimator shares; the two determine* methods are the
// virtual interface Registration calls (and what a drop-in subclass overrides).
template <typename PointSource, typename PointTarget, typename Scalar = float>
class CorrespondenceEstimationBase : public PCLBase<PointSource> {
public:
  using Ptr = std::shared_ptr<CorrespondenceEstimationBase>;
  using ConstPtr = std::shared_ptr<const CorrespondenceEstimationBase>;
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  using KdTreeReciprocal = pcl::search::KdTree<PointSource>;
  using KdTreeReciprocalPtr = typename KdTreeReciprocal::Ptr;
  using PointCloudSourceConstPtr = typename pcl::PointCloud<PointSource>::ConstPtr;
  using PointCloudTargetConstPtr = typename pcl::PointCloud<PointTarget>::ConstPtr;

  CorrespondenceEstimationBase() : tree_(new KdTree), tree_reciprocal_(new KdTreeReciprocal) {}

  void setInputSource(const PointCloudSourceConstPtr& cloud)
  {
    source_cloud_updated_ = true;
    PCLBase<PointSource>::setInputCloud(cloud);
  }
  PointCloudSourceConstPtr const getInputSource() { return this->input_; }
  void setInputTarget(const PointCloudTargetConstPtr& cloud)
  {
    if (!cloud || cloud->empty()) {
      std::fprintf(stderr, "[pcl::registration::CorrespondenceEstimation::setInputTarget] Invalid or empty point cloud dataset given!\n");
      return;
    }
    target_ = cloud;
    target_cloud_updated_ = true;
  }
  PointCloudTargetConstPtr const getInputTarget() { return target_; }
  void setIndicesSource(const IndicesPtr& indices) { this->setIndices(indices); }
  IndicesPtr const getIndicesSource() { return this->indices_; }                  // correspondence_estimation.h:200-204
  void setIndicesTarget(const IndicesPtr& indices) { target_cloud_updated_ = true; target_indices_ = indices; }
  IndicesPtr const getIndicesTarget() { return target_indices_; }                 // :217-221
  // correspondence_estimation.h:296-318: the representation under which the target (and, for the reciprocal search, the
  // source) tree compares points; handed to the tree when it is (re)built (impl/correspondence_estimation.hpp:66-67,102-103)
  using PointRepresentationConstPtr = typename KdTree::PointRepresentationConstPtr;
  using PointRepresentationReciprocalConstPtr = typename KdTreeReciprocal::PointRepresentationConstPtr;
  void setPointRepresentation(const PointRepresentationConstPtr& point_representation)
  {
    point_representation_ = point_representation;
    target_cloud_updated_ = true;
  }
  void setPointRepresentationReciprocal(const PointRepresentationReciprocalConstPtr& point_representation_reciprocal)
  {
    point_representation_reciprocal_ = point_representation_reciprocal;
    source_cloud_updated_ = true;
  }
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false)
  {
    tree_ = tree;
    force_no_recompute_ = force_no_recompute;
    target_cloud_updated_ = true;
  }
  KdTreePtr getSearchMethodTarget() const { return tree_; }
  void setSearchMethodSource(const KdTreeReciprocalPtr& tree, bool force_no_recompute = false)
  {
    tree_reciprocal_ = tree;
    force_no_recompute_reciprocal_ = force_no_recompute;
    source_cloud_updated_ = true;
  }
  KdTreeReciprocalPtr getSearchMethodSource() const { return tree_reciprocal_; }
  void setNumberOfThreads(unsigned int) {}  // one device launch replaces the OpenMP loop (:163-165)
  virtual Ptr clone() const = 0;            // correspondence_estimation.h:315
  virtual bool requiresSourceNormals() const { return false; }
  virtual bool requiresTargetNormals() const { return false; }
  // correspondence_estimation.h:277-322: the type-erased route by which IterativeClosestPoint hands an estimator the
  // normals it requires (impl/icp.hpp:142,169); estimators that need none ignore the blob
  virtual void setSourceNormals(pcl::PCLPointCloud2::ConstPtr /*cloud2*/) {}
  virtual void setTargetNormals(pcl::PCLPointCloud2::ConstPtr /*cloud2*/) {}
  virtual void determineCorrespondences(pcl::Correspondences& correspondences,
                                        double max_distance = std::numeric_limits<double>::max()) = 0;
  virtual void determineReciprocalCorrespondences(pcl::Correspondences& correspondences,
                                                  double max_distance = std::numeric_limits<double>::max()) = 0;
  // which estimator IterativeClosestPoint's fused device loop runs for this object (PCLB200_CORR_*) and its k
  virtual int abiKind() const { return PCLB200_CORR_NEAREST; }
  virtual int abiK() const { return 1; }

protected:
  const std::string& getClassName() const { return corr_name_; }                  // correspondence_estimation.h:353-357
  std::string corr_name_ = "CorrespondenceEstimationBase";
  PointRepresentationConstPtr point_representation_;
  PointRepresentationReciprocalConstPtr point_representation_reciprocal_;
  bool initCompute()
  {
    if (!target_) {
      std::fprintf(stderr, "[pcl::registration::CorrespondenceEstimation::compute] No input target dataset was given!\n");
      return false;
    }
    if (target_cloud_updated_ && !force_no_recompute_) {  // :63-75
      if (point_representation_) tree_->setPointRepresentation(point_representation_);
      tree_->setInputCloud(target_, target_indices_);
      target_cloud_updated_ = false;
    }
    return PCLBase<PointSource>::initCompute();
  }
  bool initComputeReciprocal()
  {
    if (source_cloud_updated_ && !force_no_recompute_reciprocal_) {  // :98-110
      if (point_representation_reciprocal_) tree_reciprocal_->setPointRepresentation(point_representation_reciprocal_);
      tree_reciprocal_->setInputCloud(this->input_, this->use_indices_ && !this->fake_indices_ ? IndicesConstPtr(this->indices_) : IndicesConstPtr());
      source_cloud_updated_ = false;
    }
    return true;
  }
  KdTreePtr tree_;
  KdTreeReciprocalPtr tree_reciprocal_;
  PointCloudTargetConstPtr target_;
  IndicesPtr target_indices_;
  bool target_cloud_updated_ = true, source_cloud_updated_ = true;
  bool force_no_recompute_ = false, force_no_recompute_reciprocal_ = false;
};

// correspondence_estimation.h:407-504 — nearest neighbour (+ reciprocal) estimator
template <typename PointSource, typename PointTarget, typename Scalar = float>
class CorrespondenceEstimation : public CorrespondenceEstimationBase<PointSource, PointTarget, Scalar> {
public:
  using Base = CorrespondenceEstimationBase<PointSource, PointTarget, Scalar>;
  using Ptr = std::shared_ptr<CorrespondenceEstimation>;
  using ConstPtr = std::shared_ptr<const CorrespondenceEstimation>;
  CorrespondenceEstimation() { this->corr_name_ = "CorrespondenceEstimation"; }
  typename Base::Ptr clone() const override { return typename Base::Ptr(new CorrespondenceEstimation(*this)); }  // :493-498
  // impl/correspondence_estimation.hpp:145-218
  void determineCorrespondences(pcl::Correspondences& correspondences,
                                double max_distance = std::numeric_limits<double>::max()) override
  {
    run(correspondences, max_distance, false);
  }
  // impl/correspondence_estimation.hpp:220-311
  void determineReciprocalCorrespondences(pcl::Correspondences& correspondences,
                                          double max_distance = std::numeric_limits<double>::max()) override
  {
    run(correspondences, max_distance, true);
  }

protected:
  void run(pcl::Correspondences& out, double max_distance, bool reciprocal)
  {
    out.clear();
    if (!this->initCompute()) return;
    if (reciprocal && !this->initComputeReciprocal()) return;
    if (!this->tree_->deviceIndex() || (reciprocal && !this->tree_reciprocal_->deviceIndex())) return;
    if (this->tree_->usesRepresentationVectors() || (reciprocal && this->tree_reciprocal_->usesRepresentationVectors())) {
      runThroughTrees(out, max_distance, reciprocal);
      return;
    }
    out.resize(this->indices_->size());
    std::size_t n_out = 0;
    // max_distance = DBL_MAX squares to +inf (no gate), exactly like `max_distance * max_distance` at :161
    int rc = pclb200_correspondences(b200::Context::get(), this->tree_->deviceIndex(),
                                     reciprocal ? this->tree_reciprocal_->deviceIndex() : nullptr, this->input_->points.data(),
                                     this->input_->size(), sizeof(PointSource), this->abiIndices(), this->abiIndexCount(),
                                     this->input_->is_dense ? 1 : 0, max_distance,
                                     reinterpret_cast<pclb200_corr*>(out.data()), &n_out);
    if (rc != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::registration::CorrespondenceEstimation] %s\n", pclb200_last_error());
      n_out = 0;
    }
    out.resize(n_out);
  }

  // With a rescaling / lower-dimensional PointRepresentation the trees compare representation vectors, so the queries
  // go through the trees' own batch searches (which vectorise them the same way) and the pairing rules of
  // impl/correspondence_estimation.hpp:167-215 / :247-306 are applied on the host.  Distances are those of the
  // representation space, as in the reference.
  void runThroughTrees(pcl::Correspondences& out, double max_distance, bool reciprocal)
  {
    const double max_d2 = max_distance * max_distance;
    const Indices& idx = *this->indices_;
    pcl::PointCloud<PointTarget> q;
    q.points.reserve(idx.size());
    std::vector<index_t> kept;
    kept.reserve(idx.size());
    for (index_t i : idx) {
      const PointSource& p = (*this->input_)[i];
      if (!this->input_->is_dense && !(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
      PointTarget t;
      t.x = p.x; t.y = p.y; t.z = p.z;
      q.points.push_back(t);
      kept.push_back(i);
    }
    std::vector<Indices> ki;
    std::vector<std::vector<float>> kd;
    this->tree_->nearestKSearch(q, Indices(), 1, ki, kd);
    std::vector<index_t> cand_src, cand_tgt;
    std::vector<float> cand_d;
    for (std::size_t j = 0; j < kept.size(); ++j) {
      if (ki[j].empty() || static_cast<double>(kd[j][0]) > max_d2) continue;
      cand_src.push_back(kept[j]);
      cand_tgt.push_back(ki[j][0]);
      cand_d.push_back(kd[j][0]);
    }
    if (reciprocal) {  // the matched target point must find this source point back (:259-269)
      pcl::PointCloud<PointSource> back;
      back.points.reserve(cand_tgt.size());
      for (index_t t : cand_tgt) {
        PointSource sp;
        sp.x = (*this->target_)[t].x; sp.y = (*this->target_)[t].y; sp.z = (*this->target_)[t].z;
        back.points.push_back(sp);
      }
      std::vector<Indices> bi;
      std::vector<std::vector<float>> bd;
      this->tree_reciprocal_->nearestKSearch(back, Indices(), 1, bi, bd);
      for (std::size_t j = 0; j < cand_src.size(); ++j)
        if (!bi[j].empty() && static_cast<double>(bd[j][0]) <= max_d2 && bi[j][0] == cand_src[j])
          out.emplace_back(cand_src[j], cand_tgt[j], cand_d[j]);
    }
    else
      for (std::size_t j = 0; j < cand_src.size(); ++j) out.emplace_back(cand_src[j], cand_tgt[j], cand_d[j]);
  }
};

}  // namespace registration
}  // namespace pcl
