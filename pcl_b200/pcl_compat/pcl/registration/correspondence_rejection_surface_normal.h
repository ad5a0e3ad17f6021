// pcl/registration/correspondence_rejection_surface_normal.h — CorrespondenceRejectorSurfaceNormal on the device.
// Reference: registration/include/pcl/registration/correspondence_rejection_surface_normal.h:56-346,
// registration/src/correspondence_rejection_surface_normal.cpp:43-66 and the DataContainer it scores with
// (correspondence_rejection.h:233-417).  Inside IterativeClosestPoint the rejector runs in the fused device loop on
// the rotated source normals and the target normals of the two PointNormal clouds (icp.hpp:166-201).
#pragma once
#include <vector>

#include "../PCLPointCloud2.h"
#include "../point_cloud.h"
#include "../point_types.h"
#include "correspondence_rejection.h"

namespace pcl {
namespace registration {

class CorrespondenceRejectorSurfaceNormal : public CorrespondenceRejector {
public:
  using Ptr = std::shared_ptr<CorrespondenceRejectorSurfaceNormal>;
  using ConstPtr = std::shared_ptr<const CorrespondenceRejectorSurfaceNormal>;
  CorrespondenceRejectorSurfaceNormal() { rejection_name_ = "CorrespondenceRejectorSurfaceNormal"; }

  void setThreshold(double threshold) { threshold_ = threshold; }  // cosine of the largest accepted angle
  double getThreshold() const { return threshold_; }

  // the reference keeps these inside a type-erased DataContainer; here only the normals are ever read
  template <typename PointT, typename NormalT> void initializeDataContainer() { initialized_ = true; }
  template <typename PointT> void setInputSource(const typename pcl::PointCloud<PointT>::ConstPtr&) {}
  template <typename PointT> void setInputTarget(const typename pcl::PointCloud<PointT>::ConstPtr&) {}
  template <typename PointT> void setInputCloud(const typename pcl::PointCloud<PointT>::ConstPtr&) {}
  template <typename PointT, typename NormalT> void setInputNormals(const typename pcl::PointCloud<NormalT>::ConstPtr& normals)
  {
    copyNormals<NormalT>(*normals, source_normals_);
  }
  template <typename PointT, typename NormalT> void setTargetNormals(const typename pcl::PointCloud<NormalT>::ConstPtr& normals)
  {
    copyNormals<NormalT>(*normals, target_normals_);
  }
  bool requiresSourceNormals() const override { return true; }
  bool requiresTargetNormals() const override { return true; }
  // correspondence_rejection_surface_normal.h:236-276: the blob route IterativeClosestPoint uses (impl/icp.hpp:142-155) —
  // the normal fields are looked up by name, the data container is set up on first use
  void setSourceNormals(pcl::PCLPointCloud2::ConstPtr cloud2) override
  {
    pcl::PointCloud<pcl::Normal> cloud;
    fromPCLPointCloud2(*cloud2, cloud);
    copyNormals<pcl::Normal>(cloud, source_normals_);
    initialized_ = true;
  }
  void setTargetNormals(pcl::PCLPointCloud2::ConstPtr cloud2) override
  {
    pcl::PointCloud<pcl::Normal> cloud;
    fromPCLPointCloud2(*cloud2, cloud);
    copyNormals<pcl::Normal>(cloud, target_normals_);
    initialized_ = true;
  }
  bool requiresSourcePoints() const override { return true; }   // :204-234: the container also takes the clouds
  bool requiresTargetPoints() const override { return true; }
  void setSourcePoints(pcl::PCLPointCloud2::ConstPtr /*cloud2*/) override {}   // only the normals are ever read
  void setTargetPoints(pcl::PCLPointCloud2::ConstPtr /*cloud2*/) override {}

  void getRemainingCorrespondences(const pcl::Correspondences& in, pcl::Correspondences& out) override
  {
    out.clear();
    if (!initialized_ || source_normals_.empty() || target_normals_.empty()) {  // .cpp:49-54
      std::fprintf(stderr, "[pcl::registration::%s::getRemainingCorrespondences] DataContainer object is not initialized!\n",
                   getClassName().c_str());
      return;
    }
    out.resize(in.size());
    std::size_t n = 0;
    if (pclb200_reject_surface_normal(b200::Context::get(), reinterpret_cast<const pclb200_corr*>(in.data()), in.size(),
                                      source_normals_.data(), source_normals_.size() / 4, 16, target_normals_.data(),
                                      target_normals_.size() / 4, 16, threshold_, reinterpret_cast<pclb200_corr*>(out.data()),
                                      &n) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::registration::%s::getRemainingCorrespondences] %s\n", getClassName().c_str(), pclb200_last_error());
      n = 0;
    }
    out.resize(n);
  }
  pclb200_rejector abiRejector() const override { return pclb200_rejector{PCLB200_REJ_SURFACE_NORMAL, 0, threshold_}; }

protected:
  template <typename NormalT> static void copyNormals(const pcl::PointCloud<NormalT>& cloud, std::vector<float>& dst)
  {
    dst.resize(4 * cloud.size());
    for (std::size_t i = 0; i < cloud.size(); ++i) {
      dst[4 * i] = cloud[i].normal_x;
      dst[4 * i + 1] = cloud[i].normal_y;
      dst[4 * i + 2] = cloud[i].normal_z;
      dst[4 * i + 3] = 0.f;
    }
  }
  double threshold_ = 1.0;  // correspondence_rejection_surface_normal.h:345
  bool initialized_ = false;
  std::vector<float> source_normals_, target_normals_;
};

}  // namespace registration
}  // namespace pcl
