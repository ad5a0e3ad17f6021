// pcl/registration/transformation_estimation*.h — TransformationEstimation, ...SVD, ...PointToPlaneLLS on the device.
// Reference: registration/include/pcl/registration/transformation_estimation.h:61-121,
// impl/transformation_estimation_svd.hpp:50-225 (Umeyama path), impl/transformation_estimation_point_to_plane_lls.hpp:50-310.
#pragma once
#include <cstdio>
#include <memory>
#include <vector>

#include "../b200/context.h"
#include "../correspondence.h"
#include "../eigen_lite.h"
#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {
namespace registration {

template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationEstimation {
public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using Ptr = std::shared_ptr<TransformationEstimation>;
  using ConstPtr = std::shared_ptr<const TransformationEstimation>;
  virtual ~TransformationEstimation() = default;
  virtual void estimateRigidTransformation(const pcl::PointCloud<PointSource>& cloud_src,
                                           const pcl::PointCloud<PointTarget>& cloud_tgt, Matrix4& T) const = 0;
  virtual void estimateRigidTransformation(const pcl::PointCloud<PointSource>& cloud_src, const pcl::Indices& indices_src,
                                           const pcl::PointCloud<PointTarget>& cloud_tgt, Matrix4& T) const = 0;
  virtual void estimateRigidTransformation(const pcl::PointCloud<PointSource>& cloud_src, const pcl::Indices& indices_src,
                                           const pcl::PointCloud<PointTarget>& cloud_tgt, const pcl::Indices& indices_tgt,
                                           Matrix4& T) const = 0;
  virtual void estimateRigidTransformation(const pcl::PointCloud<PointSource>& cloud_src,
                                           const pcl::PointCloud<PointTarget>& cloud_tgt,
                                           const pcl::Correspondences& correspondences, Matrix4& T) const = 0;
  virtual int abiEstimator() const = 0;  // PCLB200_EST_* (lets ICP run the fused device loop for the stock estimators)

protected:
  static void fromRowMajor(const double* t, Matrix4& T)
  {
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) T(r, c) = static_cast<Scalar>(t[4 * r + c]);
  }
  static std::vector<pclb200_corr> pairs(const pcl::Indices& a, const pcl::Indices& b)
  {
    std::vector<pclb200_corr> c(a.size());
    for (std::size_t i = 0; i < a.size(); ++i) c[i] = pclb200_corr{a[i], b[i], 0.f};
    return c;
  }
};

template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationEstimationSVD : public TransformationEstimation<PointSource, PointTarget, Scalar> {
public:
  using Base = TransformationEstimation<PointSource, PointTarget, Scalar>;
  using Matrix4 = typename Base::Matrix4;
  using Ptr = std::shared_ptr<TransformationEstimationSVD>;
  explicit TransformationEstimationSVD(bool use_umeyama = true) : use_umeyama_(use_umeyama) {}
  int abiEstimator() const override { return PCLB200_EST_SVD; }
  bool usesUmeyama() const { return use_umeyama_; }

  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, Matrix4& T) const override
  {
    if (s.size() != t.size()) {  // impl/transformation_estimation_svd.hpp:58-66
      std::fprintf(stderr, "[pcl::TransformationEstimationSVD::estimateRigidTransformation] Number or points in source (%zu) differs than target (%zu)!\n", s.size(), t.size());
      return;
    }
    solve(s, t, nullptr, s.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::Indices& is, const pcl::PointCloud<PointTarget>& t, Matrix4& T) const override
  {
    if (is.size() != t.size()) {
      std::fprintf(stderr, "[pcl::TransformationSVD::estimateRigidTransformation] Number or points in source (%zu) differs than target (%zu)!\n", is.size(), t.size());
      return;
    }
    pcl::Indices it(t.size());
    for (std::size_t i = 0; i < it.size(); ++i) it[i] = static_cast<index_t>(i);
    auto c = Base::pairs(is, it);
    solve(s, t, c.data(), c.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::Indices& is, const pcl::PointCloud<PointTarget>& t, const pcl::Indices& it, Matrix4& T) const override
  {
    if (is.size() != it.size()) {
      std::fprintf(stderr, "[pcl::TransformationEstimationSVD::estimateRigidTransformation] Number or points in source (%zu) differs than target (%zu)!\n", is.size(), it.size());
      return;
    }
    auto c = Base::pairs(is, it);
    solve(s, t, c.data(), c.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, const pcl::Correspondences& corr, Matrix4& T) const override
  {
    solve(s, t, reinterpret_cast<const pclb200_corr*>(corr.data()), corr.size(), T);
  }

protected:
  void solve(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, const pclb200_corr* c, std::size_t n, Matrix4& T) const
  {
    double out[16];
    // use_umeyama_ = false selects getTransformationFromCorrelation (impl/transformation_estimation_svd.hpp:156-225)
    auto fn = use_umeyama_ ? pclb200_estimate_svd : pclb200_estimate_svd_correlation;
    if (n == 0 || fn(b200::Context::get(), s.points.data(), sizeof(PointSource), t.points.data(), sizeof(PointTarget), c, n,
                     sizeof(Scalar) == 8, out) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::TransformationEstimationSVD] %s\n", n ? pclb200_last_error() : "no point pairs");
      return;
    }
    Base::fromRowMajor(out, T);
  }
  bool use_umeyama_;
};

template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationEstimationPointToPlaneLLS : public TransformationEstimation<PointSource, PointTarget, Scalar> {
public:
  using Base = TransformationEstimation<PointSource, PointTarget, Scalar>;
  using Matrix4 = typename Base::Matrix4;
  using Ptr = std::shared_ptr<TransformationEstimationPointToPlaneLLS>;
  static_assert(has_normal<PointTarget>::value, "TransformationEstimationPointToPlaneLLS needs target normals");
  int abiEstimator() const override { return PCLB200_EST_POINT_TO_PLANE_LLS; }

  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, Matrix4& T) const override
  {
    if (s.size() != t.size()) {
      std::fprintf(stderr, "[pcl::TransformationEstimationPointToPlaneLLS::estimateRigidTransformation] Number or points in source (%zu) differs than target (%zu)!\n", s.size(), t.size());
      return;
    }
    solve(s, t, nullptr, s.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::Indices& is, const pcl::PointCloud<PointTarget>& t, Matrix4& T) const override
  {
    if (is.size() != t.size()) return;
    pcl::Indices it(t.size());
    for (std::size_t i = 0; i < it.size(); ++i) it[i] = static_cast<index_t>(i);
    auto c = Base::pairs(is, it);
    solve(s, t, c.data(), c.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::Indices& is, const pcl::PointCloud<PointTarget>& t, const pcl::Indices& it, Matrix4& T) const override
  {
    if (is.size() != it.size()) return;
    auto c = Base::pairs(is, it);
    solve(s, t, c.data(), c.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, const pcl::Correspondences& corr, Matrix4& T) const override
  {
    solve(s, t, reinterpret_cast<const pclb200_corr*>(corr.data()), corr.size(), T);
  }

protected:
  void solve(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, const pclb200_corr* c, std::size_t n, Matrix4& T) const
  {
    double out[16];
    const void* normals = t.empty() ? nullptr : static_cast<const void*>(&t.points[0].normal_x);
    if (n == 0 || pclb200_estimate_point_to_plane_lls(b200::Context::get(), s.points.data(), sizeof(PointSource), t.points.data(), normals,
                                                      sizeof(PointTarget), c, n, sizeof(Scalar) == 8, out) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::TransformationEstimationPointToPlaneLLS] %s\n", n ? pclb200_last_error() : "no point pairs");
      return;
    }
    Base::fromRowMajor(out, T);
  }
};

// impl/transformation_estimation_symmetric_point_to_plane_lls.hpp:50-215
template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationEstimationSymmetricPointToPlaneLLS : public TransformationEstimation<PointSource, PointTarget, Scalar> {
public:
  using Base = TransformationEstimation<PointSource, PointTarget, Scalar>;
  using Matrix4 = typename Base::Matrix4;
  using Ptr = std::shared_ptr<TransformationEstimationSymmetricPointToPlaneLLS>;
  static_assert(has_normal<PointSource>::value && has_normal<PointTarget>::value,
                "TransformationEstimationSymmetricPointToPlaneLLS needs source and target normals");
  int abiEstimator() const override { return PCLB200_EST_SYMMETRIC_POINT_TO_PLANE_LLS; }
  void setEnforceSameDirectionNormals(bool v) { enforce_same_direction_normals_ = v; }
  bool getEnforceSameDirectionNormals() const { return enforce_same_direction_normals_; }

  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, Matrix4& T) const override
  {
    if (s.size() != t.size()) {
      std::fprintf(stderr, "[pcl::TransformationEstimationSymmetricPointToPlaneLLS::estimateRigidTransformation] Number or points in source (%zu) differs than target (%zu)!\n", s.size(), t.size());
      return;
    }
    solve(s, t, nullptr, s.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::Indices& is, const pcl::PointCloud<PointTarget>& t, Matrix4& T) const override
  {
    if (is.size() != t.size()) return;
    pcl::Indices it(t.size());
    for (std::size_t i = 0; i < it.size(); ++i) it[i] = static_cast<index_t>(i);
    auto c = Base::pairs(is, it);
    solve(s, t, c.data(), c.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::Indices& is, const pcl::PointCloud<PointTarget>& t, const pcl::Indices& it, Matrix4& T) const override
  {
    if (is.size() != it.size()) return;
    auto c = Base::pairs(is, it);
    solve(s, t, c.data(), c.size(), T);
  }
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, const pcl::Correspondences& corr, Matrix4& T) const override
  {
    solve(s, t, reinterpret_cast<const pclb200_corr*>(corr.data()), corr.size(), T);
  }

protected:
  void solve(const pcl::PointCloud<PointSource>& s, const pcl::PointCloud<PointTarget>& t, const pclb200_corr* c, std::size_t n, Matrix4& T) const
  {
    double out[16];
    if (n == 0 || s.empty() || t.empty() ||
        pclb200_estimate_symmetric_point_to_plane_lls(b200::Context::get(), s.points.data(), &s.points[0].normal_x, sizeof(PointSource),
                                                      t.points.data(), &t.points[0].normal_x, sizeof(PointTarget), c, n,
                                                      enforce_same_direction_normals_ ? 1 : 0, sizeof(Scalar) == 8, out) != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::TransformationEstimationSymmetricPointToPlaneLLS] %s\n", n ? pclb200_last_error() : "no point pairs");
      return;
    }
    Base::fromRowMajor(out, T);
  }
  bool enforce_same_direction_normals_ = true;
};

}  // namespace registration
}  // namespace pcl
