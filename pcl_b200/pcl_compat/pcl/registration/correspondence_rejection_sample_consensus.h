// pcl/registration/correspondence_rejection_sample_consensus.h — CorrespondenceRejectorSampleConsensus<PointT>: RANSAC over a
// correspondence list with a rigid transform as the model, keeping the inliers of the best one.
// Reference: registration/include/pcl/registration/correspondence_rejection_sample_consensus.h:56-290,
// impl/correspondence_rejection_sample_consensus.hpp:52-135; the RANSAC loop it drives is
// sample_consensus/include/pcl/sample_consensus/impl/ransac.hpp:52-224 over SampleConsensusModelRegistration
// (sac_model_registration.h:209-257, impl/sac_model_registration.hpp:46-250).
//
// This one runs on the HOST: RANSAC is a sequential search whose trial count adapts to the best consensus so far, and the
// reference's own test for it (test/registration/test_registration.cpp:336-382) only pins tolerances.  An ICP that holds a
// rejector of this kind takes the stage-by-stage loop (Registration::computeTransformationStaged): correspondences come
// from the device searcher, this rejector filters them on the host, the estimator runs on the device again.
// The sample sequence is the reference's (mt19937 seeded 12345, draws halved like boost::uniform_int<>(0, INT_MAX), the same
// partial shuffle of a persistent index array); the 3-point model fit is a double-precision Kabsch solve instead of the float
// Umeyama, so a pair that sits on the inlier threshold may fall on the other side.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <random>
#include <unordered_map>
#include <vector>

#include "../common/centroid.h"
#include "../common/eigen.h"
#include "../conversions.h"
#include "../point_cloud.h"
#include "correspondence_rejection.h"

namespace pcl {
namespace registration {
namespace detail {
// cyclic Jacobi for a symmetric 3x3 (row-major): A = V diag(w) V^T
inline void jacobiEigenSym3(double A[9], double w[3], double V[9])
{
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[3 * p + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const double akp = A[3 * k + p], akq = A[3 * k + q];
          A[3 * k + p] = c * akp - s * akq;
          A[3 * k + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const double apk = A[3 * p + k], aqk = A[3 * q + k];
          A[3 * p + k] = c * apk - s * aqk;
          A[3 * q + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = c * vkp - s * vkq;
          V[3 * k + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[4 * i];
}
inline void cross3(const double a[3], const double b[3], double o[3])
{
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
// least-squares rigid transform q ~ R p + t over n >= 3 pairs (Kabsch: SVD of the cross-covariance through the eigenvectors
// of H^T H; the third axis of a planar configuration — three points always are one — is completed by cross products).
// T: row-major 4x4.  false if the pairs do not fix a rotation (two singular values vanish).
inline bool rigidFromPairs(const double* p, const double* q, std::size_t n, double T[16])
{
  double cp[3] = {0, 0, 0}, cq[3] = {0, 0, 0};
  for (std::size_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) { cp[d] += p[3 * i + d]; cq[d] += q[3 * i + d]; }
  for (int d = 0; d < 3; ++d) { cp[d] /= static_cast<double>(n); cq[d] /= static_cast<double>(n); }
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // sum (q - cq)(p - cp)^T
  for (std::size_t i = 0; i < n; ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) H[3 * r + c] += (q[3 * i + r] - cq[r]) * (p[3 * i + c] - cp[c]);
  double HtH[9], w[3], V[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) HtH[3 * r + c] = H[r] * H[c] + H[3 + r] * H[3 + c] + H[6 + r] * H[6 + c];
  jacobiEigenSym3(HtH, w, V);
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int a, int b) { return w[a] > w[b]; });
  double v[3][3], u[3][3];
  for (int k = 0; k < 3; ++k)
    for (int d = 0; d < 3; ++d) v[k][d] = V[3 * d + order[k]];
  const double s0 = std::sqrt(std::max(w[order[0]], 0.0)), s1 = std::sqrt(std::max(w[order[1]], 0.0));
  if (!(s1 > 1e-12 * std::max(s0, 1e-300))) return false;
  for (int k = 0; k < 2; ++k) {
    double nrm = 0;
    for (int r = 0; r < 3; ++r) { u[k][r] = H[3 * r] * v[k][0] + H[3 * r + 1] * v[k][1] + H[3 * r + 2] * v[k][2]; nrm += u[k][r] * u[k][r]; }
    nrm = std::sqrt(nrm);
    for (int r = 0; r < 3; ++r) u[k][r] /= nrm;
  }
  {  // u1 orthogonal to u0 (they are, up to round-off), then right-handed third axes on both sides: det(U) = det(V) = +1
    double dot = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2], nrm = 0;
    for (int r = 0; r < 3; ++r) { u[1][r] -= dot * u[0][r]; nrm += u[1][r] * u[1][r]; }
    nrm = std::sqrt(nrm);
    for (int r = 0; r < 3; ++r) u[1][r] /= nrm;
    cross3(v[0], v[1], v[2]);
    cross3(u[0], u[1], u[2]);
  }
  // with both frames right-handed R = U V^T is a rotation; it is the least-squares one unless the third singular value is
  // the one that would have to flip (a reflection fits better), which three non-collinear point pairs never ask for
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = u[0][r] * v[0][c] + u[1][r] * v[1][c] + u[2][r] * v[2][c];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[4 * r + c] = R[3 * r + c];
    T[4 * r + 3] = cq[r] - (R[3 * r] * cp[0] + R[3 * r + 1] * cp[1] + R[3 * r + 2] * cp[2]);
  }
  T[12] = T[13] = T[14] = 0.0;
  T[15] = 1.0;
  return true;
}
}  // namespace detail

template <typename PointT>
class CorrespondenceRejectorSampleConsensus : public CorrespondenceRejector {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  using PointCloudPtr = typename PointCloud::Ptr;
  using PointCloudConstPtr = typename PointCloud::ConstPtr;
  using Ptr = std::shared_ptr<CorrespondenceRejectorSampleConsensus<PointT>>;
  using ConstPtr = std::shared_ptr<const CorrespondenceRejectorSampleConsensus<PointT>>;

  CorrespondenceRejectorSampleConsensus()
  {
    rejection_name_ = "CorrespondenceRejectorSampleConsensus";
  }

  virtual void setInputSource(const PointCloudConstPtr& cloud) { input_ = cloud; }
  PointCloudConstPtr const getInputSource() { return input_; }
  virtual void setInputTarget(const PointCloudConstPtr& cloud) { target_ = cloud; }
  PointCloudConstPtr const getInputTarget() { return target_; }
  bool requiresSourcePoints() const override { return true; }
  void setSourcePoints(pcl::PCLPointCloud2::ConstPtr cloud2) override
  {
    PointCloudPtr cloud(new PointCloud);
    fromPCLPointCloud2(*cloud2, *cloud);
    setInputSource(cloud);
  }
  bool requiresTargetPoints() const override { return true; }
  void setTargetPoints(pcl::PCLPointCloud2::ConstPtr cloud2) override
  {
    PointCloudPtr cloud(new PointCloud);
    fromPCLPointCloud2(*cloud2, *cloud);
    setInputTarget(cloud);
  }
  void setInlierThreshold(double threshold) { inlier_threshold_ = threshold; }
  double getInlierThreshold() { return inlier_threshold_; }
  void setMaximumIterations(int max_iterations) { max_iterations_ = std::max(max_iterations, 0); }
  int getMaximumIterations() { return max_iterations_; }
  Eigen::Matrix4f getBestTransformation() { return best_transformation_; }
  void setRefineModel(const bool refine) { refine_ = refine; }
  bool getRefineModel() const { return refine_; }
  void getInliersIndices(pcl::Indices& inlier_indices) { inlier_indices = inlier_indices_; }
  void setSaveInliers(bool s) { save_inliers_ = s; }
  bool getSaveInliers() { return save_inliers_; }

  bool runsOnDevice() const override { return false; }
  pclb200_rejector abiRejector() const override { return pclb200_rejector{-1, 0, 0.0}; }  // no device form

  // impl/correspondence_rejection_sample_consensus.hpp:52-135
  void getRemainingCorrespondences(const pcl::Correspondences& original, pcl::Correspondences& remaining) override
  {
    if (!input_) {
      std::fprintf(stderr, "[pcl::registration::%s::getRemainingCorrespondences] No input cloud dataset was given!\n", getClassName().c_str());
      return;
    }
    if (!target_) {
      std::fprintf(stderr, "[pcl::registration::%s::getRemainingCorrespondences] No input target dataset was given!\n", getClassName().c_str());
      return;
    }
    if (save_inliers_) inlier_indices_.clear();
    const std::size_t n = original.size();
    std::vector<double> P(3 * n), Q(3 * n);
    pcl::Indices source_indices(n);
    for (std::size_t i = 0; i < n; ++i) {
      const PointT& s = (*input_)[original[i].index_query];
      const PointT& t = (*target_)[original[i].index_match];
      P[3 * i] = s.x; P[3 * i + 1] = s.y; P[3 * i + 2] = s.z;
      Q[3 * i] = t.x; Q[3 * i + 1] = t.y; Q[3 * i + 2] = t.z;
      source_indices[i] = original[i].index_query;
    }
    double best[16];
    std::vector<std::size_t> inliers;
    if (!ransac(P, Q, source_indices, best, inliers) || inliers.size() < 3) {  // :98-103, :114-118
      remaining = original;
      setIdentity(best_transformation_);
      return;
    }
    // :119-127: an inlier is a SOURCE POINT; the correspondence kept for it is the last one that names it
    std::unordered_map<int, int> index_to_correspondence;
    for (std::size_t i = 0; i < n; ++i) index_to_correspondence[original[i].index_query] = static_cast<int>(i);
    remaining.resize(inliers.size());
    for (std::size_t i = 0; i < inliers.size(); ++i) remaining[i] = original[index_to_correspondence[source_indices[inliers[i]]]];
    if (save_inliers_) {
      inlier_indices_.reserve(inliers.size());
      for (std::size_t i : inliers) inlier_indices_.push_back(index_to_correspondence[source_indices[i]]);
    }
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) best_transformation_(r, c) = static_cast<float>(best[4 * r + c]);
  }

protected:
  static void setIdentity(Eigen::Matrix4f& m)
  {
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) m(r, c) = r == c ? 1.f : 0.f;
  }
  // squared distance of pair i under T, in float like the reference's Vector4f arithmetic (sac_model_registration.hpp:235-244)
  static float pairDistance2(const double T[16], const double* p, const double* q)
  {
    float d2 = 0.f;
    for (int r = 0; r < 3; ++r) {
      const float tr = static_cast<float>(T[4 * r]) * static_cast<float>(p[0]) + static_cast<float>(T[4 * r + 1]) * static_cast<float>(p[1]) +
                       static_cast<float>(T[4 * r + 2]) * static_cast<float>(p[2]) + static_cast<float>(T[4 * r + 3]);
      const float e = tr - static_cast<float>(q[r]);
      d2 += e * e;
    }
    return d2;
  }
  // sac_model_registration.h:209-257: (mean of the square roots of the eigenvalues of the source covariance)^2
  double sampleDistanceThreshold(const pcl::Indices& source_indices) const
  {
    Eigen::Matrix3f cov;
    Eigen::Vector4f centroid;
    if (computeMeanAndCovarianceMatrix(*input_, source_indices, cov, centroid) == 0) return 0.0;
    float scale = 0.f;
    for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(cov[i]));
    if (!(scale > std::numeric_limits<float>::min())) return 0.0;
    Eigen::Matrix3f scaled = cov;
    for (int i = 0; i < 9; ++i) scaled[i] /= scale;
    float roots[3];
    computeRoots(scaled, roots);
    double sum = 0.0;
    for (int i = 0; i < 3; ++i) sum += std::sqrt(std::max(roots[i] * scale, 0.f));
    const double t = sum / 3.0;
    return t * t;
  }
  // ransac.hpp:52-224 with SampleConsensusModelRegistration as the model (sample size 3, probability 0.99)
  bool ransac(const std::vector<double>& P, const std::vector<double>& Q, const pcl::Indices& source_indices, double best[16],
              std::vector<std::size_t>& inliers) const
  {
    const std::size_t n = source_indices.size();
    inliers.clear();
    if (n < 3) return false;
    const double sample_dist_thresh = sampleDistanceThreshold(source_indices);
    const float thresh = static_cast<float>(inlier_threshold_ * inlier_threshold_);
    std::mt19937 rnd(12345u);  // the reference's model is built with random = false: one fixed seed per call
    std::vector<std::size_t> shuffled(n);
    for (std::size_t i = 0; i < n; ++i) shuffled[i] = i;
    const double log_probability = std::log(1.0 - 0.99);
    double k = std::numeric_limits<double>::max();
    std::size_t n_best = 0;
    int iterations = 0;
    unsigned skipped = 0;
    const unsigned max_skip = static_cast<unsigned>(max_iterations_) * 10u;
    bool have_model = false;
    auto far_apart = [&](std::size_t a, std::size_t b) {
      const double dx = P[3 * a] - P[3 * b], dy = P[3 * a + 1] - P[3 * b + 1], dz = P[3 * a + 2] - P[3 * b + 2];
      return dx * dx + dy * dy + dz * dz > sample_dist_thresh;
    };
    while (true) {
      std::size_t s[3];
      bool good = false;
      for (unsigned check = 0; check < 1000u && !good; ++check) {  // sac_model.h:162-197, max_sample_checks_
        // sac_model.h:466-476 drawIndexSample; rnd() there is boost::uniform_int<>(0, INT_MAX) over mt19937 = the raw draw halved
        for (std::size_t i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + (rnd() >> 1) % (n - i)]);
        s[0] = shuffled[0]; s[1] = shuffled[1]; s[2] = shuffled[2];
        good = far_apart(s[1], s[0]) && far_apart(s[2], s[0]) && far_apart(s[2], s[1]);
      }
      if (!good) break;  // "No samples could be selected"
      double p3[9], q3[9], T[16];
      for (int j = 0; j < 3; ++j)
        for (int d = 0; d < 3; ++d) { p3[3 * j + d] = P[3 * s[j] + d]; q3[3 * j + d] = Q[3 * s[j] + d]; }
      if (!detail::rigidFromPairs(p3, q3, 3, T)) {
        if (++skipped < max_skip) continue;
        break;
      }
      std::size_t count = 0;
      for (std::size_t i = 0; i < n; ++i) count += pairDistance2(T, &P[3 * i], &Q[3 * i]) < thresh;
      if (count > n_best) {
        n_best = count;
        have_model = true;
        for (int i = 0; i < 16; ++i) best[i] = T[i];
        const double w = static_cast<double>(n_best) / static_cast<double>(n);
        double p_outliers = 1.0 - std::pow(w, 3.0);
        p_outliers = std::max(std::numeric_limits<double>::epsilon(), p_outliers);
        p_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_outliers);
        k = log_probability / std::log(p_outliers);
      }
      ++iterations;
      if (static_cast<double>(iterations) > k) break;
      if (iterations > max_iterations_) break;
    }
    if (!have_model) return false;
    if (refine_) refine(P, Q, best);
    for (std::size_t i = 0; i < n; ++i)
      if (pairDistance2(best, &P[3 * i], &Q[3 * i]) < thresh) inliers.push_back(i);
    return true;
  }
  // sac.h refineModel (sigma = 3, at most 1000 rounds): refit on the inliers until the inlier set stops changing
  void refine(const std::vector<double>& P, const std::vector<double>& Q, double T[16]) const
  {
    const std::size_t n = P.size() / 3;
    const float thresh = static_cast<float>(inlier_threshold_ * inlier_threshold_);
    std::vector<std::size_t> prev;
    for (int round = 0; round < 1000; ++round) {
      std::vector<std::size_t> in;
      for (std::size_t i = 0; i < n; ++i)
        if (pairDistance2(T, &P[3 * i], &Q[3 * i]) < thresh) in.push_back(i);
      if (in.size() < 3 || in == prev) break;
      std::vector<double> p(3 * in.size()), q(3 * in.size());
      for (std::size_t j = 0; j < in.size(); ++j)
        for (int d = 0; d < 3; ++d) { p[3 * j + d] = P[3 * in[j] + d]; q[3 * j + d] = Q[3 * in[j] + d]; }
      double Tn[16];
      if (!detail::rigidFromPairs(p.data(), q.data(), in.size(), Tn)) break;
      for (int i = 0; i < 16; ++i) T[i] = Tn[i];
      prev.swap(in);
    }
  }

  double inlier_threshold_ = 0.05;
  int max_iterations_ = 1000;
  PointCloudConstPtr input_, target_;
  Eigen::Matrix4f best_transformation_;
  bool refine_ = false;
  pcl::Indices inlier_indices_;
  bool save_inliers_ = false;
};
}  // namespace registration
}  // namespace pcl
