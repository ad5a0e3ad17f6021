// pcl/common/eigen.h — the closed-form symmetric 3x3 eigen-solver NormalEstimation rests on
// (common/include/pcl/common/impl/eigen.hpp:52-133 computeRoots2 / computeRoots, :273-288 getLargest3x3Eigenvector,
// :293-326 eigen33: smallest eigenvalue and its eigenvector), host side.  The device kernels (search.cu: roots3_dev,
// normal_from_moments) evaluate the same expressions in the same order; this is the public function a PCL caller
// reaches directly.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <utility>

#include "../eigen_lite.h"

namespace pcl {
// roots of x^2 - b x + c = 0 (the cubic's degenerate case), ascending in roots[1], roots[2]; roots[0] = 0
template <typename Scalar, typename Roots>
inline void computeRoots2(const Scalar& b, const Scalar& c, Roots& roots)
{
  roots[0] = Scalar(0);
  Scalar d = Scalar(b * b - 4.0 * c);
  if (d < 0.0) d = 0.0;  // no real roots: the matrix is PSD, this is round-off
  const Scalar sd = std::sqrt(d);
  roots[2] = Scalar(0.5f) * (b + sd);
  roots[1] = Scalar(0.5f) * (b - sd);
}

// eigenvalues of a symmetric 3x3 matrix, ascending: trigonometric solution of the characteristic cubic with the
// clamps that keep round-off from producing NaNs
template <typename Matrix, typename Roots>
inline void computeRoots(const Matrix& m, Roots& roots)
{
  using Scalar = float;
  const Scalar c0 = m(0, 0) * m(1, 1) * m(2, 2) + Scalar(2) * m(0, 1) * m(0, 2) * m(1, 2) - m(0, 0) * m(1, 2) * m(1, 2) -
                    m(1, 1) * m(0, 2) * m(0, 2) - m(2, 2) * m(0, 1) * m(0, 1);
  const Scalar c1 = m(0, 0) * m(1, 1) - m(0, 1) * m(0, 1) + m(0, 0) * m(2, 2) - m(0, 2) * m(0, 2) + m(1, 1) * m(2, 2) -
                    m(1, 2) * m(1, 2);
  const Scalar c2 = m(0, 0) + m(1, 1) + m(2, 2);
  if (std::abs(c0) < std::numeric_limits<Scalar>::epsilon()) {  // one root is 0
    computeRoots2(c2, c1, roots);
    return;
  }
  const Scalar s_inv3 = Scalar(1.0 / 3.0);
  const Scalar s_sqrt3 = std::sqrt(Scalar(3.0));
  const Scalar c2_over_3 = c2 * s_inv3;
  Scalar a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > Scalar(0)) a_over_3 = Scalar(0);
  const Scalar half_b = Scalar(0.5) * (c0 + c2_over_3 * (Scalar(2) * c2_over_3 * c2_over_3 - c1));
  Scalar q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > Scalar(0)) q = Scalar(0);
  const Scalar rho = std::sqrt(-a_over_3);
  const Scalar theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  const Scalar cos_theta = std::cos(theta);
  const Scalar sin_theta = std::sin(theta);
  roots[0] = c2_over_3 + Scalar(2) * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0) computeRoots2(c2, c1, roots);  // a PSD matrix has no negative eigenvalue: fall back to the quadratic
}

namespace detail {
// the longest of the three pairwise cross products of the rows of a rank-2 matrix spans its null space
inline void largest3x3Eigenvector(const Eigen::Matrix3f& s, Eigen::Vector3f& v)
{
  float c[3][3];
  const int pairs[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  float len[3];
  for (int k = 0; k < 3; ++k) {
    const int a = pairs[k][0], b = pairs[k][1];
    c[k][0] = s(a, 1) * s(b, 2) - s(a, 2) * s(b, 1);
    c[k][1] = s(a, 2) * s(b, 0) - s(a, 0) * s(b, 2);
    c[k][2] = s(a, 0) * s(b, 1) - s(a, 1) * s(b, 0);
    len[k] = std::sqrt(c[k][0] * c[k][0] + c[k][1] * c[k][1] + c[k][2] * c[k][2]);
  }
  int best = 0;                      // first maximum, like Eigen's maxCoeff
  if (len[1] > len[best]) best = 1;
  if (len[2] > len[best]) best = 2;
  for (int d = 0; d < 3; ++d) v[d] = c[best][d] / len[best];
}
// Eigen::MatrixBase::unitOrthogonal for a 3-vector
inline void unitOrthogonal(const Eigen::Vector3f& v, Eigen::Vector3f& o)
{
  auto much_smaller = [](float a, float b) { return std::abs(a) <= std::abs(b) * std::numeric_limits<float>::epsilon(); };
  if (!much_smaller(v[0], v[2]) || !much_smaller(v[1], v[2])) {
    const float invnm = 1.0f / std::sqrt(v[0] * v[0] + v[1] * v[1]);
    o[0] = -v[1] * invnm; o[1] = v[0] * invnm; o[2] = 0.0f;
  }
  else {
    const float invnm = 1.0f / std::sqrt(v[1] * v[1] + v[2] * v[2]);
    o[0] = 0.0f; o[1] = -v[2] * invnm; o[2] = v[1] * invnm;
  }
}
}  // namespace detail

// smallest eigenvalue of a symmetric PSD 3x3 matrix and its unit eigenvector (eigen.hpp:293-326): the matrix is scaled
// by its largest |entry| first; a (numerically) repeated smallest eigenvalue gives a vector orthogonal to the largest
// eigenvector, a multiple of the identity gives (1, 0, 0)
inline void eigen33(const Eigen::Matrix3f& mat, float& eigenvalue, Eigen::Vector3f& eigenvector)
{
  float scale = 0.0f;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) scale = std::max(scale, std::abs(mat(i, j)));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  Eigen::Matrix3f s;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) s(i, j) = mat(i, j) / scale;
  float roots[3];
  computeRoots(s, roots);
  eigenvalue = roots[0] * scale;
  if ((roots[1] - roots[0]) > std::numeric_limits<float>::epsilon()) {
    for (int d = 0; d < 3; ++d) s(d, d) -= roots[0];
    detail::largest3x3Eigenvector(s, eigenvector);
  }
  else if ((roots[2] - roots[0]) > std::numeric_limits<float>::epsilon()) {
    for (int d = 0; d < 3; ++d) s(d, d) -= roots[2];
    Eigen::Vector3f big;
    detail::largest3x3Eigenvector(s, big);
    detail::unitOrthogonal(big, eigenvector);
  }
  else {
    eigenvector[0] = 1.0f; eigenvector[1] = 0.0f; eigenvector[2] = 0.0f;
  }
}
}  // namespace pcl
