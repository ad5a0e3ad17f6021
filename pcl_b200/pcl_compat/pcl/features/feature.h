// pcl/features/feature.h — solvePlaneParameters (features/include/pcl/features/impl/feature.hpp:52-92): least-squares
// plane of a neighbourhood from its covariance — the normal is the eigenvector of the smallest eigenvalue, the surface
// curvature that eigenvalue over the trace.  Host side; the device does the same in search.cu: normal_from_moments.
#pragma once
#include <cmath>

#include "../common/eigen.h"

namespace pcl {
inline void solvePlaneParameters(const Eigen::Matrix3f& covariance_matrix, float& nx, float& ny, float& nz, float& curvature)
{
  float eigen_value = 0.f;
  Eigen::Vector3f eigen_vector;
  pcl::eigen33(covariance_matrix, eigen_value, eigen_vector);
  nx = eigen_vector[0];
  ny = eigen_vector[1];
  nz = eigen_vector[2];
  const float eig_sum = covariance_matrix(0, 0) + covariance_matrix(1, 1) + covariance_matrix(2, 2);
  curvature = eig_sum != 0 ? std::fabs(eigen_value / eig_sum) : 0.f;
}
// Hessian normal form: plane_parameters = (n, d) with d = -n . point (point = the neighbourhood's centroid)
inline void solvePlaneParameters(const Eigen::Matrix3f& covariance_matrix, const Eigen::Vector4f& point,
                                 Eigen::Vector4f& plane_parameters, float& curvature)
{
  solvePlaneParameters(covariance_matrix, plane_parameters[0], plane_parameters[1], plane_parameters[2], curvature);
  plane_parameters[3] = 0;
  plane_parameters[3] = -1 * (plane_parameters[0] * point[0] + plane_parameters[1] * point[1] + plane_parameters[2] * point[2] +
                              plane_parameters[3] * point[3]);
}
}  // namespace pcl
