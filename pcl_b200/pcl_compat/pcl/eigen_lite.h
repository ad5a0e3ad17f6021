// pcl/eigen_lite.h — the few Eigen types PCL's registration API exposes (Matrix4f/Matrix4d/Vector4f), for
// builds where Eigen is absent (this image).  With real Eigen present, include <Eigen/Core> BEFORE the facade
// headers and define PCLB200_USE_EIGEN: the facade then uses Eigen's own types and this file is skipped.
#pragma once
#ifdef PCLB200_USE_EIGEN
#include <Eigen/Core>
#else
#include <cmath>
#include <cstddef>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>
namespace Eigen {
template <typename S, int R, int C>
struct Matrix {
  S m[R * C];  // column-major like Eigen's default
  Matrix() { for (int i = 0; i < R * C; ++i) m[i] = S(0); }
  // fixed-size vectors from their coefficients (Eigen::Vector3i(0, 0, 1), Eigen::Vector4f(x, y, z, w))
  Matrix(S a, S b, S c) { static_assert(R * C == 3, "three coefficients"); m[0] = a; m[1] = b; m[2] = c; }
  Matrix(S a, S b, S c, S d) { static_assert(R * C == 4, "four coefficients"); m[0] = a; m[1] = b; m[2] = c; m[3] = d; }
  static Matrix Identity()
  {
    Matrix r;
    for (int i = 0; i < (R < C ? R : C); ++i) r(i, i) = S(1);
    return r;
  }
  static Matrix Zero() { return Matrix(); }
  S& operator()(int r, int c) { return m[c * R + r]; }
  const S& operator()(int r, int c) const { return m[c * R + r]; }
  S& operator[](int i) { return m[i]; }
  const S& operator[](int i) const { return m[i]; }
  S& coeffRef(int r, int c) { return (*this)(r, c); }
  S coeff(int r, int c) const { return (*this)(r, c); }
  static constexpr int rows() { return R; }
  static constexpr int cols() { return C; }
  void setIdentity() { *this = Identity(); }
  bool operator==(const Matrix& o) const
  {
    for (int i = 0; i < R * C; ++i)
      if (m[i] != o.m[i]) return false;
    return true;
  }
  bool operator!=(const Matrix& o) const { return !(*this == o); }
  template <typename T>
  Matrix<T, R, C> cast() const
  {
    Matrix<T, R, C> r;
    for (int i = 0; i < R * C; ++i) r.m[i] = static_cast<T>(m[i]);
    return r;
  }
  // coefficient-based product, Eigen order ((a0*b0 + a1*b1) + a2*b2) + a3*b3
  template <int K>
  Matrix<S, R, K> operator*(const Matrix<S, C, K>& o) const
  {
    Matrix<S, R, K> r;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < K; ++j) {
        S acc = (*this)(i, 0) * o(0, j);
        for (int k = 1; k < C; ++k) acc = acc + (*this)(i, k) * o(k, j);
        r(i, j) = acc;
      }
    return r;
  }
};
// Eigen's default IOFormat: coefficients right-aligned to the widest one, columns separated by one space, rows by a newline
template <typename S, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<S, R, C>& m)
{
  std::string cell[R * C];
  std::size_t width = 0;
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      std::ostringstream one;
      one.copyfmt(os);
      one.width(0);
      one << m(r, c);
      cell[r * C + c] = one.str();
      width = width > cell[r * C + c].size() ? width : cell[r * C + c].size();
    }
  for (int r = 0; r < R; ++r) {
    if (r) os << "\n";
    for (int c = 0; c < C; ++c) {
      if (c) os << " ";
      os << std::string(width - cell[r * C + c].size(), ' ') << cell[r * C + c];
    }
  }
  return os;
}
using Matrix4f = Matrix<float, 4, 4>;
using Matrix4d = Matrix<double, 4, 4>;
using Matrix3f = Matrix<float, 3, 3>;
using Vector4f = Matrix<float, 4, 1>;
using Vector3f = Matrix<float, 3, 1>;
using Vector3i = Matrix<int, 3, 1>;
using Vector4i = Matrix<int, 4, 1>;
// run-time sized integer matrix, column-major: the lists of relative cell coordinates of pcl::VoxelGrid's neighbour
// queries (voxel_grid.h:52-100, 356-380)
struct MatrixXi {
  int r = 0, c = 0;
  std::vector<int> v;
  MatrixXi() = default;
  MatrixXi(int rows_, int cols_) : r(rows_), c(cols_), v(static_cast<std::size_t>(rows_) * cols_, 0) {}
  static MatrixXi Zero(int rows_, int cols_) { return MatrixXi(rows_, cols_); }
  template <int R, int C>
  MatrixXi(const Matrix<int, R, C>& f) : r(R), c(C), v(f.m, f.m + R * C) {}  // both column-major
  int rows() const { return r; }
  int cols() const { return c; }
  int& operator()(int i, int j) { return v[static_cast<std::size_t>(j) * r + i]; }
  int operator()(int i, int j) const { return v[static_cast<std::size_t>(j) * r + i]; }
  void conservativeResize(int rows_, int cols_)
  {
    MatrixXi n(rows_, cols_);
    for (int j = 0; j < (cols_ < c ? cols_ : c); ++j)
      for (int i = 0; i < (rows_ < r ? rows_ : r); ++i) n(i, j) = (*this)(i, j);
    *this = n;
  }
};
// the sensor pose of pcl::PointCloud (point_cloud.h:405-407) is carried, never computed with
struct Quaternionf {
  float qw = 1.f, qx = 0.f, qy = 0.f, qz = 0.f;
  Quaternionf() = default;
  Quaternionf(float w_, float x_, float y_, float z_) : qw(w_), qx(x_), qy(y_), qz(z_) {}
  static Quaternionf Identity() { return Quaternionf(); }
  float w() const { return qw; }
  float x() const { return qx; }
  float y() const { return qy; }
  float z() const { return qz; }
  float& w() { return qw; }
  float& x() { return qx; }
  float& y() { return qy; }
  float& z() { return qz; }
  bool operator==(const Quaternionf& o) const { return qw == o.qw && qx == o.qx && qy == o.qy && qz == o.qz; }
};
}  // namespace Eigen
#endif
