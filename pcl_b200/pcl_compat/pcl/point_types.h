// pcl/point_types.h — the three point layouts on the ICP path, byte-identical to
// common/include/pcl/impl/point_types.hpp:205-227 (PointXYZ, 16 B), :769-794 (Normal, 32 B),
// :824-855 (PointNormal, 48 B); all 16-byte aligned.
#pragma once
#include <cstddef>
#include <cmath>
#include <iostream>
#include <ostream>
namespace pcl {
struct alignas(16) PointXYZ {
  union {
    float data[4];
    struct { float x, y, z; };
  };
  PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
  PointXYZ(float _x, float _y, float _z) : data{_x, _y, _z, 1.f} {}
};
struct alignas(16) Normal {
  union {
    float data_n[4];
    float normal[3];
    struct { float normal_x, normal_y, normal_z; };
  };
  union {
    struct { float curvature; };
    float data_c[4];
  };
  Normal() : data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
  Normal(float nx, float ny, float nz, float c = 0.f) : data_n{nx, ny, nz, 0.f}, data_c{c, 0.f, 0.f, 0.f} {}
};
struct alignas(16) PointNormal {
  union {
    float data[4];
    struct { float x, y, z; };
  };
  union {
    float data_n[4];
    float normal[3];
    struct { float normal_x, normal_y, normal_z; };
  };
  union {
    struct { float curvature; };
    float data_c[4];
  };
  PointNormal() : data{0.f, 0.f, 0.f, 1.f}, data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
  PointNormal(float _x, float _y, float _z, float nx = 0.f, float ny = 0.f, float nz = 0.f, float c = 0.f)
  : data{_x, _y, _z, 1.f}, data_n{nx, ny, nz, 0.f}, data_c{c, 0.f, 0.f, 0.f} {}
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(Normal) == 32 && sizeof(PointNormal) == 48, "PCL layouts");

template <typename T> struct has_normal { static constexpr bool value = false; };
template <> struct has_normal<PointNormal> { static constexpr bool value = true; };
template <> struct has_normal<Normal> { static constexpr bool value = true; };
// byte offset of normal_x inside a record (pcl::Normal: 0, pcl::PointNormal: 16 — impl/point_types.hpp:769-855)
template <typename T> struct normal_offset { static constexpr std::size_t value = 16; };
template <> struct normal_offset<Normal> { static constexpr std::size_t value = 0; };

template <typename PointT> inline bool isXYZFinite(const PointT& p)
{
  return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z);
}
template <typename PointT> inline bool isFinite(const PointT& p) { return isXYZFinite(p); }
template <> inline bool isFinite<Normal>(const Normal& n) { return std::isfinite(n.normal_x) && std::isfinite(n.normal_y) && std::isfinite(n.normal_z); }   // point_tests.h:122-127
// common/src/point_types.cpp:41-46, 168-173, 189-194
inline std::ostream& operator<<(std::ostream& os, const PointXYZ& p)
{
  os << "(" << p.x << "," << p.y << "," << p.z << ")";
  return os;
}
inline std::ostream& operator<<(std::ostream& os, const Normal& p)
{
  os << "(" << p.normal[0] << "," << p.normal[1] << "," << p.normal[2] << " - " << p.curvature << ")";
  return os;
}
inline std::ostream& operator<<(std::ostream& os, const PointNormal& p)
{
  os << "(" << p.x << "," << p.y << "," << p.z << " - " << p.normal[0] << "," << p.normal[1] << "," << p.normal[2] << " - " << p.curvature << ")";
  return os;
}
}  // namespace pcl
