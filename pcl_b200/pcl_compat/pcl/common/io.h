// pcl/common/io.h — field bookkeeping of the type-erased cloud (common/include/pcl/common/io.h:60-170,
// common/src/io.cpp:44-210): field look-up, the printable field list, sizes / PCD type letters of the field datatypes,
// and concatenateFields, which a tools/iterative_closest_point.cpp-style program uses to put the aligned coordinates
// back beside the input's remaining fields.  Host code.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../PCLPointCloud2.h"
#include "../point_cloud.h"
#include "../types.h"

namespace pcl {
inline int getFieldIndex(const pcl::PCLPointCloud2& cloud, const std::string& field_name)
{
  for (std::size_t d = 0; d < cloud.fields.size(); ++d)
    if (cloud.fields[d].name == field_name) return static_cast<int>(d);
  return -1;
}
inline std::string getFieldsList(const pcl::PCLPointCloud2& cloud)
{
  std::string result;
  for (std::size_t i = 0; i < cloud.fields.size(); ++i) result += (i ? " " : "") + cloud.fields[i].name;
  return result;
}
inline int getFieldSize(const int datatype)
{
  switch (datatype) {
    case PCLPointField::INT8: case PCLPointField::UINT8: return 1;
    case PCLPointField::INT16: case PCLPointField::UINT16: return 2;
    case PCLPointField::INT32: case PCLPointField::UINT32: case PCLPointField::FLOAT32: return 4;
    case PCLPointField::FLOAT64: return 8;
    default: return 0;
  }
}
inline int getFieldType(const int size, char type)   // (SIZE, TYPE) of a PCD header -> datatype, -1 if there is none
{
  type = static_cast<char>(std::toupper(static_cast<unsigned char>(type)));
  if (type == 'I') return size == 1 ? PCLPointField::INT8 : size == 2 ? PCLPointField::INT16 : size == 4 ? PCLPointField::INT32 : -1;
  if (type == 'U') return size == 1 ? PCLPointField::UINT8 : size == 2 ? PCLPointField::UINT16 : size == 4 ? PCLPointField::UINT32 : -1;
  if (type == 'F') return size == 4 ? PCLPointField::FLOAT32 : size == 8 ? PCLPointField::FLOAT64 : -1;
  return -1;
}
inline char getFieldType(const int datatype)          // datatype -> the PCD TYPE letter
{
  switch (datatype) {
    case PCLPointField::INT8: case PCLPointField::INT16: case PCLPointField::INT32: return 'I';
    case PCLPointField::UINT8: case PCLPointField::UINT16: case PCLPointField::UINT32: return 'U';
    case PCLPointField::FLOAT32: case PCLPointField::FLOAT64: return 'F';
    default: return '?';
  }
}

// copyPointCloud (common/include/pcl/common/impl/io.hpp:112-260): same point type — all fields; an index list gives an
// unorganised cloud of the selected points; between point types the coordinates are what both sides share here
template <typename PointT>
inline void copyPointCloud(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out)
{
  cloud_out = cloud_in;
}
template <typename PointT>
inline void copyPointCloud(const pcl::PointCloud<PointT>& cloud_in, const Indices& indices, pcl::PointCloud<PointT>& cloud_out)
{
  if (indices.size() == cloud_in.size()) {   // io.hpp:141-146: a full-length list is taken for the identity
    cloud_out = cloud_in;
    return;
  }
  pcl::PointCloud<PointT> out;
  out.points.resize(indices.size());
  out.header = cloud_in.header;
  out.width = static_cast<std::uint32_t>(indices.size());
  out.height = 1;
  out.is_dense = cloud_in.is_dense;
  out.sensor_orientation_ = cloud_in.sensor_orientation_;
  out.sensor_origin_ = cloud_in.sensor_origin_;
  for (std::size_t i = 0; i < indices.size(); ++i) out.points[i] = cloud_in[static_cast<std::size_t>(indices[i])];
  cloud_out = std::move(out);
}
template <typename PointInT, typename PointOutT>
inline void copyPointCloud(const pcl::PointCloud<PointInT>& cloud_in, pcl::PointCloud<PointOutT>& cloud_out)
{
  cloud_out.header = cloud_in.header;
  cloud_out.width = cloud_in.width;
  cloud_out.height = cloud_in.height;
  cloud_out.is_dense = cloud_in.is_dense;
  cloud_out.sensor_orientation_ = cloud_in.sensor_orientation_;
  cloud_out.sensor_origin_ = cloud_in.sensor_origin_;
  cloud_out.points.assign(cloud_in.size(), PointOutT());
  for (std::size_t i = 0; i < cloud_in.size(); ++i) {
    cloud_out.points[i].x = cloud_in[i].x;
    cloud_out.points[i].y = cloud_in[i].y;
    cloud_out.points[i].z = cloud_in[i].z;
  }
}

// common/include/pcl/common/impl/io.hpp:367-399: every field of the output type that one of the inputs has, by NAME — cloud1's
// first, then cloud2's (so cloud2 wins where both have it); metadata of cloud1; dense only if both are
template <typename PointIn1T, typename PointIn2T, typename PointOutT>
inline void concatenateFields(const pcl::PointCloud<PointIn1T>& cloud1_in, const pcl::PointCloud<PointIn2T>& cloud2_in, pcl::PointCloud<PointOutT>& cloud_out)
{
  if (cloud1_in.size() != cloud2_in.size()) {
    std::fprintf(stderr, "[pcl::concatenateFields] The number of points in the two input datasets differs!\n");
    return;
  }
  cloud_out.points.resize(cloud1_in.size());
  cloud_out.header = cloud1_in.header;
  cloud_out.width = cloud1_in.width;
  cloud_out.height = cloud1_in.height;
  cloud_out.is_dense = cloud1_in.is_dense && cloud2_in.is_dense;
  const auto out_fields = detail::blob_fields<PointOutT>::get();
  auto copy_from = [&](const auto& in_fields, const unsigned char* src, unsigned char* dst) {
    for (const auto& fi : in_fields)
      for (const auto& fo : out_fields)
        if (std::strcmp(fi.name, fo.name) == 0) std::memcpy(dst + fo.offset, src + fi.offset, 4);   // every field of these types is one FLOAT32
  };
  const auto f1 = detail::blob_fields<PointIn1T>::get();
  const auto f2 = detail::blob_fields<PointIn2T>::get();
  for (std::size_t i = 0; i < cloud_out.size(); ++i) {
    unsigned char* dst = reinterpret_cast<unsigned char*>(&cloud_out.points[i]);
    copy_from(f1, reinterpret_cast<const unsigned char*>(&cloud1_in.points[i]), dst);
    copy_from(f2, reinterpret_cast<const unsigned char*>(&cloud2_in.points[i]), dst);
  }
}

// cloud_out = the records of cloud2 followed, per point, by the fields of cloud1 that cloud2 does not have (by name;
// "_" padding never carried over).  Each carried field keeps the room it had in cloud1 up to the next named field, the
// slack zero-filled (common/src/io.cpp:69-210).  Both clouds must have the same width and height.
inline bool concatenateFields(const pcl::PCLPointCloud2& cloud1, const pcl::PCLPointCloud2& cloud2, pcl::PCLPointCloud2& cloud_out)
{
  if (cloud1.width != cloud2.width || cloud1.height != cloud2.height) {
    std::fprintf(stderr, "[pcl::concatenateFields] Dimensions of input clouds do not match: cloud1 (w, %u, h, %u), cloud2 (w, %u, h, %u)\n",
                 cloud1.width, cloud1.height, cloud2.width, cloud2.height);
    return false;
  }
  if (cloud1.is_bigendian != cloud2.is_bigendian) {
    std::fprintf(stderr, "[pcl::concatenateFields] Endianness of clouds does not match\n");
    return false;
  }
  std::vector<const PCLPointField*> by_offset;
  for (const auto& f : cloud1.fields) by_offset.push_back(&f);
  std::sort(by_offset.begin(), by_offset.end(), [](const PCLPointField* a, const PCLPointField* b) { return a->offset < b->offset; });
  std::vector<const PCLPointField*> carried;
  std::vector<std::uint32_t> room;
  for (std::size_t i = 0; i < by_offset.size(); ++i) {
    const PCLPointField& f = *by_offset[i];
    if (f.name == "_" || getFieldIndex(cloud2, f.name) >= 0) continue;
    std::size_t next = i + 1;
    while (next < by_offset.size() && by_offset[next]->name == "_") ++next;
    const std::uint32_t end = next < by_offset.size() ? by_offset[next]->offset : cloud1.point_step;
    carried.push_back(&f);
    room.push_back(end - f.offset);
  }
  std::uint32_t extra = 0;
  for (std::uint32_t r : room) extra += r;
  pcl::PCLPointCloud2 out;
  out.header = cloud2.header;
  out.fields = cloud2.fields;
  out.width = cloud2.width;
  out.height = cloud2.height;
  out.is_bigendian = cloud2.is_bigendian;
  out.is_dense = cloud1.is_dense && cloud2.is_dense;
  out.point_step = cloud2.point_step + extra;
  out.row_step = out.point_step * out.width;
  const std::size_t npts = static_cast<std::size_t>(out.width) * out.height;
  out.data.assign(npts * out.point_step, 0);
  std::uint32_t offset = cloud2.point_step;
  for (std::size_t d = 0; d < carried.size(); ++d) {
    PCLPointField nf = *carried[d];
    nf.offset = offset;
    out.fields.push_back(nf);
    offset += room[d];
  }
  for (std::size_t cp = 0; cp < npts; ++cp) {
    std::uint8_t* dst = out.data.data() + cp * out.point_step;
    std::memcpy(dst, cloud2.data.data() + cp * cloud2.point_step, cloud2.point_step);
    std::uint32_t at = cloud2.point_step;
    for (std::size_t d = 0; d < carried.size(); ++d) {
      const std::uint32_t bytes = carried[d]->count * static_cast<std::uint32_t>(getFieldSize(carried[d]->datatype));
      std::memcpy(dst + at, cloud1.data.data() + cp * cloud1.point_step + carried[d]->offset, std::min(bytes, room[d]));
      at += room[d];
    }
  }
  cloud_out = std::move(out);
  return true;
}
}  // namespace pcl
