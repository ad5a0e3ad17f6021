// pcl/PCLPointCloud2.h + pcl/conversions.h — the type-erased cloud ("blob") PCL passes where the point type is not known
// statically (common/include/pcl/PCLPointCloud2.h:20-80, PCLPointField.h:13-40, conversions.h:166-330): on this path,
// IterativeClosestPoint hands normals to CorrespondenceEstimationBase::setSourceNormals / setTargetNormals as blobs
// (registration/impl/icp.hpp:142,169; correspondence_estimation.h:277-322).  Host code; fields are matched by NAME.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "point_cloud.h"
#include "point_types.h"

namespace pcl {

struct PCLPointField {
  std::string name;
  std::uint32_t offset = 0;
  std::uint8_t datatype = 0;
  std::uint32_t count = 0;
  enum PointFieldTypes { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
};

struct PCLPointCloud2 {
  using Ptr = std::shared_ptr<PCLPointCloud2>;
  using ConstPtr = std::shared_ptr<const PCLPointCloud2>;
  PCLHeader header;
  std::uint32_t height = 0;
  std::uint32_t width = 0;
  std::vector<PCLPointField> fields;
  std::uint8_t is_bigendian = 0;
  std::uint32_t point_step = 0;
  std::uint32_t row_step = 0;
  std::vector<std::uint8_t> data;
  std::uint8_t is_dense = 0;
};

namespace detail {
struct BlobField {
  const char* name;
  std::uint32_t offset;
};
template <typename P> struct blob_fields;
template <> struct blob_fields<PointXYZ> {
  static std::vector<BlobField> get() { return {{"x", 0}, {"y", 4}, {"z", 8}}; }
};
template <> struct blob_fields<Normal> {
  static std::vector<BlobField> get() { return {{"normal_x", 0}, {"normal_y", 4}, {"normal_z", 8}, {"curvature", 16}}; }
};
template <> struct blob_fields<PointNormal> {
  static std::vector<BlobField> get()
  {
    return {{"x", 0}, {"y", 4}, {"z", 8}, {"normal_x", 16}, {"normal_y", 20}, {"normal_z", 24}, {"curvature", 32}};
  }
};
}  // namespace detail

// conversions.h:275-330: the records verbatim + the field list of the point type
template <typename PointT>
void toPCLPointCloud2(const pcl::PointCloud<PointT>& cloud, pcl::PCLPointCloud2& msg)
{
  msg.header = cloud.header;
  if (cloud.width == 0 && cloud.height == 0) {
    msg.width = static_cast<std::uint32_t>(cloud.size());
    msg.height = 1;
  }
  else {
    msg.height = cloud.height;
    msg.width = cloud.width;
  }
  msg.point_step = sizeof(PointT);
  msg.row_step = msg.point_step * msg.width;
  msg.is_bigendian = 0;
  msg.is_dense = cloud.is_dense ? 1 : 0;
  msg.data.resize(sizeof(PointT) * cloud.size());
  if (!cloud.empty())
    std::memcpy(msg.data.data(), cloud.points.data(), msg.data.size());
  msg.fields.clear();
  for (const auto& f : detail::blob_fields<PointT>::get()) {
    PCLPointField pf;
    pf.name = f.name;
    pf.offset = f.offset;
    pf.datatype = PCLPointField::FLOAT32;
    pf.count = 1;
    msg.fields.push_back(pf);
  }
}

// conversions.h:166-250 (createMapping + fromPCLPointCloud2): every field of PointT is looked up by name and must be a
// FLOAT32 of count 1; fields the blob lacks are reported and left default-initialised, as the reference's warning does
template <typename PointT>
void fromPCLPointCloud2(const pcl::PCLPointCloud2& msg, pcl::PointCloud<PointT>& cloud)
{
  cloud.header = msg.header;
  cloud.width = msg.width;
  cloud.height = msg.height;
  cloud.is_dense = msg.is_dense == 1;
  const std::size_t n = static_cast<std::size_t>(msg.width) * msg.height;
  cloud.points.assign(n, PointT());
  if (msg.point_step == 0 || msg.data.size() < n * static_cast<std::size_t>(msg.point_step)) {
    if (n) std::fprintf(stderr, "[pcl::fromPCLPointCloud2] blob holds fewer bytes than width x height x point_step\n");
    cloud.points.clear();
    cloud.width = cloud.height = 0;
    return;
  }
  for (const auto& want : detail::blob_fields<PointT>::get()) {
    const PCLPointField* src = nullptr;
    for (const auto& f : msg.fields)
      if (f.name == want.name && f.datatype == PCLPointField::FLOAT32 && f.count >= 1 && f.offset + 4 <= msg.point_step) src = &f;
    if (!src) {
      std::fprintf(stderr, "Failed to find match for field '%s'.\n", want.name);
      continue;
    }
    for (std::size_t i = 0; i < n; ++i)
      std::memcpy(reinterpret_cast<unsigned char*>(&cloud.points[i]) + want.offset,
                  msg.data.data() + i * msg.point_step + src->offset, 4);
  }
}

}  // namespace pcl
