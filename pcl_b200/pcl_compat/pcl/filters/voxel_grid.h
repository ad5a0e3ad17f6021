// pcl/filters/voxel_grid.h — pcl::VoxelGrid<PointT> (filters/include/pcl/filters/voxel_grid.h:220-530,
// impl/voxel_grid.hpp:596-814) on the device.  Supported: setLeafSize, setMinimumPointsNumberPerVoxel, indices,
// filter().  downsample_all_data_ (default true, voxel_grid.hpp:796-806): for point types that carry a normal and a
// curvature (PointNormal) those fields are averaged on the device too, like CentroidPoint does (normal = normalised
// 4-vector sum, curvature = mean); with setDownsampleAllData(false) only xyz is produced and the other fields are
// default-initialised (voxel_grid.hpp:784-794 copies just the 4-float centroid).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../b200/context.h"
#include "../point_cloud.h"
#include "../point_types.h"
#include "../PCLPointCloud2.h"
#include "../common/io.h"

namespace pcl {
// voxel_grid.h:52-100: the 13 cells of the "upper half" of a cell's 26-neighbourhood, and all 26 + the cell itself
inline Eigen::MatrixXi getHalfNeighborCellIndices()
{
  Eigen::MatrixXi rel(3, 13);
  int idx = 0;
  for (int i = -1; i < 2; ++i)      // 0 - 8
    for (int j = -1; j < 2; ++j) {
      rel(0, idx) = i; rel(1, idx) = j; rel(2, idx) = -1;
      ++idx;
    }
  for (int i = -1; i < 2; ++i) {    // 9 - 11
    rel(0, idx) = i; rel(1, idx) = -1; rel(2, idx) = 0;
    ++idx;
  }
  rel(0, idx) = -1; rel(1, idx) = 0; rel(2, idx) = 0;  // 12
  return rel;
}
inline Eigen::MatrixXi getAllNeighborCellIndices()
{
  const Eigen::MatrixXi half = getHalfNeighborCellIndices();
  Eigen::MatrixXi all(3, 27);   // the 13, the cell itself (zeros), the 13 mirrored
  for (int j = 0; j < 13; ++j)
    for (int i = 0; i < 3; ++i) {
      all(i, j) = half(i, j);
      all(i, 14 + j) = -half(i, j);
    }
  return all;
}

namespace b200 {
// Host-side bookkeeping of one VoxelGrid::applyFilter call (impl/voxel_grid.hpp:612-650, 700-780): the integer box of
// the used points, the divisions, the linear-index multipliers and — on request — the leaf layout (cell -> position of
// its centroid in the output, -1 for empty cells and cells with fewer than min_points points).  The arithmetic is the
// reference's: float products, floor, the minimum subtracted as a float.
struct VoxelLayout {
  Eigen::Vector4i min_b, max_b, div_b, divb_mul;
  std::vector<int> leaf_layout;
  std::size_t cells_kept = 0;
};
template <typename PointT>
inline bool voxel_layout(const std::vector<PointT>& pts, const index_t* idx, std::size_t n, bool is_dense, const float inv_leaf[3],
                         unsigned int min_points, bool want_layout, VoxelLayout& L)
{
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  auto point = [&](std::size_t j) -> const PointT& { return pts[idx ? static_cast<std::size_t>(idx[j]) : j]; };
  auto finite = [](const PointT& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); };
  bool any = false;
  for (std::size_t j = 0; j < n; ++j) {
    const PointT& p = point(j);
    if (!is_dense && !finite(p)) continue;
    any = true;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  if (!any) return false;
  for (int a = 0; a < 3; ++a) {
    L.min_b[a] = static_cast<int>(std::floor(mn[a] * inv_leaf[a]));
    L.max_b[a] = static_cast<int>(std::floor(mx[a] * inv_leaf[a]));
    L.div_b[a] = L.max_b[a] - L.min_b[a] + 1;
  }
  L.min_b[3] = L.max_b[3] = L.div_b[3] = 0;
  L.divb_mul[0] = 1; L.divb_mul[1] = L.div_b[0]; L.divb_mul[2] = L.div_b[0] * L.div_b[1]; L.divb_mul[3] = 0;
  L.leaf_layout.clear();
  L.cells_kept = 0;
  // an oversize grid makes the filter return its input unfiltered (voxel_grid.hpp:620-629): no layout then
  const long long cells = static_cast<long long>(L.div_b[0]) * L.div_b[1] * L.div_b[2];
  if (!want_layout || cells > static_cast<long long>(std::numeric_limits<std::int32_t>::max())) return true;
  std::vector<int> keys;
  keys.reserve(n);
  for (std::size_t j = 0; j < n; ++j) {
    const PointT& p = point(j);
    if (!is_dense && !finite(p)) continue;
    const int i0 = static_cast<int>(std::floor(p.x * inv_leaf[0]) - static_cast<float>(L.min_b[0]));
    const int i1 = static_cast<int>(std::floor(p.y * inv_leaf[1]) - static_cast<float>(L.min_b[1]));
    const int i2 = static_cast<int>(std::floor(p.z * inv_leaf[2]) - static_cast<float>(L.min_b[2]));
    keys.push_back(i0 * L.divb_mul[0] + i1 * L.divb_mul[1] + i2 * L.divb_mul[2]);
  }
  std::sort(keys.begin(), keys.end());
  L.leaf_layout.assign(static_cast<std::size_t>(L.div_b[0]) * L.div_b[1] * L.div_b[2], -1);
  int out = 0;
  for (std::size_t a = 0; a < keys.size();) {
    std::size_t b = a + 1;
    while (b < keys.size() && keys[b] == keys[a]) ++b;
    if (b - a >= min_points) L.leaf_layout[static_cast<std::size_t>(keys[a])] = out++;
    a = b;
  }
  L.cells_kept = static_cast<std::size_t>(out);
  return true;
}
}  // namespace b200

template <typename PointT>
class VoxelGrid : public PCLBase<PointT> {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  // voxel_grid.h:296-425: grid geometry of the last filter() call and, with setSaveLeafLayout(true), the leaf layout
  bool getDownsampleAllData() const { return downsample_all_data_; }
  void setSaveLeafLayout(bool save_leaf_layout) { save_leaf_layout_ = save_leaf_layout; }
  bool getSaveLeafLayout() const { return save_leaf_layout_; }
  Eigen::Vector3i getMinBoxCoordinates() const { return head3(grid_.min_b); }
  Eigen::Vector3i getMaxBoxCoordinates() const { return head3(grid_.max_b); }
  Eigen::Vector3i getNrDivisions() const { return head3(grid_.div_b); }
  Eigen::Vector3i getDivisionMultiplier() const { return head3(grid_.divb_mul); }
  std::vector<int> getLeafLayout() const { return grid_.leaf_layout; }
  Eigen::Vector3i getGridCoordinates(float x, float y, float z) const
  {
    Eigen::Vector3i g;
    g[0] = static_cast<int>(std::floor(x * invLeaf(0)));
    g[1] = static_cast<int>(std::floor(y * invLeaf(1)));
    g[2] = static_cast<int>(std::floor(z * invLeaf(2)));
    return g;
  }
  int getCentroidIndexAt(const Eigen::Vector3i& ijk) const
  {
    const long long idx = linearIndex(ijk[0], ijk[1], ijk[2]);
    if (idx < 0 || idx >= static_cast<long long>(grid_.leaf_layout.size())) return -1;
    return grid_.leaf_layout[static_cast<std::size_t>(idx)];
  }
  // index of the centroid of the cell that holds p; like the reference (std::vector::at) it throws for a cell outside
  // the grid or when no layout was saved
  int getCentroidIndex(const PointT& p) const
  {
    const Eigen::Vector3i g = getGridCoordinates(p.x, p.y, p.z);
    return grid_.leaf_layout.at(static_cast<std::size_t>(linearIndex(g[0], g[1], g[2])));
  }
  std::vector<int> getNeighborCentroidIndices(const PointT& reference_point, const Eigen::MatrixXi& relative_coordinates) const
  {
    const Eigen::Vector3i g = getGridCoordinates(reference_point.x, reference_point.y, reference_point.z);
    std::vector<int> neighbors(static_cast<std::size_t>(relative_coordinates.cols()), -1);
    for (int ni = 0; ni < relative_coordinates.cols(); ++ni) {
      bool inside = true;
      int c[3];
      for (int a = 0; a < 3; ++a) {
        c[a] = g[a] + relative_coordinates(a, ni);
        inside = inside && c[a] >= grid_.min_b[a] && c[a] <= grid_.max_b[a];
      }
      if (inside && !grid_.leaf_layout.empty())
        neighbors[static_cast<std::size_t>(ni)] = grid_.leaf_layout[static_cast<std::size_t>(linearIndex(c[0], c[1], c[2]))];
    }
    return neighbors;
  }
  bool getFilterLimitsNegative() const { return filter_limit_negative_; }
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void setMinimumPointsNumberPerVoxel(unsigned int n) { min_points_per_voxel_ = n; }
  unsigned int getMinimumPointsNumberPerVoxel() const { return min_points_per_voxel_; }
  void setDownsampleAllData(bool v) { downsample_all_data_ = v; }
  void setLeafSize(const Eigen::Vector4f& l) { leaf_[0] = l[0]; leaf_[1] = l[1]; leaf_[2] = l[2]; }
  Eigen::Vector3f getLeafSize() const
  {
    Eigen::Vector3f l;
    l[0] = leaf_[0]; l[1] = leaf_[1]; l[2] = leaf_[2];
    return l;
  }
  // voxel_grid.h:393-470: pass-through limits on one coordinate field before the grid.  Only "x", "y", "z" exist on
  // the point types of this path; min/max (and therefore the grid origin) are taken over the selected points, exactly
  // like getMinMax3D with a filter field (voxel_grid.hpp:614-617, 663-690).
  void setFilterFieldName(const std::string& name) { filter_field_name_ = name; }
  const std::string& getFilterFieldName() const { return filter_field_name_; }
  void setFilterLimits(const double& lo, const double& hi) { filter_limit_min_ = lo; filter_limit_max_ = hi; }
  void getFilterLimits(double& lo, double& hi) const { lo = filter_limit_min_; hi = filter_limit_max_; }
  void setFilterLimitsNegative(bool v) { filter_limit_negative_ = v; }
  void filter(PointCloud& output)
  {
    if (!this->input_) {
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] No input dataset given!\n");
      output.clear();
      return;
    }
    PCLBase<PointT>::initCompute();
    output.header = this->input_->header;
    Indices selected;
    const index_t* abi_idx = this->abiIndices();
    std::size_t abi_cnt = this->abiIndexCount();
    if (!filter_field_name_.empty()) {
      const int f = filter_field_name_ == "x" ? 0 : filter_field_name_ == "y" ? 1 : filter_field_name_ == "z" ? 2 : -1;
      if (f < 0) {
        std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] Invalid filter field name (%s).\n", filter_field_name_.c_str());
        output.clear();
        return;
      }
      for (index_t i : *this->indices_) {
        const PointT& p = (*this->input_)[i];
        if (!this->input_->is_dense && !isXYZFinite(p)) continue;
        const float v = p.data[f];
        if (filter_limit_negative_) { if (v < filter_limit_max_ && v > filter_limit_min_) continue; }
        else if (v > filter_limit_max_ || v < filter_limit_min_) continue;
        selected.push_back(i);
      }
      if (selected.empty()) { output.clear(); return; }
      abi_idx = selected.data();
      abi_cnt = selected.size();
    }
    const std::size_t n = abi_idx ? abi_cnt : this->indices_->size();
    {  // grid geometry (always) and leaf layout (on request) of this call, like impl/voxel_grid.hpp:612-650, 757-776
      const float inv[3] = {invLeaf(0), invLeaf(1), invLeaf(2)};
      const index_t* used = abi_idx ? abi_idx : this->indices_->data();
      if (!b200::voxel_layout(this->input_->points, used, n, this->input_->is_dense, inv, min_points_per_voxel_,
                              save_leaf_layout_, grid_))
        grid_ = b200::VoxelLayout();
    }
    std::vector<float> xyz1(4 * (n ? n : 1));
    std::vector<float> ncurv;
    std::size_t m = 0;
    int rc;
    constexpr bool has_normal = pcl::has_normal<PointT>::value && sizeof(PointT) == 48;  // pcl::PointNormal: {xyz1 | normal4 | curvature, pad}
    if (has_normal && downsample_all_data_) {
      ncurv.resize(8 * (n ? n : 1));
      const unsigned char* base = reinterpret_cast<const unsigned char*>(this->input_->points.data());
      rc = pclb200_voxelgrid_normals(b200::Context::get(), base, this->input_->size(), sizeof(PointT), base + 16,
                                     sizeof(PointT), abi_idx, abi_cnt, this->input_->is_dense ? 1 : 0, leaf_,
                                     min_points_per_voxel_, xyz1.data(), ncurv.data(), &m);
    }
    else
      rc = pclb200_voxelgrid(b200::Context::get(), this->input_->points.data(), this->input_->size(), sizeof(PointT),
                             abi_idx, abi_cnt, this->input_->is_dense ? 1 : 0, leaf_, min_points_per_voxel_,
                             xyz1.data(), &m);
    if (rc == PCLB200_ERR_LEAF_TOO_SMALL) {  // voxel_grid.hpp:620-629: warn and return the input unfiltered
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] Leaf size is too small for the input dataset. Integer indices would overflow.\n");
      output = *this->input_;
      return;
    }
    if (rc != PCLB200_OK) {
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] %s\n", pclb200_last_error());
      output.clear();
      return;
    }
    if (save_leaf_layout_ && grid_.cells_kept != m)  // cannot happen: both sides run the same float arithmetic
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] leaf layout holds %zu cells, the filter produced %zu\n",
                   grid_.cells_kept, m);
    output.points.assign(m, PointT());
    for (std::size_t i = 0; i < m; ++i) {
      output.points[i].x = xyz1[4 * i];
      output.points[i].y = xyz1[4 * i + 1];
      output.points[i].z = xyz1[4 * i + 2];
      if (!ncurv.empty())  // bytes 16..47 of a PointNormal: normal4 | curvature
        std::memcpy(reinterpret_cast<unsigned char*>(&output.points[i]) + 16, &ncurv[8 * i], 32);
    }
    output.width = static_cast<std::uint32_t>(m);
    output.height = 1;       // downsampling breaks the organized structure (:609)
    output.is_dense = true;  // (:610)
  }

protected:
  static Eigen::Vector3i head3(const Eigen::Vector4i& v)
  {
    Eigen::Vector3i r;
    r[0] = v[0]; r[1] = v[1]; r[2] = v[2];
    return r;
  }
  float invLeaf(int a) const { return 1.0f / leaf_[a]; }  // inverse_leaf_size_ = 1 / leaf_size_ in float (voxel_grid.h:266-283)
  long long linearIndex(int i, int j, int k) const
  {
    return static_cast<long long>(i - grid_.min_b[0]) * grid_.divb_mul[0] + static_cast<long long>(j - grid_.min_b[1]) * grid_.divb_mul[1] +
           static_cast<long long>(k - grid_.min_b[2]) * grid_.divb_mul[2];
  }
  b200::VoxelLayout grid_;
  bool save_leaf_layout_ = false;
  float leaf_[3] = {0.f, 0.f, 0.f};
  unsigned int min_points_per_voxel_ = 0;
  bool downsample_all_data_ = true;
  std::string filter_field_name_;
  double filter_limit_min_ = std::numeric_limits<float>::lowest(), filter_limit_max_ = std::numeric_limits<float>::max();
  bool filter_limit_negative_ = false;
};

// VoxelGrid<pcl::PCLPointCloud2> (filters/include/pcl/filters/voxel_grid.h:507-870, filters/src/voxel_grid.cpp:211-545): the
// type-erased form the PCL tutorials use.  The blob's x / y / z (FLOAT32) — and, when it carries them and all data is to be
// downsampled, normal_x / normal_y / normal_z / curvature — go through the typed device filter above; the output blob holds
// exactly those fields.  Other fields of the input (intensity, rgb, ...) are not carried: the reference averages them when
// setDownsampleAllData(true), this class does not and says so once per filter() call.
template <>
class VoxelGrid<pcl::PCLPointCloud2> {
public:
  using PCLPointCloud2 = pcl::PCLPointCloud2;
  using PCLPointCloud2Ptr = PCLPointCloud2::Ptr;
  using PCLPointCloud2ConstPtr = PCLPointCloud2::ConstPtr;

  void setInputCloud(const PCLPointCloud2ConstPtr& cloud) { input_ = cloud; }
  PCLPointCloud2ConstPtr const getInputCloud() const { return input_; }
  void setIndices(const IndicesPtr& indices) { indices_ = indices; }
  void setIndices(const IndicesConstPtr& indices) { indices_.reset(new Indices(*indices)); }
  IndicesPtr getIndices() { return indices_ ? indices_ : last_indices_; }   // without setIndices: the list the last filter() ran on
  void setLeafSize(float lx, float ly, float lz) { xyz_.setLeafSize(lx, ly, lz); xyzn_.setLeafSize(lx, ly, lz); }
  void setLeafSize(const Eigen::Vector4f& l) { setLeafSize(l[0], l[1], l[2]); }
  Eigen::Vector3f getLeafSize() const { return xyz_.getLeafSize(); }
  void setDownsampleAllData(bool v) { downsample_all_data_ = v; }
  bool getDownsampleAllData() const { return downsample_all_data_; }
  void setMinimumPointsNumberPerVoxel(unsigned int n) { xyz_.setMinimumPointsNumberPerVoxel(n); xyzn_.setMinimumPointsNumberPerVoxel(n); }
  unsigned int getMinimumPointsNumberPerVoxel() const { return xyz_.getMinimumPointsNumberPerVoxel(); }
  void setSaveLeafLayout(bool v) { xyz_.setSaveLeafLayout(v); xyzn_.setSaveLeafLayout(v); }
  bool getSaveLeafLayout() const { return xyz_.getSaveLeafLayout(); }
  void setFilterFieldName(const std::string& name) { xyz_.setFilterFieldName(name); xyzn_.setFilterFieldName(name); }
  const std::string& getFilterFieldName() const { return xyz_.getFilterFieldName(); }
  void setFilterLimits(const double& lo, const double& hi) { xyz_.setFilterLimits(lo, hi); xyzn_.setFilterLimits(lo, hi); }
  void getFilterLimits(double& lo, double& hi) const { xyz_.getFilterLimits(lo, hi); }
  void setFilterLimitsNegative(bool v) { xyz_.setFilterLimitsNegative(v); xyzn_.setFilterLimitsNegative(v); }
  bool getFilterLimitsNegative() const { return xyz_.getFilterLimitsNegative(); }
  Eigen::Vector3i getMinBoxCoordinates() const { return used_normals_ ? xyzn_.getMinBoxCoordinates() : xyz_.getMinBoxCoordinates(); }
  Eigen::Vector3i getMaxBoxCoordinates() const { return used_normals_ ? xyzn_.getMaxBoxCoordinates() : xyz_.getMaxBoxCoordinates(); }
  Eigen::Vector3i getNrDivisions() const { return used_normals_ ? xyzn_.getNrDivisions() : xyz_.getNrDivisions(); }
  Eigen::Vector3i getDivisionMultiplier() const { return used_normals_ ? xyzn_.getDivisionMultiplier() : xyz_.getDivisionMultiplier(); }
  std::vector<int> getLeafLayout() const { return used_normals_ ? xyzn_.getLeafLayout() : xyz_.getLeafLayout(); }
  Eigen::Vector3i getGridCoordinates(float x, float y, float z) const { return xyz_.getGridCoordinates(x, y, z); }
  int getCentroidIndexAt(const Eigen::Vector3i& ijk) const { return used_normals_ ? xyzn_.getCentroidIndexAt(ijk) : xyz_.getCentroidIndexAt(ijk); }
  int getCentroidIndex(float x, float y, float z) const
  {
    return used_normals_ ? xyzn_.getCentroidIndex(PointNormal(x, y, z)) : xyz_.getCentroidIndex(PointXYZ(x, y, z));
  }
  std::vector<int> getNeighborCentroidIndices(float x, float y, float z, const Eigen::MatrixXi& relative_coordinates) const
  {
    return used_normals_ ? xyzn_.getNeighborCentroidIndices(PointNormal(x, y, z), relative_coordinates)
                         : xyz_.getNeighborCentroidIndices(PointXYZ(x, y, z), relative_coordinates);
  }

  void filter(PCLPointCloud2& output)
  {
    output = PCLPointCloud2();
    if (!input_) {
      std::fprintf(stderr, "[pcl::VoxelGrid<pcl::PCLPointCloud2>::applyFilter] No input dataset given!\n");
      return;
    }
    const int ix = getFieldIndex(*input_, "x"), iy = getFieldIndex(*input_, "y"), iz = getFieldIndex(*input_, "z");
    auto is_float = [&](int i) { return i >= 0 && input_->fields[static_cast<std::size_t>(i)].datatype == PCLPointField::FLOAT32; };
    if (!is_float(ix) || !is_float(iy) || !is_float(iz)) {
      std::fprintf(stderr, "[pcl::VoxelGrid<pcl::PCLPointCloud2>::applyFilter] Input dataset doesn't have x-y-z coordinates!\n");
      return;
    }
    const int inx = getFieldIndex(*input_, "normal_x"), iny = getFieldIndex(*input_, "normal_y"), inz = getFieldIndex(*input_, "normal_z"),
              icv = getFieldIndex(*input_, "curvature");
    used_normals_ = downsample_all_data_ && is_float(inx) && is_float(iny) && is_float(inz) && is_float(icv);
    if (downsample_all_data_) {
      std::size_t carried = used_normals_ ? 7 : 3, named = 0;
      for (const auto& f : input_->fields) named += f.name != "_";
      if (named > carried)
        std::fprintf(stderr, "[pcl::VoxelGrid<pcl::PCLPointCloud2>::applyFilter] %zu field(s) besides x y z%s are not carried into the output.\n",
                     named - carried, used_normals_ ? " normal_x normal_y normal_z curvature" : "");
    }
    const std::size_t n = static_cast<std::size_t>(input_->width) * input_->height;
    auto at = [&](std::size_t i, int field) {
      float v;
      std::memcpy(&v, &input_->data[i * input_->point_step + input_->fields[static_cast<std::size_t>(field)].offset], 4);
      return v;
    };
    auto run = [&](auto& grid, auto& cloud) {
      using PointT = typename std::decay<decltype(cloud->points[0])>::type;
      cloud->header = input_->header;
      cloud->points.resize(n);
      cloud->width = input_->width;
      cloud->height = input_->height;
      cloud->is_dense = input_->is_dense != 0;
      for (std::size_t i = 0; i < n; ++i) {
        PointT& p = cloud->points[i];
        p.x = at(i, ix); p.y = at(i, iy); p.z = at(i, iz);
      }
      grid.setInputCloud(cloud);
      if (indices_) grid.setIndices(indices_);
      pcl::PointCloud<PointT> out;
      grid.filter(out);
      last_indices_ = grid.getIndices();   // the index list the filter ran on (PCLBase::initCompute)
      toPCLPointCloud2(out, output);
      // the output blob holds the downsampled fields only, tightly packed
      PCLPointCloud2 packed;
      packed.header = output.header;
      packed.width = output.width;
      packed.height = output.height;
      packed.is_dense = output.is_dense;
      packed.is_bigendian = 0;
      std::uint32_t off = 0;
      std::vector<std::uint32_t> src_off;
      for (const auto& f : output.fields) {
        if (f.name == "_") continue;
        PCLPointField g = f;
        src_off.push_back(f.offset);
        g.offset = off;
        off += 4;
        packed.fields.push_back(g);
      }
      packed.point_step = off;
      packed.row_step = off * packed.width;
      packed.data.resize(static_cast<std::size_t>(packed.row_step) * packed.height);
      const std::size_t m = static_cast<std::size_t>(output.width) * output.height;
      for (std::size_t i = 0; i < m; ++i)
        for (std::size_t k = 0; k < src_off.size(); ++k)
          std::memcpy(&packed.data[i * off + 4 * k], &output.data[i * output.point_step + src_off[k]], 4);
      output = std::move(packed);
    };
    if (used_normals_) {
      auto cloud = std::make_shared<pcl::PointCloud<PointNormal>>();
      cloud->points.resize(n);
      for (std::size_t i = 0; i < n; ++i) {
        PointNormal& p = cloud->points[i];
        p.normal_x = at(i, inx); p.normal_y = at(i, iny); p.normal_z = at(i, inz); p.curvature = at(i, icv);
      }
      xyzn_.setDownsampleAllData(true);
      run(xyzn_, cloud);
    }
    else {
      auto cloud = std::make_shared<pcl::PointCloud<PointXYZ>>();
      run(xyz_, cloud);
    }
  }

protected:
  PCLPointCloud2ConstPtr input_;
  IndicesPtr indices_, last_indices_;
  bool downsample_all_data_ = true;
  bool used_normals_ = false;
  VoxelGrid<PointXYZ> xyz_;
  VoxelGrid<PointNormal> xyzn_;
};
}  // namespace pcl
