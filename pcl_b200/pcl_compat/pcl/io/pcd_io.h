// pcl/io/pcd_io.h — minimal PCD reader/writer (ASCII and binary, float fields) so the reference's own fixtures
// and tutorials (doc/tutorials/content/sources/iterative_closest_point) run end to end.  io/src/pcd_io.cpp proper
// (compressed PCD, PCLPointCloud2 blobs) is outside the accelerated path.
#pragma once
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {
namespace io {
namespace detail {
template <typename P> inline void setField(P&, const std::string&, float) {}
inline void setXYZ(float* x, float* y, float* z, const std::string& f, float v)
{
  if (f == "x") *x = v; else if (f == "y") *y = v; else if (f == "z") *z = v;
}
inline void setField(PointXYZ& p, const std::string& f, float v) { setXYZ(&p.x, &p.y, &p.z, f, v); }
inline void setField(PointNormal& p, const std::string& f, float v)
{
  setXYZ(&p.x, &p.y, &p.z, f, v);
  if (f == "normal_x") p.normal_x = v; else if (f == "normal_y") p.normal_y = v; else if (f == "normal_z") p.normal_z = v;
  else if (f == "curvature") p.curvature = v;
}
inline void setField(Normal& p, const std::string& f, float v)
{
  if (f == "normal_x") p.normal_x = v; else if (f == "normal_y") p.normal_y = v; else if (f == "normal_z") p.normal_z = v;
  else if (f == "curvature") p.curvature = v;
}
}  // namespace detail

template <typename PointT>
int loadPCDFile(const std::string& file, pcl::PointCloud<PointT>& cloud)
{
  std::ifstream in(file, std::ios::binary);
  if (!in) { std::fprintf(stderr, "[pcl::PCDReader::read] Could not find file '%s'.\n", file.c_str()); return -1; }
  std::vector<std::string> fields;
  std::vector<int> sizes, counts;
  std::vector<char> types;
  std::size_t npts = 0, width = 0, height = 1;
  std::string line, mode;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key == "FIELDS" || key == "COLUMNS") { std::string f; while (ss >> f) fields.push_back(f); }
    else if (key == "SIZE") { int v; while (ss >> v) sizes.push_back(v); }
    else if (key == "TYPE") { char v; while (ss >> v) types.push_back(v); }
    else if (key == "COUNT") { int v; while (ss >> v) counts.push_back(v); }
    else if (key == "WIDTH") ss >> width;
    else if (key == "HEIGHT") ss >> height;
    else if (key == "POINTS") ss >> npts;
    else if (key == "DATA") { ss >> mode; break; }
  }
  if (npts == 0) npts = width * height;
  if (sizes.empty()) sizes.assign(fields.size(), 4);
  if (types.empty()) types.assign(fields.size(), 'F');
  if (counts.empty()) counts.assign(fields.size(), 1);
  cloud.points.assign(npts, PointT());
  cloud.width = static_cast<std::uint32_t>(width ? width : npts);
  cloud.height = static_cast<std::uint32_t>(height);
  cloud.is_dense = true;
  if (mode == "ascii") {
    for (std::size_t i = 0; i < npts; ++i)
      for (std::size_t f = 0; f < fields.size(); ++f)
        for (int c = 0; c < counts[f]; ++c) {
          std::string tok;
          in >> tok;
          float v = std::strtof(tok.c_str(), nullptr);
          if (c == 0) detail::setField(cloud.points[i], fields[f], v);
        }
  }
  else if (mode == "binary") {
    std::size_t rec = 0;
    for (std::size_t f = 0; f < fields.size(); ++f) rec += static_cast<std::size_t>(sizes[f]) * counts[f];
    std::vector<char> buf(rec);
    for (std::size_t i = 0; i < npts; ++i) {
      in.read(buf.data(), static_cast<std::streamsize>(rec));
      std::size_t off = 0;
      for (std::size_t f = 0; f < fields.size(); ++f) {
        if (types[f] == 'F' && sizes[f] == 4) {
          float v;
          std::memcpy(&v, buf.data() + off, 4);
          detail::setField(cloud.points[i], fields[f], v);
        }
        off += static_cast<std::size_t>(sizes[f]) * counts[f];
      }
    }
  }
  else {
    std::fprintf(stderr, "[pcl::PCDReader::read] unsupported DATA mode '%s'\n", mode.c_str());
    return -1;
  }
  for (const auto& p : cloud.points)
    if (!std::isfinite(p.data[0] + p.data[1] + p.data[2])) { cloud.is_dense = false; break; }
  return 0;
}

inline int savePCDFileBinary(const std::string& file, const pcl::PointCloud<PointXYZ>& cloud)
{
  std::ofstream out(file, std::ios::binary);
  if (!out) return -1;
  out << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH "
      << cloud.size() << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << cloud.size() << "\nDATA binary\n";
  for (const auto& p : cloud.points) out.write(reinterpret_cast<const char*>(&p.x), 12);
  return out ? 0 : -1;
}
}  // namespace io
}  // namespace pcl
