// pcl/io/pcd_io.h — PCD reader / writer for the point types of the ICP path: DATA ascii, binary and
// binary_compressed (LZF, structure-of-arrays), every numeric field type, COUNT > 1 and "_" padding fields, the
// VIEWPOINT line, organised (WIDTH x HEIGHT) clouds.
// Reference behaviour: io/src/pcd_io.cpp:120-395 (header), :443-576 (ASCII body), :580-668 (binary and
// binary_compressed bodies, is_dense), io/include/pcl/io/pcd_io.h:633-800 (load/save free functions),
// io/include/pcl/io/impl/pcd_io.hpp:66-130 (header text), :232-428 (writeBinaryCompressed), :430-560 (writeASCII).
// The LZF coder below is written from the published stream format (a control byte < 32 starts a literal run of
// ctrl + 1 bytes; otherwise a back reference of length (ctrl >> 5) + 2 — 7 means "add the next byte" — at distance
// ((ctrl & 31) << 8 | next byte) + 1); any valid stream decodes with the reference's lzfDecompress and vice versa.
// Clouds are read straight into pcl::PointCloud<PointT>, fields matched by name (what PCDReader::read + fromPCLPointCloud2
// amount to); the blob type itself lives in pcl/PCLPointCloud2.h.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../PCLPointCloud2.h"
#include "../common/io.h"
#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {
namespace io {
namespace detail {

// ---- LZF ----------------------------------------------------------------------------------------------------------
// returns the number of bytes written, 0 when `out` is too small
inline std::size_t lzfCompress(const unsigned char* in, std::size_t n, unsigned char* out, std::size_t cap)
{
  constexpr std::size_t kMaxOff = 1u << 13, kMaxLen = 264, kHashBits = 16;
  std::vector<std::int64_t> table(std::size_t(1) << kHashBits, -1);
  std::size_t op = 0, lit = 0;
  auto flush = [&](std::size_t end) -> bool {  // literal runs of at most 32 bytes
    while (lit < end) {
      const std::size_t run = std::min<std::size_t>(32, end - lit);
      if (op + 1 + run > cap) return false;
      out[op++] = static_cast<unsigned char>(run - 1);
      std::memcpy(out + op, in + lit, run);
      op += run;
      lit += run;
    }
    return true;
  };
  auto slot = [&](std::size_t i) -> std::int64_t& {  // hash of the three bytes at i
    const std::uint32_t v = (std::uint32_t(in[i]) << 16) | (std::uint32_t(in[i + 1]) << 8) | in[i + 2];
    return table[((v * 2654435761u) >> (32 - kHashBits)) & ((1u << kHashBits) - 1)];
  };
  std::size_t ip = 0;
  while (ip + 2 < n) {
    std::int64_t& entry = slot(ip);
    const std::int64_t ref = entry;
    entry = static_cast<std::int64_t>(ip);
    if (ref >= 0 && ip - static_cast<std::size_t>(ref) <= kMaxOff && in[ref] == in[ip] && in[ref + 1] == in[ip + 1] &&
        in[ref + 2] == in[ip + 2]) {
      if (!flush(ip)) return 0;
      std::size_t len = 3;
      const std::size_t max_len = std::min<std::size_t>(kMaxLen, n - ip);
      while (len < max_len && in[ref + len] == in[ip + len]) ++len;
      const std::size_t off = ip - static_cast<std::size_t>(ref) - 1, l = len - 2;
      if (op + 3 > cap) return 0;
      if (l < 7)
        out[op++] = static_cast<unsigned char>((l << 5) | (off >> 8));
      else {
        out[op++] = static_cast<unsigned char>((7u << 5) | (off >> 8));
        out[op++] = static_cast<unsigned char>(l - 7);
      }
      out[op++] = static_cast<unsigned char>(off & 0xff);
      for (std::size_t j = ip + 1; j < ip + len && j + 2 < n; ++j) slot(j) = static_cast<std::int64_t>(j);  // keep the table fresh
      ip += len;
      lit = ip;
    }
    else
      ++ip;
  }
  if (!flush(n)) return 0;
  return op;
}

// returns the number of bytes produced, 0 on a corrupt stream or when `out` is too small
inline std::size_t lzfDecompress(const unsigned char* in, std::size_t n, unsigned char* out, std::size_t cap)
{
  std::size_t ip = 0, op = 0;
  while (ip < n) {
    const unsigned ctrl = in[ip++];
    if (ctrl < 32) {
      const std::size_t run = ctrl + 1;
      if (ip + run > n || op + run > cap) return 0;
      std::memcpy(out + op, in + ip, run);
      ip += run;
      op += run;
    }
    else {
      std::size_t len = ctrl >> 5;
      if (len == 7) {
        if (ip >= n) return 0;
        len += in[ip++];
      }
      if (ip >= n) return 0;
      const std::size_t dist = ((std::size_t(ctrl & 0x1f) << 8) | in[ip++]) + 1;
      len += 2;
      if (dist > op || op + len > cap) return 0;
      for (std::size_t i = 0; i < len; ++i, ++op) out[op] = out[op - dist];  // may overlap: byte by byte
    }
  }
  return op;
}

// ---- fields of the supported point types (all FLOAT32, COUNT 1; pcl::getFields<PointT>) ---------------------------
struct FieldDesc {
  const char* name;
  std::size_t offset;
};
template <typename P> struct point_fields;
template <> struct point_fields<PointXYZ> {
  static std::vector<FieldDesc> get() { return {{"x", 0}, {"y", 4}, {"z", 8}}; }
};
template <> struct point_fields<Normal> {
  static std::vector<FieldDesc> get() { return {{"normal_x", 0}, {"normal_y", 4}, {"normal_z", 8}, {"curvature", 16}}; }
};
template <> struct point_fields<PointNormal> {
  static std::vector<FieldDesc> get()
  {
    return {{"x", 0}, {"y", 4}, {"z", 8}, {"normal_x", 16}, {"normal_y", 20}, {"normal_z", 24}, {"curvature", 32}};
  }
};

struct FileField {
  std::string name;
  int size = 4;
  char type = 'F';
  int count = 1;
  std::size_t offset = 0;  // inside one binary record
};

struct Header {
  std::vector<FileField> fields;
  std::size_t width = 0, height = 1, points = 0, point_step = 0;
  bool width_read = false, height_read = false;
  float viewpoint[7] = {0, 0, 0, 1, 0, 0, 0};
  int data_type = -1;  // 0 ascii, 1 binary, 2 binary_compressed
  std::size_t data_offset = 0;
};

inline double readValue(const unsigned char* p, char type, int size, bool* finite)
{
  *finite = true;
  switch (type) {
    case 'F':
      if (size == 4) { float v; std::memcpy(&v, p, 4); *finite = std::isfinite(v); return v; }
      if (size == 8) { double v; std::memcpy(&v, p, 8); *finite = std::isfinite(v); return v; }
      break;
    case 'I':
      if (size == 1) { std::int8_t v; std::memcpy(&v, p, 1); return v; }
      if (size == 2) { std::int16_t v; std::memcpy(&v, p, 2); return v; }
      if (size == 4) { std::int32_t v; std::memcpy(&v, p, 4); return v; }
      if (size == 8) { std::int64_t v; std::memcpy(&v, p, 8); return static_cast<double>(v); }
      break;
    case 'U':
      if (size == 1) { std::uint8_t v; std::memcpy(&v, p, 1); return v; }
      if (size == 2) { std::uint16_t v; std::memcpy(&v, p, 2); return v; }
      if (size == 4) { std::uint32_t v; std::memcpy(&v, p, 4); return v; }
      if (size == 8) { std::uint64_t v; std::memcpy(&v, p, 8); return static_cast<double>(v); }
      break;
    default: break;
  }
  return 0.0;
}

// io/src/pcd_io.cpp:120-395
inline int readHeader(std::istream& fs, Header& h)
{
  std::string line;
  bool sizes_read = false, types_read = false;
  auto fail = [](const char* msg) {
    std::fprintf(stderr, "[pcl::PCDReader::readHeader] %s\n", msg);
    return -1;
  };
  auto finish_offsets = [&]() {
    std::size_t off = 0;
    for (auto& f : h.fields) {
      f.offset = off;
      off += static_cast<std::size_t>(f.size) * static_cast<std::size_t>(std::max(f.count, 0));
    }
    h.point_step = off;
  };
  bool points_read = false;
  while (std::getline(fs, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key.empty() || key[0] == '#') continue;
    std::vector<std::string> tok;
    for (std::string t; ss >> t;) tok.push_back(t);
    if (key.compare(0, 7, "VERSION") == 0) continue;
    if (key.compare(0, 6, "FIELDS") == 0 || key.compare(0, 7, "COLUMNS") == 0) {
      h.fields.assign(tok.size(), FileField());
      for (std::size_t i = 0; i < tok.size(); ++i) h.fields[i].name = tok[i];
      finish_offsets();  // float32 / count 1 until SIZE, TYPE, COUNT say otherwise (:176-185)
      continue;
    }
    if (key.compare(0, 4, "SIZE") == 0) {
      if (tok.size() != h.fields.size()) return fail("The number of elements in <SIZE> differs than the number of elements in <FIELDS>!");
      for (std::size_t i = 0; i < tok.size(); ++i) h.fields[i].size = std::atoi(tok[i].c_str());
      sizes_read = true;
      finish_offsets();
      continue;
    }
    if (key.compare(0, 4, "TYPE") == 0) {
      if (!sizes_read) return fail("TYPE of FIELDS specified before SIZE in header!");
      if (tok.size() != h.fields.size()) return fail("The number of elements in <TYPE> differs than the number of elements in <FIELDS>!");
      for (std::size_t i = 0; i < tok.size(); ++i) h.fields[i].type = tok[i][0];
      types_read = true;
      continue;
    }
    if (key.compare(0, 5, "COUNT") == 0) {
      if (!sizes_read || !types_read) return fail("COUNT of FIELDS specified before SIZE or TYPE in header!");
      if (tok.size() != h.fields.size()) return fail("The number of elements in <COUNT> differs than the number of elements in <FIELDS>!");
      for (std::size_t i = 0; i < tok.size(); ++i) h.fields[i].count = std::atoi(tok[i].c_str());
      finish_offsets();
      continue;
    }
    if (key.compare(0, 5, "WIDTH") == 0) {
      if (tok.empty()) return fail("Invalid WIDTH value specified.");
      h.width = std::strtoull(tok[0].c_str(), nullptr, 10);
      h.width_read = true;
      continue;
    }
    if (key.compare(0, 6, "HEIGHT") == 0) {
      if (tok.empty()) return fail("Invalid HEIGHT value specified.");
      h.height = std::strtoull(tok[0].c_str(), nullptr, 10);
      h.height_read = true;
      continue;
    }
    if (key.compare(0, 9, "VIEWPOINT") == 0) {
      if (tok.size() < 7) return fail("Not enough number of elements in <VIEWPOINT>! Need 7 values (tx ty tz qw qx qy qz).");
      for (int i = 0; i < 7; ++i) h.viewpoint[i] = std::strtof(tok[i].c_str(), nullptr);
      continue;
    }
    if (key.compare(0, 6, "POINTS") == 0) {
      if (!h.point_step) return fail("Number of POINTS specified before COUNT in header!");
      if (tok.empty()) return fail("Invalid POINTS value specified.");
      h.points = std::strtoull(tok[0].c_str(), nullptr, 10);
      points_read = true;
      continue;
    }
    if (key.compare(0, 4, "DATA") == 0) {
      if (tok.empty()) return fail("Unknown DATA format");
      if (tok[0].compare(0, 17, "binary_compressed") == 0) h.data_type = 2;
      else if (tok[0].compare(0, 6, "binary") == 0) h.data_type = 1;
      else if (tok[0].compare(0, 5, "ascii") == 0) h.data_type = 0;
      else return fail("Unknown DATA format");
      h.data_offset = static_cast<std::size_t>(fs.tellg());
      break;  // DATA is the last header entry
    }
  }
  if (h.data_type < 0) return fail("no DATA line");
  (void)points_read;
  // An untrusted header must not be able to wrap an offset or request an absurd allocation: sizes are the four PCD
  // allows, counts are positive, the type fits its size (io/src/pcd_io.cpp:200-260 rejects the same things), and every
  // product below is checked for overflow.
  if (h.fields.empty()) return fail("no FIELDS");
  for (const auto& f : h.fields) {
    if (!(f.size == 1 || f.size == 2 || f.size == 4 || f.size == 8)) return fail("invalid SIZE (must be 1, 2, 4 or 8)");
    if (f.count < 1 || f.count > (1 << 20)) return fail("invalid COUNT (must be >= 1)");
    if (!(f.type == 'F' || f.type == 'I' || f.type == 'U')) return fail("invalid TYPE (must be F, I or U)");
    if (f.type == 'F' && f.size < 4) return fail("TYPE F needs SIZE 4 or 8");
  }
  if (h.point_step == 0 || h.point_step > (std::size_t(1) << 30)) return fail("invalid point step");
  if (h.width != 0 && h.height > std::numeric_limits<std::size_t>::max() / h.width) return fail("WIDTH x HEIGHT overflows");
  if (h.points > std::numeric_limits<std::size_t>::max() / h.point_step) return fail("POINTS x point step overflows");
  // compatibility with older files (:351-383)
  if (!h.width_read && !h.height_read) { h.width = h.points; h.height = 1; }
  if (!h.height_read) { h.height = 1; if (h.width == 0) h.width = h.points; }
  else if (h.width == 0 && h.points != 0) return fail("HEIGHT given but no WIDTH!");
  if (h.points == 0) h.points = h.width * h.height;
  if (h.width * h.height != h.points) return fail("HEIGHT x WIDTH != number of points");
  return 0;
}

// header text of the writers — io/include/pcl/io/impl/pcd_io.hpp:66-130
inline std::string headerText(const std::vector<FieldDesc>& fields, std::size_t width, std::size_t height, std::size_t n,
                              const float vp[7], const char* data)
{
  std::ostringstream os;
  os << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS";
  for (const auto& f : fields) os << ' ' << f.name;
  os << "\nSIZE";
  for (std::size_t i = 0; i < fields.size(); ++i) os << " 4";
  os << "\nTYPE";
  for (std::size_t i = 0; i < fields.size(); ++i) os << " F";
  os << "\nCOUNT";
  for (std::size_t i = 0; i < fields.size(); ++i) os << " 1";
  os << "\nWIDTH " << width << "\nHEIGHT " << height << "\nVIEWPOINT " << vp[0] << ' ' << vp[1] << ' ' << vp[2] << ' ' << vp[3]
     << ' ' << vp[4] << ' ' << vp[5] << ' ' << vp[6] << "\nPOINTS " << n << "\nDATA " << data << "\n";
  return os.str();
}

template <typename PointT> inline void writerGeometry(const pcl::PointCloud<PointT>& cloud, std::size_t& w, std::size_t& h, float vp[7])
{
  const std::size_t n = cloud.size();
  w = cloud.width;
  h = cloud.height;
  if (w * h != n) { w = n; h = 1; }  // pcd_io.hpp:76-84: an inconsistent cloud is written as unorganised
  vp[0] = cloud.sensor_origin_[0]; vp[1] = cloud.sensor_origin_[1]; vp[2] = cloud.sensor_origin_[2];
  vp[3] = 1.f; vp[4] = vp[5] = vp[6] = 0.f;
}
}  // namespace detail

// pcl::io::loadPCDFile<PointT> — io/include/pcl/io/pcd_io.h:662-667 (PCDReader::read + fromPCLPointCloud2)
template <typename PointT>
int loadPCDFile(const std::string& file, pcl::PointCloud<PointT>& cloud)
{
  std::ifstream in(file, std::ios::binary);
  if (!in) { std::fprintf(stderr, "[pcl::PCDReader::read] Could not find file '%s'.\n", file.c_str()); return -1; }
  detail::Header h;
  if (detail::readHeader(in, h) != 0) return -1;
  const std::size_t npts = h.points;
  {
    // a point needs at least one byte of file per field in every encoding (ascii: a digit; binary: >= 1 byte;
    // compressed: checked against the declared uncompressed size below) — refuse before allocating
    const std::streampos here = in.tellg();
    in.seekg(0, std::ios::end);
    const std::size_t fsz = static_cast<std::size_t>(in.tellg());
    in.seekg(here);
    const std::size_t body = fsz > h.data_offset ? fsz - h.data_offset : 0;
    // compressed: the body is u32 csize, u32 usize, csize bytes; usize must be the planes' size and an LZF stream cannot
    // expand more than ~132x, so the point count is bounded by the FILE size before anything is allocated
    const bool too_many = h.data_type == 1 ? npts * h.point_step > body
                          : h.data_type == 0 ? npts > body
                                             : (body < 8 || npts > (body - 8) * 256 + 64);
    if (too_many) {
      std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file: %zu points do not fit %zu bytes of data.\n", npts, body);
      return -1;
    }
  }
  cloud.points.assign(npts, PointT());
  cloud.width = static_cast<std::uint32_t>(h.width);
  cloud.height = static_cast<std::uint32_t>(h.height);
  cloud.is_dense = true;
  cloud.sensor_origin_[0] = h.viewpoint[0];
  cloud.sensor_origin_[1] = h.viewpoint[1];
  cloud.sensor_origin_[2] = h.viewpoint[2];
  cloud.sensor_origin_[3] = 0.f;
  // file field -> byte offset inside PointT (or -1: not part of this point type, skipped like fromPCLPointCloud2 does)
  const auto want = detail::point_fields<PointT>::get();
  std::vector<std::ptrdiff_t> dst(h.fields.size(), -1);
  std::size_t matched = 0;
  for (std::size_t f = 0; f < h.fields.size(); ++f)
    for (const auto& w : want)
      if (h.fields[f].name == w.name && h.fields[f].count >= 1) { dst[f] = static_cast<std::ptrdiff_t>(w.offset); ++matched; }
  if (matched < want.size())
    std::fprintf(stderr, "[pcl::PCDReader::read] Failed to find match for some fields of the point type in '%s'.\n", file.c_str());
  auto store = [&](std::size_t i, std::size_t f, double v) {
    const float fv = static_cast<float>(v);
    std::memcpy(reinterpret_cast<unsigned char*>(&cloud.points[i]) + dst[f], &fv, 4);
  };
  if (h.data_type == 0) {  // io/src/pcd_io.cpp:443-576
    std::size_t per_line = 0;
    for (const auto& f : h.fields) per_line += static_cast<std::size_t>(std::max(f.count, 0));
    std::string line;
    std::size_t i = 0;
    std::vector<std::string> tok;
    while (i < npts && std::getline(in, line)) {
      tok.clear();
      std::istringstream ss(line);
      for (std::string t; ss >> t;) tok.push_back(t);
      if (tok.empty()) continue;
      if (tok.size() != per_line) {  // :487-493: malformed line, the point keeps its defaults
        std::fprintf(stderr, "[pcl::PCDReader::readBodyASCII] Possibly malformed PCD file: point number %zu has %zu elements, but should have %zu\n",
                     i + 1, tok.size(), per_line);
        ++i;
        continue;
      }
      std::size_t t = 0;
      for (std::size_t f = 0; f < h.fields.size(); ++f) {
        const int cnt = std::max(h.fields[f].count, 0);
        if (h.fields[f].name != "_" && cnt > 0) {
          for (int c = 0; c < cnt; ++c) {
            const double v = std::strtod(tok[t + c].c_str(), nullptr);
            if (!std::isfinite(v)) cloud.is_dense = false;
            if (c == 0 && dst[f] >= 0) store(i, f, v);
          }
        }
        t += static_cast<std::size_t>(cnt);
      }
      ++i;
    }
    if (i != npts) {
      std::fprintf(stderr, "[pcl::PCDReader::read] Number of points read (%zu) is different than expected (%zu)\n", i, npts);
      return -1;
    }
    return 0;
  }
  // binary bodies — io/src/pcd_io.cpp:580-668
  in.seekg(0, std::ios::end);
  const std::size_t file_size = static_cast<std::size_t>(in.tellg());
  in.seekg(static_cast<std::streamoff>(h.data_offset));
  if (h.data_type == 1) {
    const std::size_t bytes = npts * h.point_step;
    if (h.data_offset + bytes > file_size) {
      std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file. The file is smaller than expected!\n");
      return -1;
    }
    std::vector<unsigned char> buf(bytes ? bytes : 1);
    in.read(reinterpret_cast<char*>(buf.data()), static_cast<std::streamsize>(bytes));
    for (std::size_t i = 0; i < npts; ++i)
      for (std::size_t f = 0; f < h.fields.size(); ++f) {
        const auto& ff = h.fields[f];
        for (int c = 0; c < ff.count; ++c) {
          bool fin;
          const double v = detail::readValue(buf.data() + i * h.point_step + ff.offset + static_cast<std::size_t>(c) * ff.size, ff.type, ff.size, &fin);
          if (!fin) cloud.is_dense = false;
          if (c == 0 && dst[f] >= 0) store(i, f, v);
        }
      }
    return 0;
  }
  // binary_compressed: u32 compressed size, u32 uncompressed size, LZF stream of the fields as planes (all x, all y, …),
  // "_" padding fields not stored (:587-631)
  if (h.data_offset + 8 > file_size) { std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file.\n"); return -1; }
  std::uint32_t csize = 0, usize = 0;
  in.read(reinterpret_cast<char*>(&csize), 4);
  in.read(reinterpret_cast<char*>(&usize), 4);
  if (h.data_offset + 8 + csize > file_size) {
    std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file. The file is smaller than expected!\n");
    return -1;
  }
  std::size_t plane_bytes = 0;
  for (const auto& f : h.fields)
    if (f.name != "_") plane_bytes += static_cast<std::size_t>(f.size) * static_cast<std::size_t>(std::max(f.count, 0));
  if (static_cast<std::size_t>(usize) != plane_bytes * npts)
    std::fprintf(stderr, "[pcl::PCDReader::read] The estimated cloud.data size (%zu) is different than the saved uncompressed value (%u)! Data corruption?\n",
                 plane_bytes * npts, usize);
  if (usize == 0) return 0;
  // an LZF back-reference of 2-3 bytes yields at most 264: a stream cannot expand by more than ~132x
  if (static_cast<std::uint64_t>(usize) > static_cast<std::uint64_t>(csize) * 256u + 64u) {
    std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file: %u compressed bytes cannot hold %u uncompressed.\n", csize, usize);
    return -1;
  }
  std::vector<unsigned char> cbuf(csize ? csize : 1), buf(usize);
  in.read(reinterpret_cast<char*>(cbuf.data()), static_cast<std::streamsize>(csize));
  const std::size_t got = detail::lzfDecompress(cbuf.data(), csize, buf.data(), usize);
  if (got != usize) {
    std::fprintf(stderr, "[pcl::PCDReader::read] Size of decompressed lzf data (%zu) does not match value stored in PCD header (%u).\n", got, usize);
    return -1;
  }
  if (plane_bytes * npts > usize) return -1;
  std::size_t plane = 0;
  for (std::size_t f = 0; f < h.fields.size(); ++f) {
    const auto& ff = h.fields[f];
    if (ff.name == "_" || ff.count < 1) continue;
    const std::size_t fs = static_cast<std::size_t>(ff.size) * static_cast<std::size_t>(ff.count);
    for (std::size_t i = 0; i < npts; ++i)
      for (int c = 0; c < ff.count; ++c) {
        bool fin;
        const double v = detail::readValue(buf.data() + plane + i * fs + static_cast<std::size_t>(c) * ff.size, ff.type, ff.size, &fin);
        if (!fin) cloud.is_dense = false;
        if (c == 0 && dst[f] >= 0) store(i, f, v);
      }
    plane += fs * npts;
  }
  return 0;
}

// PCDWriter::writeASCII — io/include/pcl/io/impl/pcd_io.hpp:430-560 (precision 8, NaN written as "nan")
template <typename PointT>
int savePCDFileASCII(const std::string& file, const pcl::PointCloud<PointT>& cloud, int precision = 8)
{
  if (cloud.empty()) std::fprintf(stderr, "[pcl::PCDWriter::writeASCII] Input point cloud has no data!\n");
  std::ofstream out(file, std::ios::binary);
  if (!out) { std::fprintf(stderr, "[pcl::PCDWriter::writeASCII] Could not open file for writing!\n"); return -1; }
  const auto fields = detail::point_fields<PointT>::get();
  std::size_t w, h;
  float vp[7];
  detail::writerGeometry(cloud, w, h, vp);
  out << detail::headerText(fields, w, h, cloud.size(), vp, "ascii");
  out.precision(precision);
  for (const auto& p : cloud.points) {
    for (std::size_t f = 0; f < fields.size(); ++f) {
      float v;
      std::memcpy(&v, reinterpret_cast<const unsigned char*>(&p) + fields[f].offset, 4);
      if (f) out << ' ';
      if (std::isnan(v)) out << "nan";
      else out << v;
    }
    out << '\n';
  }
  return out ? 0 : -1;
}

// PCDWriter::writeBinary — io/include/pcl/io/impl/pcd_io.hpp:132-230: records of the point type's fields, no padding
template <typename PointT>
int savePCDFileBinary(const std::string& file, const pcl::PointCloud<PointT>& cloud)
{
  if (cloud.empty()) std::fprintf(stderr, "[pcl::PCDWriter::writeBinary] Input point cloud has no data!\n");
  std::ofstream out(file, std::ios::binary);
  if (!out) { std::fprintf(stderr, "[pcl::PCDWriter::writeBinary] Could not open file for writing!\n"); return -1; }
  const auto fields = detail::point_fields<PointT>::get();
  std::size_t w, h;
  float vp[7];
  detail::writerGeometry(cloud, w, h, vp);
  out << detail::headerText(fields, w, h, cloud.size(), vp, "binary");
  std::vector<unsigned char> buf(cloud.size() * fields.size() * 4 + 1);
  std::size_t o = 0;
  for (const auto& p : cloud.points)
    for (const auto& f : fields) {
      std::memcpy(buf.data() + o, reinterpret_cast<const unsigned char*>(&p) + f.offset, 4);
      o += 4;
    }
  out.write(reinterpret_cast<const char*>(buf.data()), static_cast<std::streamsize>(o));
  return out ? 0 : -1;
}

// PCDWriter::writeBinaryCompressed — io/include/pcl/io/impl/pcd_io.hpp:232-428: planes per field, LZF, two u32 sizes
template <typename PointT>
int savePCDFileBinaryCompressed(const std::string& file, const pcl::PointCloud<PointT>& cloud)
{
  if (cloud.empty()) std::fprintf(stderr, "[pcl::PCDWriter::writeBinaryCompressed] Input point cloud has no data!\n");
  const auto fields = detail::point_fields<PointT>::get();
  const std::size_t n = cloud.size(), data_size = n * fields.size() * 4;
  if (data_size * 3 / 2 > std::numeric_limits<std::uint32_t>::max()) {  // :296-300
    std::fprintf(stderr, "[pcl::PCDWriter::writeBinaryCompressed] The input data exceeds the maximum size for compressed version 0.7 pcds.\n");
    return -2;
  }
  std::ofstream out(file, std::ios::binary);
  if (!out) { std::fprintf(stderr, "[pcl::PCDWriter::writeBinaryCompressed] Could not open file for writing!\n"); return -1; }
  std::size_t w, h;
  float vp[7];
  detail::writerGeometry(cloud, w, h, vp);
  out << detail::headerText(fields, w, h, n, vp, "binary_compressed");
  std::vector<unsigned char> planes(data_size ? data_size : 1);
  for (std::size_t f = 0; f < fields.size(); ++f)
    for (std::size_t i = 0; i < n; ++i)
      std::memcpy(planes.data() + (f * n + i) * 4, reinterpret_cast<const unsigned char*>(&cloud.points[i]) + fields[f].offset, 4);
  std::vector<unsigned char> comp(data_size + data_size / 16 + 64);
  std::uint32_t csize = 0;
  const std::uint32_t usize = static_cast<std::uint32_t>(data_size);
  if (data_size) {
    csize = static_cast<std::uint32_t>(detail::lzfCompress(planes.data(), data_size, comp.data(), comp.size()));
    if (csize == 0) { std::fprintf(stderr, "[pcl::PCDWriter::writeBinaryCompressed] Error during compression!\n"); return -1; }
  }
  out.write(reinterpret_cast<const char*>(&csize), 4);
  out.write(reinterpret_cast<const char*>(&usize), 4);
  out.write(reinterpret_cast<const char*>(comp.data()), csize);
  return out ? 0 : -1;
}

// io/include/pcl/io/pcd_io.h:708-713
template <typename PointT>
int savePCDFile(const std::string& file, const pcl::PointCloud<PointT>& cloud, bool binary_mode = false)
{
  return binary_mode ? savePCDFileBinary(file, cloud) : savePCDFileASCII(file, cloud);
}
// ---- the type-erased route (PCDReader::read / PCDWriter::write* on pcl::PCLPointCloud2, io/src/pcd_io.cpp) -------------
// The blob keeps the file's own record layout: every named field with its datatype, count and byte offset, "_" padding
// skipped in the field list but kept in point_step; origin / orientation come from the VIEWPOINT line.
namespace detail {
inline bool blobFieldFinite(const unsigned char* p, std::uint8_t datatype)
{
  if (datatype == PCLPointField::FLOAT32) { float v; std::memcpy(&v, p, 4); return std::isfinite(v); }
  if (datatype == PCLPointField::FLOAT64) { double v; std::memcpy(&v, p, 8); return std::isfinite(v); }
  return true;
}
inline void storeAscii(const std::string& tok, std::uint8_t datatype, unsigned char* dst, bool* finite)
{
  *finite = true;
  switch (datatype) {
    case PCLPointField::FLOAT32: { const float v = std::strtof(tok.c_str(), nullptr); *finite = std::isfinite(v); std::memcpy(dst, &v, 4); break; }
    case PCLPointField::FLOAT64: { const double v = std::strtod(tok.c_str(), nullptr); *finite = std::isfinite(v); std::memcpy(dst, &v, 8); break; }
    case PCLPointField::INT8: { const std::int8_t v = static_cast<std::int8_t>(std::strtol(tok.c_str(), nullptr, 10)); std::memcpy(dst, &v, 1); break; }
    case PCLPointField::UINT8: { const std::uint8_t v = static_cast<std::uint8_t>(std::strtoul(tok.c_str(), nullptr, 10)); std::memcpy(dst, &v, 1); break; }
    case PCLPointField::INT16: { const std::int16_t v = static_cast<std::int16_t>(std::strtol(tok.c_str(), nullptr, 10)); std::memcpy(dst, &v, 2); break; }
    case PCLPointField::UINT16: { const std::uint16_t v = static_cast<std::uint16_t>(std::strtoul(tok.c_str(), nullptr, 10)); std::memcpy(dst, &v, 2); break; }
    case PCLPointField::INT32: { const std::int32_t v = static_cast<std::int32_t>(std::strtol(tok.c_str(), nullptr, 10)); std::memcpy(dst, &v, 4); break; }
    case PCLPointField::UINT32: { const std::uint32_t v = static_cast<std::uint32_t>(std::strtoul(tok.c_str(), nullptr, 10)); std::memcpy(dst, &v, 4); break; }
    default: break;
  }
}
inline void printAscii(std::ostream& os, const unsigned char* p, std::uint8_t datatype)
{
  switch (datatype) {
    case PCLPointField::FLOAT32: { float v; std::memcpy(&v, p, 4); if (std::isnan(v)) os << "nan"; else os << v; break; }
    case PCLPointField::FLOAT64: { double v; std::memcpy(&v, p, 8); if (std::isnan(v)) os << "nan"; else os << v; break; }
    case PCLPointField::INT8: { std::int8_t v; std::memcpy(&v, p, 1); os << static_cast<int>(v); break; }
    case PCLPointField::UINT8: { std::uint8_t v; std::memcpy(&v, p, 1); os << static_cast<unsigned>(v); break; }
    case PCLPointField::INT16: { std::int16_t v; std::memcpy(&v, p, 2); os << v; break; }
    case PCLPointField::UINT16: { std::uint16_t v; std::memcpy(&v, p, 2); os << v; break; }
    case PCLPointField::INT32: { std::int32_t v; std::memcpy(&v, p, 4); os << v; break; }
    case PCLPointField::UINT32: { std::uint32_t v; std::memcpy(&v, p, 4); os << v; break; }
    default: break;
  }
}
// named fields in record order
inline std::vector<PCLPointField> fieldsByOffset(const pcl::PCLPointCloud2& cloud)
{
  std::vector<PCLPointField> f;
  for (const auto& x : cloud.fields)
    if (x.name != "_") f.push_back(x);
  std::stable_sort(f.begin(), f.end(), [](const PCLPointField& a, const PCLPointField& b) { return a.offset < b.offset; });
  return f;
}
inline std::uint32_t fieldCount(const PCLPointField& f) { return f.count ? f.count : 1u; }  // 0 counts of old converters mean 1
// header of a blob; with_padding: "_" byte fields fill the gaps so that the records can be written verbatim (binary)
inline std::string blobHeader(const pcl::PCLPointCloud2& cloud, const Eigen::Vector4f& origin, const Eigen::Quaternionf& orientation,
                              bool with_padding, const char* data)
{
  std::ostringstream names, sizes, types, counts, os;
  std::uint32_t at = 0;
  for (const auto& f : fieldsByOffset(cloud)) {
    if (with_padding && f.offset > at) {
      names << " _"; sizes << " 1"; types << " U"; counts << ' ' << (f.offset - at);
      at = f.offset;
    }
    names << ' ' << f.name;
    sizes << ' ' << getFieldSize(f.datatype);
    types << ' ' << getFieldType(static_cast<int>(f.datatype));
    counts << ' ' << fieldCount(f);
    at += fieldCount(f) * static_cast<std::uint32_t>(getFieldSize(f.datatype));
  }
  if (with_padding && at < cloud.point_step) {
    names << " _"; sizes << " 1"; types << " U"; counts << ' ' << (cloud.point_step - at);
  }
  os << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS" << names.str() << "\nSIZE" << sizes.str() << "\nTYPE"
     << types.str() << "\nCOUNT" << counts.str() << "\nWIDTH " << cloud.width << "\nHEIGHT " << cloud.height << "\nVIEWPOINT "
     << origin[0] << ' ' << origin[1] << ' ' << origin[2] << ' ' << orientation.w() << ' ' << orientation.x() << ' '
     << orientation.y() << ' ' << orientation.z() << "\nPOINTS " << static_cast<std::size_t>(cloud.width) * cloud.height << "\nDATA "
     << data << "\n";
  return os.str();
}
inline bool blobWritable(const pcl::PCLPointCloud2& cloud, const char* who)
{
  if (cloud.fields.empty()) {
    std::fprintf(stderr, "[pcl::PCDWriter::%s] Input point cloud has no field data!\n", who);
    return false;
  }
  const std::size_t npts = static_cast<std::size_t>(cloud.width) * cloud.height;
  if (cloud.data.size() < npts * cloud.point_step) {
    std::fprintf(stderr, "[pcl::PCDWriter::%s] The blob holds fewer bytes than width x height x point_step!\n", who);
    return false;
  }
  if (npts == 0) return true;  // an empty blob (the reference writes it: test_io.cpp:737-790): no record is ever read, only the header goes out
  for (const auto& f : cloud.fields)
    if (f.name != "_" && (getFieldSize(f.datatype) == 0 ||
                          f.offset + fieldCount(f) * static_cast<std::uint32_t>(getFieldSize(f.datatype)) > cloud.point_step)) {
      std::fprintf(stderr, "[pcl::PCDWriter::%s] Field '%s' does not fit the point step!\n", who, f.name.c_str());
      return false;
    }
  return true;
}
}  // namespace detail

// PCDReader::read (file, blob, origin, orientation) — io/src/pcd_io.cpp:120-395, 443-700
inline int loadPCDFile(const std::string& file, pcl::PCLPointCloud2& cloud, Eigen::Vector4f& origin, Eigen::Quaternionf& orientation)
{
  std::ifstream in(file, std::ios::binary);
  if (!in) { std::fprintf(stderr, "[pcl::PCDReader::read] Could not find file '%s'.\n", file.c_str()); return -1; }
  detail::Header h;
  if (detail::readHeader(in, h) != 0) return -1;
  origin[0] = h.viewpoint[0]; origin[1] = h.viewpoint[1]; origin[2] = h.viewpoint[2]; origin[3] = 0.f;
  orientation = Eigen::Quaternionf(h.viewpoint[3], h.viewpoint[4], h.viewpoint[5], h.viewpoint[6]);
  cloud = pcl::PCLPointCloud2();
  for (const auto& f : h.fields) {
    if (f.name == "_") continue;
    const int dt = getFieldType(f.size, f.type);
    if (dt < 0) {
      std::fprintf(stderr, "[pcl::PCDReader::read] Field '%s' has a SIZE / TYPE combination (%d, %c) a PCLPointField cannot hold.\n",
                   f.name.c_str(), f.size, f.type);
      return -1;
    }
    PCLPointField pf;
    pf.name = f.name;
    pf.offset = static_cast<std::uint32_t>(f.offset);
    pf.datatype = static_cast<std::uint8_t>(dt);
    pf.count = static_cast<std::uint32_t>(f.count);
    cloud.fields.push_back(pf);
  }
  const std::size_t npts = h.points;
  in.seekg(0, std::ios::end);
  const std::size_t file_size = static_cast<std::size_t>(in.tellg());
  in.seekg(static_cast<std::streamoff>(h.data_offset));
  const std::size_t body = file_size > h.data_offset ? file_size - h.data_offset : 0;
  if ((h.data_type == 1 && npts * h.point_step > body) || (h.data_type == 0 && npts > body) ||
      (h.data_type == 2 && (body < 8 || npts > (body - 8) * 256 + 64))) {   // an LZF stream expands < 256x: bounded by the file
    std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file: %zu points do not fit %zu bytes of data.\n", npts, body);
    return -1;
  }
  cloud.width = static_cast<std::uint32_t>(h.width);
  cloud.height = static_cast<std::uint32_t>(h.height);
  cloud.point_step = static_cast<std::uint32_t>(h.point_step);
  cloud.row_step = cloud.point_step * cloud.width;
  cloud.is_bigendian = 0;
  cloud.is_dense = 1;
  cloud.data.assign(npts * h.point_step, 0);
  if (h.data_type == 0) {
    std::size_t per_line = 0;
    for (const auto& f : h.fields) per_line += static_cast<std::size_t>(f.count);
    std::string line;
    std::vector<std::string> tok;
    std::size_t i = 0;
    while (i < npts && std::getline(in, line)) {
      tok.clear();
      std::istringstream ss(line);
      for (std::string t; ss >> t;) tok.push_back(t);
      if (tok.empty()) continue;
      if (tok.size() != per_line) {
        std::fprintf(stderr, "[pcl::PCDReader::readBodyASCII] Possibly malformed PCD file: point number %zu has %zu elements, but should have %zu\n",
                     i + 1, tok.size(), per_line);
        ++i;
        continue;
      }
      std::size_t t = 0;
      for (const auto& f : h.fields) {
        if (f.name != "_") {
          const int dt = getFieldType(f.size, f.type);
          for (int c = 0; c < f.count; ++c) {
            bool fin;
            detail::storeAscii(tok[t + c], static_cast<std::uint8_t>(dt), cloud.data.data() + i * h.point_step + f.offset + static_cast<std::size_t>(c) * f.size, &fin);
            if (!fin) cloud.is_dense = 0;
          }
        }
        t += static_cast<std::size_t>(f.count);
      }
      ++i;
    }
    if (i != npts) {
      std::fprintf(stderr, "[pcl::PCDReader::read] Number of points read (%zu) is different than expected (%zu)\n", i, npts);
      return -1;
    }
    return 0;
  }
  if (h.data_type == 1) {
    in.read(reinterpret_cast<char*>(cloud.data.data()), static_cast<std::streamsize>(cloud.data.size()));
    if (!in && !cloud.data.empty()) { std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file. The file is smaller than expected!\n"); return -1; }
  }
  else {
    if (h.data_offset + 8 > file_size) { std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file.\n"); return -1; }
    std::uint32_t csize = 0, usize = 0;
    in.read(reinterpret_cast<char*>(&csize), 4);
    in.read(reinterpret_cast<char*>(&usize), 4);
    if (h.data_offset + 8 + csize > file_size) {
      std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file. The file is smaller than expected!\n");
      return -1;
    }
    std::size_t plane_bytes = 0;
    for (const auto& f : h.fields)
      if (f.name != "_") plane_bytes += static_cast<std::size_t>(f.size) * static_cast<std::size_t>(f.count);
    if (static_cast<std::size_t>(usize) != plane_bytes * npts) {
      std::fprintf(stderr, "[pcl::PCDReader::read] The estimated cloud.data size (%zu) is different than the saved uncompressed value (%u)! Data corruption?\n",
                   plane_bytes * npts, usize);
      return -1;
    }
    if (usize) {
      if (static_cast<std::uint64_t>(usize) > static_cast<std::uint64_t>(csize) * 256u + 64u) {
        std::fprintf(stderr, "[pcl::PCDReader::read] Corrupted PCD file: %u compressed bytes cannot hold %u uncompressed.\n", csize, usize);
        return -1;
      }
      std::vector<unsigned char> cbuf(csize ? csize : 1), buf(usize);
      in.read(reinterpret_cast<char*>(cbuf.data()), static_cast<std::streamsize>(csize));
      if (detail::lzfDecompress(cbuf.data(), csize, buf.data(), usize) != usize) {
        std::fprintf(stderr, "[pcl::PCDReader::read] Size of decompressed lzf data does not match value stored in PCD header (%u).\n", usize);
        return -1;
      }
      std::size_t plane = 0;   // planes (all values of field 0, then of field 1, ...) -> records
      for (const auto& f : h.fields) {
        if (f.name == "_") continue;
        const std::size_t fs = static_cast<std::size_t>(f.size) * static_cast<std::size_t>(f.count);
        for (std::size_t i = 0; i < npts; ++i)
          std::memcpy(cloud.data.data() + i * h.point_step + f.offset, buf.data() + plane + i * fs, fs);
        plane += fs * npts;
      }
    }
  }
  for (std::size_t i = 0; i < npts && cloud.is_dense; ++i)      // io/src/pcd_io.cpp:668-700: a non-finite float clears is_dense
    for (const auto& f : cloud.fields)
      for (std::uint32_t c = 0; c < f.count; ++c)
        if (!detail::blobFieldFinite(cloud.data.data() + i * h.point_step + f.offset + c * static_cast<std::uint32_t>(getFieldSize(f.datatype)), f.datatype))
          cloud.is_dense = 0;
  return 0;
}
inline int loadPCDFile(const std::string& file, pcl::PCLPointCloud2& cloud)
{
  Eigen::Vector4f origin;
  Eigen::Quaternionf orientation;
  return loadPCDFile(file, cloud, origin, orientation);
}

// PCDWriter::writeASCII (blob) — io/src/pcd_io.cpp:1099-1215
inline int savePCDFileASCII(const std::string& file, const pcl::PCLPointCloud2& cloud, const Eigen::Vector4f& origin = Eigen::Vector4f::Zero(),
                            const Eigen::Quaternionf& orientation = Eigen::Quaternionf::Identity(), int precision = 8)
{
  if (cloud.data.empty()) std::fprintf(stderr, "[pcl::PCDWriter::writeASCII] Input point cloud has no data!\n");
  if (!detail::blobWritable(cloud, "writeASCII")) return -1;
  std::ofstream out(file, std::ios::binary);
  if (!out) { std::fprintf(stderr, "[pcl::PCDWriter::writeASCII] Could not open file for writing!\n"); return -1; }
  out << detail::blobHeader(cloud, origin, orientation, false, "ascii");
  out.precision(precision);
  const auto fields = detail::fieldsByOffset(cloud);
  const std::size_t npts = static_cast<std::size_t>(cloud.width) * cloud.height;
  for (std::size_t i = 0; i < npts; ++i) {
    bool first = true;
    for (const auto& f : fields)
      for (std::uint32_t c = 0; c < detail::fieldCount(f); ++c) {
        if (!first) out << ' ';
        first = false;
        detail::printAscii(out, cloud.data.data() + i * cloud.point_step + f.offset + c * static_cast<std::uint32_t>(getFieldSize(f.datatype)), f.datatype);
      }
    out << '\n';
  }
  return out ? 0 : -1;
}
// PCDWriter::writeBinary (blob) — io/src/pcd_io.cpp:958-1041, 1241-1330: the records verbatim, gaps declared as "_" fields
inline int savePCDFileBinary(const std::string& file, const pcl::PCLPointCloud2& cloud, const Eigen::Vector4f& origin = Eigen::Vector4f::Zero(),
                             const Eigen::Quaternionf& orientation = Eigen::Quaternionf::Identity())
{
  if (cloud.data.empty()) std::fprintf(stderr, "[pcl::PCDWriter::writeBinary] Input point cloud has no data!\n");
  if (!detail::blobWritable(cloud, "writeBinary")) return -1;
  std::ofstream out(file, std::ios::binary);
  if (!out) { std::fprintf(stderr, "[pcl::PCDWriter::writeBinary] Could not open file for writing!\n"); return -1; }
  out << detail::blobHeader(cloud, origin, orientation, true, "binary");
  out.write(reinterpret_cast<const char*>(cloud.data.data()),
            static_cast<std::streamsize>(static_cast<std::size_t>(cloud.width) * cloud.height * cloud.point_step));
  return out ? 0 : -1;
}
// PCDWriter::writeBinaryCompressed (blob) — io/src/pcd_io.cpp:1045-1098, 1332-1480: one plane per named field, LZF
inline int savePCDFileBinaryCompressed(const std::string& file, const pcl::PCLPointCloud2& cloud,
                                       const Eigen::Vector4f& origin = Eigen::Vector4f::Zero(),
                                       const Eigen::Quaternionf& orientation = Eigen::Quaternionf::Identity())
{
  if (cloud.data.empty()) std::fprintf(stderr, "[pcl::PCDWriter::writeBinaryCompressed] Input point cloud has no data!\n");
  if (!detail::blobWritable(cloud, "writeBinaryCompressed")) return -1;
  const auto fields = detail::fieldsByOffset(cloud);
  const std::size_t npts = static_cast<std::size_t>(cloud.width) * cloud.height;
  std::size_t rec = 0;
  for (const auto& f : fields) rec += detail::fieldCount(f) * static_cast<std::size_t>(getFieldSize(f.datatype));
  const std::size_t data_size = rec * npts;
  if (data_size * 3 / 2 > std::numeric_limits<std::uint32_t>::max()) {
    std::fprintf(stderr, "[pcl::PCDWriter::writeBinaryCompressed] The input data exceeds the maximum size for compressed version 0.7 pcds.\n");
    return -2;
  }
  std::ofstream out(file, std::ios::binary);
  if (!out) { std::fprintf(stderr, "[pcl::PCDWriter::writeBinaryCompressed] Could not open file for writing!\n"); return -1; }
  out << detail::blobHeader(cloud, origin, orientation, false, "binary_compressed");
  std::vector<unsigned char> planes(data_size ? data_size : 1);
  std::size_t plane = 0;
  for (const auto& f : fields) {
    const std::size_t fs = detail::fieldCount(f) * static_cast<std::size_t>(getFieldSize(f.datatype));
    for (std::size_t i = 0; i < npts; ++i) std::memcpy(planes.data() + plane + i * fs, cloud.data.data() + i * cloud.point_step + f.offset, fs);
    plane += fs * npts;
  }
  std::vector<unsigned char> comp(data_size + data_size / 16 + 64);
  std::uint32_t csize = 0;
  const std::uint32_t usize = static_cast<std::uint32_t>(data_size);
  if (data_size) {
    csize = static_cast<std::uint32_t>(detail::lzfCompress(planes.data(), data_size, comp.data(), comp.size()));
    if (csize == 0) { std::fprintf(stderr, "[pcl::PCDWriter::writeBinaryCompressed] Error during compression!\n"); return -1; }
  }
  out.write(reinterpret_cast<const char*>(&csize), 4);
  out.write(reinterpret_cast<const char*>(&usize), 4);
  out.write(reinterpret_cast<const char*>(comp.data()), csize);
  return out ? 0 : -1;
}
// io/include/pcl/io/pcd_io.h:687-697
inline int savePCDFile(const std::string& file, const pcl::PCLPointCloud2& cloud, const Eigen::Vector4f& origin = Eigen::Vector4f::Zero(),
                       const Eigen::Quaternionf& orientation = Eigen::Quaternionf::Identity(), const bool binary_mode = false)
{
  return binary_mode ? savePCDFileBinary(file, cloud, origin, orientation) : savePCDFileASCII(file, cloud, origin, orientation);
}
}  // namespace io

// io/include/pcl/io/pcd_io.h:71-600 — the class forms of the same calls
class PCDReader {
public:
  template <typename PointT> int read(const std::string& file, pcl::PointCloud<PointT>& cloud, const int = 0) { return io::loadPCDFile(file, cloud); }
  int read(const std::string& file, pcl::PCLPointCloud2& cloud, Eigen::Vector4f& origin, Eigen::Quaternionf& orientation, int& pcd_version,
           const int = 0)
  {
    pcd_version = 7;   // PCD_V7: the only version with a VIEWPOINT line (io/include/pcl/io/pcd_io.h:84-110)
    return io::loadPCDFile(file, cloud, origin, orientation);
  }
  int read(const std::string& file, pcl::PCLPointCloud2& cloud, const int = 0) { return io::loadPCDFile(file, cloud); }
};
class PCDWriter {
public:
  template <typename PointT> int write(const std::string& file, const pcl::PointCloud<PointT>& cloud, bool binary = false) { return io::savePCDFile(file, cloud, binary); }
  template <typename PointT> int writeASCII(const std::string& file, const pcl::PointCloud<PointT>& cloud, int precision = 8) { return io::savePCDFileASCII(file, cloud, precision); }
  template <typename PointT> int writeBinary(const std::string& file, const pcl::PointCloud<PointT>& cloud) { return io::savePCDFileBinary(file, cloud); }
  template <typename PointT> int writeBinaryCompressed(const std::string& file, const pcl::PointCloud<PointT>& cloud) { return io::savePCDFileBinaryCompressed(file, cloud); }
  int write(const std::string& file, const pcl::PCLPointCloud2& cloud, const Eigen::Vector4f& origin = Eigen::Vector4f::Zero(),
            const Eigen::Quaternionf& orientation = Eigen::Quaternionf::Identity(), const bool binary = false)
  {
    return io::savePCDFile(file, cloud, origin, orientation, binary);
  }
  int writeASCII(const std::string& file, const pcl::PCLPointCloud2& cloud, const Eigen::Vector4f& origin = Eigen::Vector4f::Zero(),
                 const Eigen::Quaternionf& orientation = Eigen::Quaternionf::Identity(), const int precision = 8)
  {
    return io::savePCDFileASCII(file, cloud, origin, orientation, precision);
  }
  int writeBinary(const std::string& file, const pcl::PCLPointCloud2& cloud, const Eigen::Vector4f& origin = Eigen::Vector4f::Zero(),
                  const Eigen::Quaternionf& orientation = Eigen::Quaternionf::Identity())
  {
    return io::savePCDFileBinary(file, cloud, origin, orientation);
  }
  int writeBinaryCompressed(const std::string& file, const pcl::PCLPointCloud2& cloud, const Eigen::Vector4f& origin = Eigen::Vector4f::Zero(),
                            const Eigen::Quaternionf& orientation = Eigen::Quaternionf::Identity())
  {
    return io::savePCDFileBinaryCompressed(file, cloud, origin, orientation);
  }
};
}  // namespace pcl
