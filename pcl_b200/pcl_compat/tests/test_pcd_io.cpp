// CPU-only checks of pcl/io/pcd_io.h (no device code is reached): LZF coder, the three DATA modes for the point types of
// the path, organised clouds, NaN handling, foreign field types.  Mirrors the round trips of test/io/test_io.cpp
// (PCL, IO / LZF / LZFExtended) on this facade's types.  `dump` mode serves tests/test_pcd_io.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include <pcl/io/pcd_io.h>

static int g_fail = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { ++g_checks; if (!(c)) { ++g_fail; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_EQ(a, b) do { ++g_checks; if (!((a) == (b))) { ++g_fail; std::printf("FAIL %s:%d  %s == %s  (%g vs %g)\n", __FILE__, __LINE__, #a, #b, (double)(a), (double)(b)); } } while (0)

using namespace pcl;

static bool same_bits(float a, float b) { return std::memcmp(&a, &b, 4) == 0 || (std::isnan(a) && std::isnan(b)); }

static int selftest(const std::string& dir)
{
  std::mt19937 rng(7);
  {  // LZF: random, repetitive and tiny buffers survive a round trip; corrupt streams are refused
    for (std::size_t n : {std::size_t(1), std::size_t(2), std::size_t(3), std::size_t(31), std::size_t(33), std::size_t(1000), std::size_t(70000), std::size_t(300000)}) {
      for (int kind = 0; kind < 3; ++kind) {
        std::vector<unsigned char> in(n), comp(n + n / 16 + 64), out(n);
        for (std::size_t i = 0; i < n; ++i)
          in[i] = kind == 0 ? static_cast<unsigned char>(rng()) : kind == 1 ? static_cast<unsigned char>((i / 7) % 5) : static_cast<unsigned char>(i % 251 == 0 ? rng() : 42);
        const std::size_t c = io::detail::lzfCompress(in.data(), n, comp.data(), comp.size());
        EXPECT_TRUE(c > 0);
        if (kind == 1 && n >= 1000) EXPECT_TRUE(c < n * 3 / 4);  // short period: the single-entry hash table finds short matches
        if (kind == 2 && n >= 1000) EXPECT_TRUE(c < n / 4);
        const std::size_t d = io::detail::lzfDecompress(comp.data(), c, out.data(), n);
        EXPECT_EQ(d, n);
        EXPECT_TRUE(std::memcmp(in.data(), out.data(), n) == 0);
        if (n > 40) {
          EXPECT_EQ(io::detail::lzfDecompress(comp.data(), c, out.data(), n - 1), 0u);  // output too small
          EXPECT_EQ(io::detail::lzfCompress(in.data(), n, comp.data(), 8), 0u);          // no room
        }
      }
    }
    const unsigned char bad[] = {0xe0, 0x05, 0x10};  // back reference before the start of the output
    unsigned char o[64];
    EXPECT_EQ(io::detail::lzfDecompress(bad, 3, o, 64), 0u);
    // a hand-built stream: literal "abc", then a match of length 6 at distance 3 -> "abcabcabc"
    const unsigned char hand[] = {0x02, 'a', 'b', 'c', static_cast<unsigned char>((4u << 5) | 0), 0x02};
    EXPECT_EQ(io::detail::lzfDecompress(hand, 6, o, 64), 9u);
    EXPECT_TRUE(std::memcmp(o, "abcabcabc", 9) == 0);
  }
  std::uniform_real_distribution<float> U(-5.f, 5.f);
  {  // PointXYZ, organised, with NaN points: all three modes
    PointCloud<PointXYZ> cloud;
    for (int i = 0; i < 640 * 48; ++i) cloud.points.emplace_back(U(rng), U(rng), U(rng));
    cloud.width = 640;
    cloud.height = 48;
    cloud.points[5].x = std::numeric_limits<float>::quiet_NaN();
    cloud.is_dense = false;
    cloud.sensor_origin_[0] = 1.f; cloud.sensor_origin_[1] = 2.f; cloud.sensor_origin_[2] = 3.f;
    const char* names[3] = {"a.pcd", "b.pcd", "c.pcd"};
    for (int mode = 0; mode < 3; ++mode) {
      const std::string f = dir + "/" + names[mode];
      const int rc = mode == 0 ? io::savePCDFileASCII(f, cloud, 9) : mode == 1 ? io::savePCDFileBinary(f, cloud) : io::savePCDFileBinaryCompressed(f, cloud);
      EXPECT_EQ(rc, 0);
      PointCloud<PointXYZ> back;
      EXPECT_EQ(io::loadPCDFile(f, back), 0);
      EXPECT_EQ(back.size(), cloud.size());
      EXPECT_EQ(back.width, 640u);
      EXPECT_EQ(back.height, 48u);
      EXPECT_TRUE(!back.is_dense);
      EXPECT_EQ(back.sensor_origin_[0], 1.f);
      EXPECT_EQ(back.sensor_origin_[2], 3.f);
      bool ok = back.size() == cloud.size();
      for (std::size_t i = 0; ok && i < cloud.size(); ++i)
        ok = same_bits(back[i].x, cloud[i].x) && same_bits(back[i].y, cloud[i].y) && same_bits(back[i].z, cloud[i].z) && back[i].data[3] == 1.f;
      EXPECT_TRUE(ok);  // 9 significant digits round-trip a float exactly in ASCII too
    }
  }
  {  // PointNormal and Normal records; a PointXYZ reader of a PointNormal file takes the fields it knows
    PointCloud<PointNormal> cloud;
    for (int i = 0; i < 1000; ++i) cloud.push_back(PointNormal(U(rng), U(rng), U(rng), U(rng), U(rng), U(rng), U(rng)));
    for (int mode = 0; mode < 3; ++mode) {
      const std::string f = dir + "/pn.pcd";
      EXPECT_EQ(mode == 0 ? io::savePCDFile(f, cloud) : mode == 1 ? io::savePCDFile(f, cloud, true) : io::savePCDFileBinaryCompressed(f, cloud), 0);
      PointCloud<PointNormal> back;
      EXPECT_EQ(io::loadPCDFile(f, back), 0);
      EXPECT_TRUE(back.is_dense);
      EXPECT_EQ(back.size(), 1000u);
      bool ok = back.size() == 1000u;
      const float tol = mode == 0 ? 1e-6f : 0.f;
      for (std::size_t i = 0; ok && i < 1000; ++i)
        ok = std::fabs(back[i].x - cloud[i].x) <= tol * 5 && std::fabs(back[i].normal_y - cloud[i].normal_y) <= tol * 5 &&
             std::fabs(back[i].curvature - cloud[i].curvature) <= tol * 5 && std::fabs(back[i].z - cloud[i].z) <= tol * 5;
      EXPECT_TRUE(ok);
      PointCloud<PointXYZ> xyz;
      EXPECT_EQ(io::loadPCDFile(f, xyz), 0);
      EXPECT_EQ(xyz.size(), 1000u);
      EXPECT_TRUE(std::fabs(xyz[999].y - cloud[999].y) <= tol * 5);
      PointCloud<Normal> nrm;
      EXPECT_EQ(io::loadPCDFile(f, nrm), 0);
      EXPECT_TRUE(std::fabs(nrm[17].normal_z - cloud[17].normal_z) <= tol * 5 && std::fabs(nrm[17].curvature - cloud[17].curvature) <= tol * 5);
    }
    PCDWriter w;
    PCDReader r;
    EXPECT_EQ(w.writeBinaryCompressed(dir + "/w.pcd", cloud), 0);
    PointCloud<PointNormal> back;
    EXPECT_EQ(r.read(dir + "/w.pcd", back), 0);
    EXPECT_EQ(back.size(), 1000u);
  }
  {  // a hand-written binary file with foreign field types, a padding field and COUNT > 1
    const std::string f = dir + "/foreign.pcd";
    std::ofstream out(f, std::ios::binary);
    out << "# test\nVERSION .7\nFIELDS x y z _ rgb intensity label hist\nSIZE 8 4 4 1 4 2 1 4\nTYPE F F F U U U I F\nCOUNT 1 1 1 3 1 1 1 2\n"
           "WIDTH 2\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 2\nDATA binary\n";
    for (int i = 0; i < 2; ++i) {
      const double x = 1.5 + i;
      const float y = 2.5f + i, z = -3.f - i, hist[2] = {0.1f, 0.2f};
      const unsigned char pad[3] = {0, 0, 0};
      const std::uint32_t rgb = 0x00ff8040u;
      const std::uint16_t inten = 777;
      const std::int8_t label = -3;
      out.write(reinterpret_cast<const char*>(&x), 8);
      out.write(reinterpret_cast<const char*>(&y), 4);
      out.write(reinterpret_cast<const char*>(&z), 4);
      out.write(reinterpret_cast<const char*>(pad), 3);
      out.write(reinterpret_cast<const char*>(&rgb), 4);
      out.write(reinterpret_cast<const char*>(&inten), 2);
      out.write(reinterpret_cast<const char*>(&label), 1);
      out.write(reinterpret_cast<const char*>(hist), 8);
    }
    out.close();
    PointCloud<PointXYZ> c;
    EXPECT_EQ(io::loadPCDFile(f, c), 0);
    EXPECT_EQ(c.size(), 2u);
    EXPECT_EQ(c[1].x, 2.5f);
    EXPECT_EQ(c[1].y, 3.5f);
    EXPECT_EQ(c[0].z, -3.f);
    EXPECT_TRUE(c.is_dense);
  }
  {  // old-style ASCII header (v.5: COLUMNS, no SIZE/TYPE/COUNT/WIDTH/HEIGHT) like test/bun4.pcd; truncated files
    const std::string f = dir + "/old.pcd";
    std::ofstream out(f);
    out << "# .PCD v.5 - Point Cloud Data file format\nCOLUMNS x y z\nPOINTS 3\nDATA ascii\n1 2 3\n4 5 nan\n7 8 9\n";
    out.close();
    PointCloud<PointXYZ> c;
    EXPECT_EQ(io::loadPCDFile(f, c), 0);
    EXPECT_EQ(c.size(), 3u);
    EXPECT_EQ(c.width, 3u);
    EXPECT_EQ(c.height, 1u);
    EXPECT_TRUE(!c.is_dense && std::isnan(c[1].z) && c[2].y == 8.f);
    std::ofstream t(dir + "/trunc.pcd");
    t << "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 5\nHEIGHT 1\nPOINTS 5\nDATA binary\nabc";
    t.close();
    EXPECT_EQ(io::loadPCDFile(dir + "/trunc.pcd", c), -1);
    EXPECT_EQ(io::loadPCDFile(dir + "/does_not_exist.pcd", c), -1);
  }
  {  // TEST (PCL, PCDReaderWriter) — test/io/test_io.cpp:879-946 with PointNormal in the place of PointXYZI: a 640 x 480 organised
     // cloud -> blob (whole records, padding included) -> binary file -> blob -> cloud
    PCLPointCloud2 cloud_blob;
    PointCloud<PointNormal> cloud;
    cloud.width = 640;
    cloud.height = 480;
    cloud.points.resize(static_cast<std::size_t>(cloud.width) * cloud.height);
    cloud.is_dense = true;
    const std::size_t nr_p = cloud.size();
    for (std::size_t i = 0; i < nr_p; ++i) {
      cloud[i].x = static_cast<float>(1024 * (rng() % 32768) / 32768.0);
      cloud[i].y = static_cast<float>(1024 * (rng() % 32768) / 32768.0);
      cloud[i].z = static_cast<float>(1024 * (rng() % 32768) / 32768.0);
      cloud[i].normal_x = 0.f; cloud[i].normal_y = 0.6f; cloud[i].normal_z = 0.8f;
      cloud[i].curvature = static_cast<float>(i);
    }
    const PointNormal first = cloud[0], last = cloud[nr_p - 1];
    toPCLPointCloud2(cloud, cloud_blob);
    EXPECT_EQ(cloud_blob.width, cloud.width);
    EXPECT_EQ(cloud_blob.height, cloud.height);
    EXPECT_EQ(bool(cloud_blob.is_dense), cloud.is_dense);
    EXPECT_EQ(cloud_blob.data.size(), static_cast<std::size_t>(cloud_blob.width) * cloud_blob.height * sizeof(PointNormal));
    const std::string f = dir + "/test_pcl_io.pcd";
    PCDWriter writer;
    EXPECT_EQ(writer.write(f, cloud_blob, Eigen::Vector4f::Zero(), Eigen::Quaternionf::Identity(), true), 0);
    PCDReader reader;
    EXPECT_EQ(reader.read(f, cloud_blob), 0);
    EXPECT_EQ(cloud_blob.width, cloud.width);
    EXPECT_EQ(cloud_blob.height, cloud.height);
    EXPECT_EQ(bool(cloud_blob.is_dense), cloud.is_dense);
    EXPECT_EQ(cloud_blob.data.size(), static_cast<std::size_t>(cloud_blob.width) * cloud_blob.height * sizeof(PointNormal));
    fromPCLPointCloud2(cloud_blob, cloud);
    EXPECT_EQ(cloud.width, cloud_blob.width);
    EXPECT_EQ(cloud.height, cloud_blob.height);
    EXPECT_EQ(cloud.is_dense, bool(cloud_blob.is_dense));
    EXPECT_EQ(cloud.size(), nr_p);
    EXPECT_TRUE(cloud[0].x == first.x && cloud[0].y == first.y && cloud[0].z == first.z && cloud[0].curvature == first.curvature && cloud[0].normal_z == first.normal_z);
    EXPECT_TRUE(cloud[nr_p - 1].x == last.x && cloud[nr_p - 1].y == last.y && cloud[nr_p - 1].z == last.z && cloud[nr_p - 1].curvature == last.curvature);
    std::remove(f.c_str());
  }
  {  // TEST (PCL, EmptyCloudToPCD) — test/io/test_io.cpp:692-880: empty clouds through every writer and back, headers without
     // WIDTH / HEIGHT, a non-numeric HEIGHT, blobs without fields
    PointCloud<PointXYZ> cloud;
    const std::string f = dir + "/empty.pcd";
    for (int mode = 0; mode < 3; ++mode) {
      const int res = mode == 0 ? io::savePCDFileASCII(f, cloud) : mode == 1 ? io::savePCDFileBinary(f, cloud) : io::savePCDFileBinaryCompressed(f, cloud);
      EXPECT_EQ(res, 0);
      PointCloud<PointXYZ> in;
      in.width = 10;    // loadPCDFile must overwrite these
      in.height = 10;
      EXPECT_EQ(io::loadPCDFile(f, in), 0);
      EXPECT_EQ(in.width, cloud.width);
      EXPECT_EQ(in.height, cloud.height);
      EXPECT_EQ(in.size(), 0u);
      std::remove(f.c_str());
    }
    PCLPointCloud2 cloud2;
    for (const char* name : {"x", "y", "z"}) {
      PCLPointField fld;
      fld.name = name;
      fld.datatype = PCLPointField::FLOAT32;
      cloud2.fields.push_back(fld);
    }
    cloud2.is_dense = true;
    for (int mode = 0; mode < 3; ++mode) {
      const int res = mode == 0 ? io::savePCDFile(f, cloud2, Eigen::Vector4f::Zero(), Eigen::Quaternionf::Identity())
                      : mode == 1 ? io::savePCDFile(f, cloud2, Eigen::Vector4f::Zero(), Eigen::Quaternionf::Identity(), true)
                                  : PCDWriter().writeBinaryCompressed(f, cloud2);
      EXPECT_EQ(res, 0);
      PCLPointCloud2 in2;
      in2.width = 10;
      in2.height = 10;
      EXPECT_EQ(io::loadPCDFile(f, in2), 0);
      EXPECT_EQ(in2.width, cloud2.width);
      EXPECT_EQ(in2.height, cloud2.height);
      std::remove(f.c_str());
    }
    auto write_text = [&](const char* text) { std::ofstream fs(f); fs << text; };
    {  // WIDTH and HEIGHT not defined
      write_text("# .PCD v0.5 - Point Cloud Data file format\nVERSION 0.5\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nPOINTS 2\nDATA ascii\n1 2 3 4\n5 6 7 8");
      PCLPointCloud2 in2;
      EXPECT_EQ(io::loadPCDFile(f, in2), 0);
      EXPECT_EQ(in2.width, 2u);
      EXPECT_EQ(in2.height, 1u);
      EXPECT_TRUE(in2.is_dense);
      EXPECT_EQ(in2.data.size(), std::size_t(2 * 4 * 4));
    }
    {  // HEIGHT not defined
      write_text("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 2\nPOINTS 2\nDATA ascii\n1 2 3 4\n5 6 7 8");
      PCLPointCloud2 in2;
      EXPECT_EQ(io::loadPCDFile(f, in2), 0);
      EXPECT_EQ(in2.width, 2u);
      EXPECT_EQ(in2.height, 1u);
      EXPECT_TRUE(in2.is_dense);
      EXPECT_EQ(in2.data.size(), std::size_t(2 * 4 * 4));
    }
    {  // invalid height
      write_text("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 2\nHEIGHT a\nPOINTS 2\nDATA ascii\n1 2 3 4\n5 6 7 8");
      PCLPointCloud2 in2;
      EXPECT_EQ(io::loadPCDFile(f, in2), -1);
    }
    std::remove(f.c_str());
    {  // no field data: every writer refuses
      PCLPointCloud2 empty_cloud;
      EXPECT_EQ(io::savePCDFile(f, empty_cloud, Eigen::Vector4f::Zero(), Eigen::Quaternionf::Identity()), -1);
      EXPECT_EQ(io::savePCDFile(f, empty_cloud, Eigen::Vector4f::Zero(), Eigen::Quaternionf::Identity(), true), -1);
      EXPECT_EQ(PCDWriter().writeBinaryCompressed(f, empty_cloud), -1);
      std::remove(f.c_str());
    }
  }
  {  // the type-erased route: a blob with mixed field types, a COUNT > 1 field and padding survives all three encodings;
     // typed and blob readers agree; concatenateFields / getFieldsList / getFieldIndex (common/src/io.cpp)
    PCLPointCloud2 b;
    auto add = [&](const char* name, std::uint32_t off, std::uint8_t dt, std::uint32_t cnt) {
      PCLPointField f;
      f.name = name; f.offset = off; f.datatype = dt; f.count = cnt;
      b.fields.push_back(f);
    };
    add("x", 0, PCLPointField::FLOAT32, 1);
    add("y", 4, PCLPointField::FLOAT32, 1);
    add("z", 8, PCLPointField::FLOAT32, 1);
    add("ring", 16, PCLPointField::UINT16, 1);          // bytes 12..15 and 18..19 are padding
    add("label", 20, PCLPointField::INT32, 1);
    add("hist", 24, PCLPointField::UINT8, 3);
    add("t", 32, PCLPointField::FLOAT64, 1);            // bytes 27..31 padding
    add("tag", 40, PCLPointField::INT8, 1);
    b.point_step = 48;
    b.width = 7;
    b.height = 3;
    b.row_step = b.point_step * b.width;
    b.is_dense = 1;
    const std::size_t n = 21;
    b.data.assign(n * b.point_step, 0xAB);              // padding bytes are arbitrary
    for (std::size_t i = 0; i < n; ++i) {
      unsigned char* r = b.data.data() + i * b.point_step;
      const float x = 0.5f * i - 3.f, y = 1.f / (i + 1.f), z = (i == 4) ? std::numeric_limits<float>::quiet_NaN() : -7.25f * i;
      const std::uint16_t ring = static_cast<std::uint16_t>(60000 + i);
      const std::int32_t label = -100000 * static_cast<std::int32_t>(i) + 7;
      const std::uint8_t hist[3] = {static_cast<std::uint8_t>(i), static_cast<std::uint8_t>(255 - i), 128};
      const double t = 1e-3 * i + 1234567.125;
      const std::int8_t tag = static_cast<std::int8_t>(i - 10);
      std::memcpy(r, &x, 4); std::memcpy(r + 4, &y, 4); std::memcpy(r + 8, &z, 4); std::memcpy(r + 16, &ring, 2);
      std::memcpy(r + 20, &label, 4); std::memcpy(r + 24, hist, 3); std::memcpy(r + 32, &t, 8); std::memcpy(r + 40, &tag, 1);
    }
    EXPECT_TRUE(getFieldsList(b) == "x y z ring label hist t tag");
    EXPECT_EQ(getFieldIndex(b, "label"), 4);
    EXPECT_EQ(getFieldIndex(b, "nope"), -1);
    Eigen::Vector4f org;
    org[0] = 1.5f; org[1] = -2.f; org[2] = 0.25f;
    const Eigen::Quaternionf ori(0.5f, 0.5f, -0.5f, 0.5f);
    auto same_fields = [&](const PCLPointCloud2& a, bool offsets_too) {
      bool ok = a.fields.size() == b.fields.size();
      for (std::size_t f = 0; ok && f < a.fields.size(); ++f)
        ok = a.fields[f].name == b.fields[f].name && a.fields[f].datatype == b.fields[f].datatype && a.fields[f].count == b.fields[f].count &&
             (!offsets_too || a.fields[f].offset == b.fields[f].offset);
      return ok;
    };
    auto same_values = [&](const PCLPointCloud2& a) {
      bool ok = a.width == b.width && a.height == b.height && a.data.size() == std::size_t(a.point_step) * n;
      for (std::size_t i = 0; ok && i < n; ++i)
        for (std::size_t f = 0; ok && f < b.fields.size(); ++f) {
          const std::size_t bytes = b.fields[f].count * getFieldSize(b.fields[f].datatype);
          const unsigned char *pa = a.data.data() + i * a.point_step + a.fields[f].offset, *pb = b.data.data() + i * b.point_step + b.fields[f].offset;
          if (b.fields[f].name == "z" && i == 4) { float v; std::memcpy(&v, pa, 4); ok = std::isnan(v); }
          else ok = std::memcmp(pa, pb, bytes) == 0;
        }
      return ok;
    };
    for (int mode = 0; mode < 3; ++mode) {
      const std::string f = dir + "/blob" + std::to_string(mode) + ".pcd";
      const int rc = mode == 0 ? io::savePCDFileASCII(f, b, org, ori, 17) : mode == 1 ? io::savePCDFileBinary(f, b, org, ori) : io::savePCDFileBinaryCompressed(f, b, org, ori);
      EXPECT_EQ(rc, 0);
      PCLPointCloud2 r;
      Eigen::Vector4f o2;
      Eigen::Quaternionf q2;
      EXPECT_EQ(io::loadPCDFile(f, r, o2, q2), 0);
      EXPECT_TRUE(same_fields(r, mode == 1));          // only the binary form keeps the padded layout
      EXPECT_EQ(r.point_step, mode == 1 ? 48u : 4u + 4 + 4 + 2 + 4 + 3 + 8 + 1);
      EXPECT_TRUE(same_values(r));
      EXPECT_EQ(r.is_dense, 0);                        // the NaN
      EXPECT_TRUE(o2[0] == 1.5f && o2[1] == -2.f && o2[2] == 0.25f && q2 == ori);
      PointCloud<PointXYZ> typed;                      // the typed reader sees the same coordinates
      EXPECT_EQ(io::loadPCDFile(f, typed), 0);
      EXPECT_TRUE(typed.size() == n && typed.width == 7 && typed.height == 3 && typed[3].x == -1.5f && std::isnan(typed[4].z));
      PointCloud<PointXYZ> conv;
      fromPCLPointCloud2(r, conv);
      EXPECT_TRUE(conv.size() == n && same_bits(conv[20].y, typed[20].y) && same_bits(conv[7].z, typed[7].z));
    }
    // PCDReader / PCDWriter class forms
    PCDWriter w;
    PCDReader rd;
    PCLPointCloud2 r2;
    int ver = 0;
    Eigen::Vector4f o3;
    Eigen::Quaternionf q3;
    EXPECT_EQ(w.writeBinaryCompressed(dir + "/blobw.pcd", b, org, ori), 0);
    EXPECT_EQ(rd.read(dir + "/blobw.pcd", r2, o3, q3, ver), 0);
    EXPECT_TRUE(ver == 7 && same_values(r2));
    EXPECT_EQ(w.write(dir + "/blobw2.pcd", b), 0);     // ascii, default pose
    EXPECT_EQ(rd.read(dir + "/blobw2.pcd", r2), 0);
    EXPECT_TRUE(r2.fields.size() == 8 && r2.width == 7);
    // concatenateFields: the aligned coordinates (cloud2) take the place of x y z, the rest of cloud1 follows
    PointCloud<PointXYZ> moved;
    moved.resize(7, 3, PointXYZ(9.f, 8.f, 7.f));
    PCLPointCloud2 mb, cat;
    toPCLPointCloud2(moved, mb);
    EXPECT_TRUE(concatenateFields(b, mb, cat));
    EXPECT_TRUE(getFieldsList(cat) == "x y z ring label hist t tag");
    EXPECT_EQ(cat.point_step, 16u + (4 + 4 + 8 + 8 + 8));   // PointXYZ record + each carried field with its room in cloud1
    EXPECT_TRUE(cat.width == 7 && cat.height == 3 && cat.data.size() == n * cat.point_step);
    {
      float x;
      std::uint16_t ring;
      double t;
      const unsigned char* r = cat.data.data() + 5 * cat.point_step;
      std::memcpy(&x, r + cat.fields[0].offset, 4);
      std::memcpy(&ring, r + cat.fields[getFieldIndex(cat, "ring")].offset, 2);
      std::memcpy(&t, r + cat.fields[getFieldIndex(cat, "t")].offset, 8);
      EXPECT_TRUE(x == 9.f && ring == 60005 && t == 1e-3 * 5 + 1234567.125);
    }
    PCLPointCloud2 other = mb;
    other.width = 3; other.height = 1;
    EXPECT_TRUE(!concatenateFields(b, other, cat));
    // a field type the blob cannot describe is refused
    std::ofstream odd(dir + "/i8.pcd");
    odd << "FIELDS x big\nSIZE 4 8\nTYPE F I\nCOUNT 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2\n";
    odd.close();
    EXPECT_EQ(io::loadPCDFile(dir + "/i8.pcd", r2), -1);
  }
  std::printf("%s: %d checks, %d failures\n", g_fail ? "FAILED" : "PASSED", g_checks, g_fail);
  return g_fail ? 1 : 0;
}

int main(int argc, char** argv)
{
  if (argc >= 4 && std::string(argv[1]) == "dumpblob") {  // <in.pcd> <out.bin>: like dump, through the blob reader + fromPCLPointCloud2
    PCLPointCloud2 b;
    if (io::loadPCDFile(argv[2], b) != 0) return 2;
    PointCloud<PointXYZ> c;
    fromPCLPointCloud2(b, c);
    std::ofstream out(argv[3], std::ios::binary);
    const std::uint64_t n = c.size();
    const std::uint32_t w = c.width, h = c.height, d = b.is_dense ? 1 : 0;
    out.write(reinterpret_cast<const char*>(&n), 8);
    out.write(reinterpret_cast<const char*>(&w), 4);
    out.write(reinterpret_cast<const char*>(&h), 4);
    out.write(reinterpret_cast<const char*>(&d), 4);
    for (const auto& p : c.points) out.write(reinterpret_cast<const char*>(&p.x), 12);
    return 0;
  }
  if (argc >= 3 && std::string(argv[1]) == "selftest") return selftest(argv[2]);
  if (argc >= 4 && std::string(argv[1]) == "dump") {  // <in.pcd> <out.bin>: u64 n, u32 w, u32 h, u32 dense, n * xyz floats
    PointCloud<PointXYZ> c;
    if (io::loadPCDFile(argv[2], c) != 0) return 2;
    std::ofstream out(argv[3], std::ios::binary);
    const std::uint64_t n = c.size();
    const std::uint32_t w = c.width, h = c.height, d = c.is_dense ? 1 : 0;
    out.write(reinterpret_cast<const char*>(&n), 8);
    out.write(reinterpret_cast<const char*>(&w), 4);
    out.write(reinterpret_cast<const char*>(&h), 4);
    out.write(reinterpret_cast<const char*>(&d), 4);
    for (const auto& p : c.points) out.write(reinterpret_cast<const char*>(&p.x), 12);
    return 0;
  }
  if (argc >= 5 && std::string(argv[1]) == "recode") {  // <in.pcd> <out.pcd> <ascii|binary|compressed>
    PointCloud<PointXYZ> c;
    if (io::loadPCDFile(argv[2], c) != 0) return 2;
    const std::string m = argv[4];
    return m == "ascii" ? io::savePCDFileASCII(argv[3], c, 9) : m == "binary" ? io::savePCDFileBinary(argv[3], c) : io::savePCDFileBinaryCompressed(argv[3], c);
  }
  std::fprintf(stderr, "usage: test_pcd_io selftest <dir> | dump <in> <out> | recode <in> <out> <mode>\n");
  return 64;
}
