// CPU-only checks of pcl/io/pcd_io.h (no device code is reached): LZF coder, the three DATA modes for the point types of
// the path, organised clouds, NaN handling, foreign field types.  Mirrors the round trips of test/io/test_io.cpp
// (PCL, IO / LZF / LZFExtended) on this facade's types.  `dump` mode serves tests/test_pcd_io.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
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
  std::printf("%s: %d checks, %d failures\n", g_fail ? "FAILED" : "PASSED", g_checks, g_fail);
  return g_fail ? 1 : 0;
}

int main(int argc, char** argv)
{
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
