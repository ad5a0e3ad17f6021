// Host-only behaviour of the facade (no device library needed): the pcl::PointCloud container's width / height
// bookkeeping — the scenarios of the reference's test/common/test_pointcloud.cpp:24-395 — and the small host classes
// added beside it.  Exit code 0 and "PASSED" on success.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <vector>

#include <string>

#include <pcl/common/centroid.h>
#include <pcl/common/common.h>
#include <pcl/common/io.h>
#include <pcl/common/transforms.h>
#include <pcl/filters/filter.h>
#include <pcl/correspondence.h>
#include <pcl/features/normal_3d.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/io/pcd_io.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/search/kdtree.h>
#include <pcl/search/search.h>

using namespace pcl;

static int g_fail = 0, g_checks = 0;
#define CHECK(c) do { ++g_checks; if (!(c)) { ++g_fail; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)

// a brute-force searcher that implements only the two pure virtuals of pcl::search::Search: every other form must work
// through them (search.h:144-146, 271-273)
struct BruteForce : search::Search<PointXYZ> {
  BruteForce() : search::Search<PointXYZ>("BruteForce", true) {}
  std::vector<std::pair<float, index_t>> all(const PointXYZ& q) const
  {
    std::vector<std::pair<float, index_t>> v;
    const auto in = getInputCloud();
    const auto idx = getIndices();
    const std::size_t n = idx ? idx->size() : in->size();
    for (std::size_t j = 0; j < n; ++j) {
      const index_t i = idx ? (*idx)[j] : static_cast<index_t>(j);
      const PointXYZ& p = (*in)[i];
      v.emplace_back((p.x - q.x) * (p.x - q.x) + (p.y - q.y) * (p.y - q.y) + (p.z - q.z) * (p.z - q.z), i);
    }
    std::sort(v.begin(), v.end());
    return v;
  }
  int nearestKSearch(const PointXYZ& q, int k, Indices& ki, std::vector<float>& kd) const override
  {
    auto v = all(q);
    if ((int)v.size() > k) v.resize(k);
    ki.clear(); kd.clear();
    for (auto& e : v) { ki.push_back(e.second); kd.push_back(e.first); }
    return (int)v.size();
  }
  int radiusSearch(const PointXYZ& q, double r, Indices& ki, std::vector<float>& kd, unsigned int max_nn = 0) const override
  {
    ki.clear(); kd.clear();
    for (auto& e : all(q))
      if (e.first < r * r && (max_nn == 0 || ki.size() < max_nn)) { ki.push_back(e.second); kd.push_back(e.first); }
    return (int)ki.size();
  }
  using search::Search<PointXYZ>::nearestKSearch;
  using search::Search<PointXYZ>::radiusSearch;
  using search::Search<PointXYZ>::sortResults;
};

static PointCloud<PointXYZ> grid() { PointCloud<PointXYZ> c; c.resize(640, 480, PointXYZ(1, 1, 1)); return c; }

// "normals <cloud.pcd> <out.bin>": plane parameters and curvature of deterministic index subsets of the cloud, written
// as raw floats (5 per subset: nx ny nz d curvature) for tests/test_facade_host.py to compare with the oracle bit for bit;
// subset s holds the indices (s * 37 + j * (s % 5 + 1)) % n, j < 3 + s % 40
static int dump_normals(const char* pcd, const char* out_path)
{
  PointCloud<PointXYZ> cloud;
  if (io::loadPCDFile(pcd, cloud)) return 2;
  std::vector<float> out;
  const int n = static_cast<int>(cloud.size());
  for (int s = 0; s < 200; ++s) {
    Indices idx;
    for (int j = 0; j < 3 + s % 40; ++j) idx.push_back((s * 37 + j * (s % 5 + 1)) % n);
    Eigen::Vector4f plane;
    float curvature = 0.f;
    computePointNormal(cloud, idx, plane, curvature);
    for (int d = 0; d < 4; ++d) out.push_back(plane[d]);
    out.push_back(curvature);
  }
  {  // the whole cloud, both overloads, then flipped towards the origin as in test_normal_estimation.cpp:103-138
    Indices all(cloud.size());
    for (int i = 0; i < n; ++i) all[i] = i;
    Eigen::Vector4f plane, plane2;
    float c1 = 0.f, c2 = 0.f;
    computePointNormal(cloud, all, plane, c1);
    computePointNormal(cloud, plane2, c2);
    for (int d = 0; d < 4; ++d) out.push_back(plane[d]);
    out.push_back(c1);
    for (int d = 0; d < 4; ++d) out.push_back(plane2[d]);
    out.push_back(c2);
    flipNormalTowardsViewpoint(cloud[0], 0, 0, 0, plane);
    for (int d = 0; d < 4; ++d) out.push_back(plane[d]);
    float nx = plane2[0], ny = plane2[1], nz = plane2[2];
    flipNormalTowardsViewpoint(cloud[0], 0, 0, 0, nx, ny, nz);
    out.push_back(nx);
    NormalEstimation<PointXYZ, Normal> ne;
    float mx, my, mz, mc;
    ne.computePointNormal(cloud, all, mx, my, mz, mc);
    out.push_back(mx); out.push_back(my); out.push_back(mz); out.push_back(mc);
  }
  FILE* f = std::fopen(out_path, "wb");
  if (!f) return 2;
  std::fwrite(out.data(), sizeof(float), out.size(), f);
  std::fclose(f);
  return 0;
}

// The bodies of the reference's test/common/test_centroid.cpp for the functions on the path: TEST (PCL, compute3DCentroidFloat)
// :53-161, compute3DCentroidDouble :164-271, computeMeanAndCovariance :700-840, demeanPointCloud :1210-1236 (cloud = bun0.pcd).
template <typename Scalar>
static void centroid_body()
{
  using Vec4 = Eigen::Matrix<Scalar, 4, 1>;
  Indices indices;
  PointXYZ point;
  PointCloud<PointXYZ> cloud;
  Vec4 centroid;
  centroid[0] = Scalar(0.125); centroid[1] = Scalar(-0.5); centroid[2] = Scalar(0.75); centroid[3] = Scalar(-0.25);   // "Random()"
  const Vec4 old_centroid = centroid;
  cloud.is_dense = true;    // empty and dense
  CHECK(compute3DCentroid(cloud, centroid) == 0 && old_centroid == centroid);
  cloud.is_dense = false;   // empty, not dense
  CHECK(compute3DCentroid(cloud, centroid) == 0 && old_centroid == centroid);
  point.x = point.y = point.z = std::numeric_limits<float>::quiet_NaN();   // only invalid points
  cloud.push_back(point);
  CHECK(compute3DCentroid(cloud, centroid) == 0 && old_centroid == centroid);
  cloud.push_back(point);
  indices.push_back(1);
  CHECK(compute3DCentroid(cloud, indices, centroid) == 0 && old_centroid == centroid);
  cloud.clear();
  indices.clear();
  for (point.x = -1; point.x < 2; point.x += 2)
    for (point.y = -1; point.y < 2; point.y += 2)
      for (point.z = -1; point.z < 2; point.z += 2) cloud.push_back(point);
  cloud.is_dense = true;
  auto reset = [&] { centroid[0] = -100; centroid[1] = -200; centroid[2] = -300; };
  reset();
  CHECK(compute3DCentroid(cloud, centroid) == 8 && centroid[0] == 0 && centroid[1] == 0 && centroid[2] == 0 && centroid[3] == 1);
  reset();
  indices = {2, 3, 6, 7};   // only positive y values
  CHECK(compute3DCentroid(cloud, indices, centroid) == 4 && centroid[0] == 0 && centroid[1] == 1 && centroid[2] == 0 && centroid[3] == 1);
  point.x = point.y = point.z = std::numeric_limits<float>::quiet_NaN();
  cloud.push_back(point);
  cloud.is_dense = false;
  reset();
  CHECK(compute3DCentroid(cloud, centroid) == 8 && centroid[0] == 0 && centroid[1] == 0 && centroid[2] == 0 && centroid[3] == 1);
  reset();
  indices.push_back(8);   // the NaN
  CHECK(compute3DCentroid(cloud, indices, centroid) == 4 && centroid[0] == 0 && centroid[1] == 1 && centroid[2] == 0 && centroid[3] == 1);
}

// TYPED_TEST (Transforms, PointCloudXYZDense / DenseIndexed / Sparse / XYZRGBNormalDense / DenseIndexed) —
// test/common/test_transforms.cpp:58-209 for the Matrix<float, 4, 4> and Matrix<double, 4, 4> instances (PointNormal stands in for
// PointXYZRGBNormal; the expected values are the rigid transform applied in double and narrowed, the bound 10 epsilon of Scalar)
template <typename Scalar>
static void transforms_body(unsigned seed)
{
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  const double rx = U(rng), ry = U(rng), rz = U(rng), tx = U(rng), ty = U(rng), tz = U(rng);
  // getTransformation (x, y, z, roll, pitch, yaw): R = Rz(yaw) Ry(pitch) Rx(roll)
  const double A = std::cos(rz), B = std::sin(rz), C = std::cos(ry), D = std::sin(ry), E = std::cos(rx), F = std::sin(rx), DE = D * E, DF = D * F;
  const double R[9] = {A * C, A * DF - B * E, B * F + A * DE, B * C, A * E + B * DF, B * DE - A * F, -D, C * F, C * E};
  Eigen::Matrix<Scalar, 4, 4> tf = Eigen::Matrix<Scalar, 4, 4>::Identity();
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) tf(r, c) = static_cast<Scalar>(R[3 * r + c]);
    tf(r, 3) = static_cast<Scalar>(r == 0 ? tx : r == 1 ? ty : tz);
  }
  const std::size_t CLOUD_SIZE = 100;
  const double ABS_ERROR = static_cast<double>(std::numeric_limits<Scalar>::epsilon()) * 10;
  PointCloud<PointXYZ> p_xyz, p_xyz_trans;
  PointCloud<PointNormal> p_n, p_n_trans;
  for (std::size_t i = 0; i < CLOUD_SIZE; ++i) {
    const float v[3] = {static_cast<float>(U(rng)), static_cast<float>(U(rng)), static_cast<float>(U(rng))};
    double nn[3] = {U(rng), U(rng), U(rng)};
    const double nl = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
    const float n[3] = {static_cast<float>(nn[0] / nl), static_cast<float>(nn[1] / nl), static_cast<float>(nn[2] / nl)};
    PointNormal a, b;
    a.x = v[0]; a.y = v[1]; a.z = v[2]; a.normal_x = n[0]; a.normal_y = n[1]; a.normal_z = n[2]; a.curvature = static_cast<float>(U(rng));
    b = a;
    float* bo[3] = {&b.x, &b.y, &b.z};
    float* bn[3] = {&b.normal_x, &b.normal_y, &b.normal_z};
    for (int r = 0; r < 3; ++r) {   // the expectation in the Scalar of the transform, as the reference's fixture computes it
      *bo[r] = static_cast<float>(static_cast<Scalar>(static_cast<Scalar>(tf(r, 0)) * v[0] + static_cast<Scalar>(tf(r, 1)) * v[1] + static_cast<Scalar>(tf(r, 2)) * v[2] + tf(r, 3)));
      *bn[r] = static_cast<float>(static_cast<Scalar>(static_cast<Scalar>(tf(r, 0)) * n[0] + static_cast<Scalar>(tf(r, 1)) * n[1] + static_cast<Scalar>(tf(r, 2)) * n[2]));
    }
    p_n.push_back(a);
    p_n_trans.push_back(b);
    p_xyz.push_back(PointXYZ(a.x, a.y, a.z));
    p_xyz_trans.push_back(PointXYZ(b.x, b.y, b.z));
  }
  Indices indices(CLOUD_SIZE / 2);
  for (std::size_t i = 0; i < indices.size(); ++i) indices[i] = static_cast<index_t>(i * 2);
  auto xyz_near = [&](const auto& a, const auto& b) { return std::fabs(a.x - b.x) <= ABS_ERROR && std::fabs(a.y - b.y) <= ABS_ERROR && std::fabs(a.z - b.z) <= ABS_ERROR; };
  auto n_near = [&](const PointNormal& a, const PointNormal& b) {
    return std::fabs(a.normal_x - b.normal_x) <= ABS_ERROR && std::fabs(a.normal_y - b.normal_y) <= ABS_ERROR && std::fabs(a.normal_z - b.normal_z) <= ABS_ERROR;
  };
  {  // PointCloudXYZDense
    PointCloud<PointXYZ> p;
    transformPointCloud(p_xyz, p, tf);
    CHECK(p.width == p_xyz.width && p.height == p_xyz.height && p.is_dense == p_xyz.is_dense && p.size() == p_xyz.size());
    int bad = 0;
    for (std::size_t i = 0; i < p.size(); ++i) bad += !xyz_near(p[i], p_xyz_trans[i]);
    CHECK(bad == 0);
  }
  {  // PointCloudXYZDenseIndexed
    PointCloud<PointXYZ> p;
    transformPointCloud(p_xyz, indices, p, tf);
    CHECK(p.size() == indices.size() && p.width == indices.size() && p.height == 1);
    int bad = 0;
    for (std::size_t i = 0; i < p.size(); ++i) bad += !xyz_near(p[i], p_xyz_trans[i * 2]);
    CHECK(bad == 0);
  }
  {  // PointCloudXYZSparse
    PointCloud<PointXYZ> sparse = p_xyz, p;
    sparse.is_dense = false;
    sparse[0].x = std::numeric_limits<float>::quiet_NaN();
    transformPointCloud(sparse, p, tf);
    CHECK(p.width == sparse.width && p.height == sparse.height && !p.is_dense && p.size() == sparse.size());
    CHECK(!(std::isfinite(p[0].x) && std::isfinite(p[0].y) && std::isfinite(p[0].z)));
    int bad = 0;
    for (std::size_t i = 1; i < p.size(); ++i) bad += !(std::isfinite(p[i].x) && xyz_near(p[i], p_xyz_trans[i]));
    CHECK(bad == 0);
  }
  {  // PointCloudXYZRGBNormalDense: coordinates and normals move, the other fields are carried
    PointCloud<PointNormal> p;
    transformPointCloudWithNormals(p_n, p, tf);
    CHECK(p.width == p_n.width && p.height == p_n.height && p.size() == p_n.size());
    int bad = 0;
    for (std::size_t i = 0; i < p.size(); ++i) bad += !(xyz_near(p[i], p_n_trans[i]) && n_near(p[i], p_n_trans[i]) && p[i].curvature == p_n_trans[i].curvature);
    CHECK(bad == 0);
  }
  {  // PointCloudXYZRGBNormalDenseIndexed
    PointCloud<PointNormal> p;
    transformPointCloudWithNormals(p_n, indices, p, tf);
    CHECK(p.size() == indices.size() && p.width == indices.size() && p.height == 1);
    int bad = 0;
    for (std::size_t i = 0; i < p.size(); ++i) bad += !(xyz_near(p[i], p_n_trans[i * 2]) && n_near(p[i], p_n_trans[i * 2]) && p[i].curvature == p_n_trans[i * 2].curvature);
    CHECK(bad == 0);
  }
}

static int reference_centroid_tests(const char* bun0_pcd)
{
  centroid_body<float>();
  centroid_body<double>();
  for (unsigned seed = 1; seed <= 5; ++seed) { transforms_body<float>(seed); transforms_body<double>(seed); }
  {  // computeMeanAndCovariance
    PointCloud<PointXYZ> cloud;
    PointXYZ point;
    Indices indices;
    Eigen::Matrix3f cov;
    Eigen::Vector4f centroid;
    for (int i = 0; i < 9; ++i) cov[i] = 0.1f * static_cast<float>(i) - 0.3f;
    centroid[0] = 0.125f; centroid[1] = -0.5f; centroid[2] = 0.75f; centroid[3] = -0.25f;
    const Eigen::Matrix3f old_cov = cov;
    const Eigen::Vector4f old_centroid = centroid;
    cloud.is_dense = true;
    CHECK(computeMeanAndCovarianceMatrix(cloud, cov, centroid) == 0 && old_cov == cov && old_centroid == centroid);
    cloud.is_dense = false;
    CHECK(computeMeanAndCovarianceMatrix(cloud, cov, centroid) == 0 && old_cov == cov && old_centroid == centroid);
    point.x = point.y = point.z = std::numeric_limits<float>::quiet_NaN();
    cloud.push_back(point);
    CHECK(computeMeanAndCovarianceMatrix(cloud, cov, centroid) == 0 && old_cov == cov && old_centroid == centroid);
    cloud.push_back(point);
    indices.push_back(1);
    CHECK(computeMeanAndCovarianceMatrix(cloud, indices, cov, centroid) == 0 && old_cov == cov && old_centroid == centroid);
    cloud.clear();
    indices.clear();
    for (point.x = -1; point.x < 2; point.x += 2)
      for (point.y = -1; point.y < 2; point.y += 2)
        for (point.z = -1; point.z < 2; point.z += 2) cloud.push_back(point);
    cloud.is_dense = true;
    auto reset = [&] { for (int i = 0; i < 9; ++i) cov[i] = -100.f - static_cast<float>(i); centroid[0] = -100; centroid[1] = -101; centroid[2] = -102; };
    auto is_diag = [&](float a, float b, float c) {
      return cov(0, 0) == a && cov(1, 1) == b && cov(2, 2) == c && cov(0, 1) == 0 && cov(0, 2) == 0 && cov(1, 0) == 0 && cov(1, 2) == 0 && cov(2, 0) == 0 && cov(2, 1) == 0;
    };
    reset();   // eight points with (0, 0, 0) as centroid and the identity as covariance
    CHECK(computeMeanAndCovarianceMatrix(cloud, cov, centroid) == 8 && centroid[0] == 0 && centroid[1] == 0 && centroid[2] == 0 && is_diag(1, 1, 1));
    indices = {2, 3, 6, 7};
    reset();
    CHECK(computeMeanAndCovarianceMatrix(cloud, indices, cov, centroid) == 4 && centroid[0] == 0 && centroid[1] == 1 && centroid[2] == 0 && is_diag(1, 0, 1));
    point.x = point.y = point.z = std::numeric_limits<float>::quiet_NaN();
    cloud.push_back(point);
    cloud.is_dense = false;
    reset();
    CHECK(computeMeanAndCovarianceMatrix(cloud, cov, centroid) == 8 && centroid[0] == 0 && centroid[1] == 0 && centroid[2] == 0 && is_diag(1, 1, 1));
    indices.push_back(8);
    reset();
    CHECK(computeMeanAndCovarianceMatrix(cloud, indices, cov, centroid) == 4 && centroid[0] == 0 && centroid[1] == 1 && centroid[2] == 0 && is_diag(1, 0, 1));
  }
  {  // TEST (PCL, ConcatenatePoints) and (PCL, ConcatenateFields) — test/io/test_io.cpp:293-381
    std::mt19937 rng(5);
    auto rnd = [&] { return static_cast<float>(1024 * (rng() % 32768) / 32768.0); };
    PointCloud<PointXYZ> cloud_a, cloud_b, cloud_c;
    cloud_a.width = 5; cloud_b.width = 3;
    cloud_a.height = cloud_b.height = 1;
    cloud_a.points.resize(5); cloud_b.points.resize(3);
    for (auto& p : cloud_a.points) { p.x = rnd(); p.y = rnd(); p.z = rnd(); }
    for (auto& p : cloud_b.points) { p.x = rnd(); p.y = rnd(); p.z = rnd(); }
    cloud_c = cloud_a;
    cloud_c += cloud_b;
    CHECK(cloud_c.size() == cloud_a.size() + cloud_b.size() && cloud_c.width == cloud_a.width + cloud_b.width && cloud_c.height == 1);
    bool same = true;
    for (std::size_t i = 0; i < cloud_a.size(); ++i) same = same && cloud_c[i].x == cloud_a[i].x && cloud_c[i].y == cloud_a[i].y && cloud_c[i].z == cloud_a[i].z;
    for (std::size_t i = cloud_a.size(); i < cloud_c.size(); ++i) {
      const PointXYZ& b = cloud_b[i - cloud_a.size()];
      same = same && cloud_c[i].x == b.x && cloud_c[i].y == b.y && cloud_c[i].z == b.z;
    }
    CHECK(same);
    PointCloud<PointXYZ> fa;
    PointCloud<Normal> fb;
    PointCloud<PointNormal> fc;
    fa.width = fb.width = 5;
    fa.height = fb.height = 1;
    fa.points.resize(5); fb.points.resize(5);
    for (auto& p : fa) { p.x = rnd(); p.y = rnd(); p.z = rnd(); }
    for (auto& p : fb) { p.normal_x = rnd(); p.normal_y = rnd(); p.normal_z = rnd(); p.curvature = rnd(); }
    concatenateFields(fa, fb, fc);
    CHECK(fc.size() == fa.size() && fc.width == fa.width && fc.height == fa.height);
    same = true;
    for (std::size_t i = 0; i < fa.size() && i < fc.size(); ++i)
      same = same && fc[i].x == fa[i].x && fc[i].y == fa[i].y && fc[i].z == fa[i].z && fc[i].normal[0] == fb[i].normal[0] && fc[i].normal[1] == fb[i].normal[1] &&
             fc[i].normal[2] == fb[i].normal[2] && fc[i].curvature == fb[i].curvature;
    CHECK(same);
    fb.points.resize(4);   // sizes differ: refused, the output is left alone
    PointCloud<PointNormal> untouched;
    concatenateFields(fa, fb, untouched);
    CHECK(untouched.empty());
  }
  {  // demeanPointCloud on bun0
    PointCloud<PointXYZ> cloud, cloud_demean;
    if (io::loadPCDFile(bun0_pcd, cloud)) return 2;
    Eigen::Vector4f centroid;
    compute3DCentroid(cloud, centroid);
    CHECK(std::fabs(centroid[0] + 0.0290809) < 1e-4 && std::fabs(centroid[1] - 0.102653) < 1e-4 && std::fabs(centroid[2] - 0.027302) < 1e-4 && std::fabs(centroid[3] - 1) < 1e-4);
    auto near_xyz = [](const PointXYZ& p, float x, float y, float z) { return std::fabs(p.x - x) < 1e-4 && std::fabs(p.y - y) < 1e-4 && std::fabs(p.z - z) < 1e-4; };
    demeanPointCloud(cloud, centroid, cloud_demean);
    CHECK(cloud_demean.width == cloud.width && cloud_demean.height == cloud.height && cloud_demean.is_dense == cloud.is_dense && cloud_demean.size() == cloud.size());
    CHECK(near_xyz(cloud_demean[0], 0.034503f, 0.010837f, 0.013447f));
    CHECK(near_xyz(cloud_demean[cloud_demean.size() - 1], -0.048849f, 0.072507f, -0.071702f));
    Indices indices(cloud.size());
    for (int i = 0; i < static_cast<int>(indices.size()); ++i) indices[i] = i;
    demeanPointCloud(cloud, indices, centroid, cloud_demean);
    CHECK(cloud_demean.is_dense == cloud.is_dense && cloud_demean.width == indices.size() && cloud_demean.height == 1 && cloud_demean.size() == cloud.size());
    CHECK(near_xyz(cloud_demean[0], 0.034503f, 0.010837f, 0.013447f));
    CHECK(near_xyz(cloud_demean[cloud_demean.size() - 1], -0.048849f, 0.072507f, -0.071702f));
  }
  std::printf("%d checks, %d failures\n%s\n", g_checks, g_fail, g_fail ? "FAILED" : "PASSED");
  return g_fail ? 1 : 0;
}

int main(int argc, char** argv)
{
  if (argc == 4 && std::string(argv[1]) == "normals") return dump_normals(argv[2], argv[3]);
  if (argc == 3 && std::string(argv[1]) == "centroid") return reference_centroid_tests(argv[2]);
  {  // organised or not is decided by height alone
    PointCloud<PointXYZ> c;
    c.width = 640; c.height = 480;
    CHECK(c.isOrganized());
    c.height = 1;
    CHECK(!c.isOrganized());
  }
  {  // clear, insert (one / n), erase, emplace, emplace_back: an unorganised cloud of the new size
    PointCloud<PointXYZ> c;
    c.insert(c.end(), PointXYZ(1, 1, 1));
    CHECK(c.size() == 1 && c.width == 1 && !c.isOrganized());
    c.clear();
    CHECK(c.width == 0 && c.height == 0 && c.empty());
    c.insert(c.end(), 5, PointXYZ(1, 1, 1));
    CHECK(c.width == 5 && c.height == 1);
    c.erase(c.end() - 1);
    CHECK(c.width == 4 && c.height == 1);
    c.erase(c.begin(), c.end());
    CHECK(c.width == 0 && c.height == 1 && c.empty());
    c.emplace(c.end(), 1.f, 2.f, 3.f);
    CHECK(c.width == 1 && c.front().y == 2.f);
    PointXYZ& nb = c.emplace_back(4.f, 5.f, 6.f);
    CHECK(c.width == 2 && &nb == &c.back() && c.back().z == 6.f);
    std::vector<PointXYZ> more(3, PointXYZ(9, 9, 9));
    c.insert(c.end(), more.begin(), more.end());
    CHECK(c.width == 5 && c.height == 1);
  }
  {  // resize: count keeps a matching grid, (w, h) sets it, fills use the value
    PointCloud<PointXYZ> c;
    c.resize(640 * 360);
    CHECK(!c.isOrganized() && c.width == 640 * 360);
    c.resize(640, 480);
    CHECK(c.isOrganized() && c.width == 640 && c.size() == 640u * 480u);
    c.resize(640 * 480);  // same number of points: the grid stays
    CHECK(c.isOrganized() && c.width == 640);
    PointCloud<PointXYZ> d;
    d.resize(640 * 360, PointXYZ(1, 1, 1));
    CHECK(!d.isOrganized() && d.width == 640 * 360 && d.back().x == 1.f);
    d.resize(640, 480, PointXYZ(2, 2, 2));
    CHECK(d.isOrganized() && d.width == 640 && d.back().x == 2.f && d.front().x == 1.f);
  }
  {  // assign in all its forms; a width that does not divide the size gives one row
    PointCloud<PointXYZ> c;
    c.assign(640 * 360, PointXYZ(1, 1, 1));
    CHECK(!c.isOrganized() && c.width == 640 * 360);
    c.assign(640, 480, PointXYZ(1, 1, 1));
    CHECK(c.isOrganized() && c.width == 640);
    std::vector<PointXYZ> v(640 * 360, PointXYZ(2, 3, 4));
    c.assign(v.begin(), v.end());
    CHECK(!c.isOrganized() && c.width == 640 * 360);
    c.assign(v.begin(), v.end(), 640);
    CHECK(c.isOrganized() && c.width == 640 && c.height == 360);
    std::vector<PointXYZ> w(640 * 480, PointXYZ(7, 7, 7));
    c.assign(w.begin(), w.end(), 460);
    CHECK(!c.isOrganized() && c.width == 640 * 480);
    c.assign(w.begin(), w.end(), 0);
    CHECK(!c.isOrganized() && c.width == 640 * 480);
    c.assign({PointXYZ(3, 4, 5), PointXYZ(3, 4, 5), PointXYZ(3, 4, 5)});
    CHECK(!c.isOrganized() && c.width == 3);
    c.assign({PointXYZ(3, 4, 5), PointXYZ(3, 4, 5), PointXYZ(3, 4, 5), PointXYZ(3, 4, 5)}, 2);
    CHECK(c.isOrganized() && c.width == 2 && c.height == 2);
    c.assign({PointXYZ(3, 4, 5), PointXYZ(3, 4, 5), PointXYZ(3, 4, 5)}, 6);
    CHECK(!c.isOrganized() && c.width == 3);
  }
  {  // push_back drops the grid, the transient_ members leave width / height alone
    PointCloud<PointXYZ> c;
    c.push_back(PointXYZ(3, 4, 5));
    CHECK(!c.isOrganized() && c.width == 1);
    c.resize(80, 80, PointXYZ(1, 1, 1));
    CHECK(c.isOrganized());
    c.push_back(PointXYZ(3, 4, 5));
    CHECK(c.width == 80 * 80 + 1 && c.height == 1);
    PointCloud<PointXYZ> g = grid();
    g.transient_push_back(PointXYZ(2, 2, 2));
    CHECK(g.isOrganized() && g.width == 640 && g.size() == 640u * 480u + 1);
    g = grid();
    PointXYZ& e = g.transient_emplace_back(3.f, 3.f, 3.f);
    CHECK(g.isOrganized() && g.width == 640 && &e == &g.back());
    g = grid();
    g.transient_insert(g.end(), PointXYZ(1, 1, 1));
    CHECK(g.isOrganized() && g.size() == 640u * 480u + 1);
    g = grid();
    g.transient_insert(g.end(), 10, PointXYZ(1, 1, 1));
    CHECK(g.isOrganized() && g.size() == 640u * 480u + 10);
    g = grid();
    g.transient_emplace(g.end(), 4.f, 4.f, 4.f);
    CHECK(g.isOrganized() && g.size() == 640u * 480u + 1 && g.back().x == 4.f);
    g = grid();
    g.transient_erase(g.end() - 1);
    CHECK(g.isOrganized() && g.width == 640 && g.size() == 640u * 480u - 1);
    g = grid();
    g.transient_erase(g.begin(), g.end());
    CHECK(g.isOrganized() && g.width == 640 && g.size() == 0);
  }
  {  // concatenation: unorganised result, newest stamp, dense only if both are; 2-D access
    PointCloud<PointXYZ> g = grid(), u;
    g.header.stamp = 7;
    u.header.stamp = 3;
    PointCloud<PointXYZ>::concatenate(u, g);
    CHECK(!u.isOrganized() && u.width == 640 * 480 && u.header.stamp == 7);
    PointCloud<PointXYZ> out;
    PointCloud<PointXYZ>::concatenate(u, g, out);
    CHECK(!out.isOrganized() && out.width == 640 * 480 * 2);
    PointCloud<PointXYZ> sum = g + u;
    CHECK(!sum.isOrganized() && sum.width == 640 * 480 * 2);
    PointCloud<PointXYZ> both = g + g;
    CHECK(!both.isOrganized() && both.size() == 614400u && both.width == 614400u);
    PointCloud<PointXYZ> nd;
    nd.is_dense = false;
    u += nd;
    CHECK(!u.is_dense);
    bool threw = false;
    try {
      out.at(5, 5);
    }
    catch (const UnorganizedPointCloudException&) {
      threw = true;
    }
    CHECK(threw);
    const PointXYZ& last = g.at(static_cast<int>(g.width - 1), static_cast<int>(g.height - 1));
    CHECK(&last == &g.back());
    g(3, 2).x = 42.f;
    CHECK(g[2 * 640 + 3].x == 42.f);
  }
  {  // subset copy constructor and swap
    PointCloud<PointXYZ> c;
    for (int i = 0; i < 10; ++i) c.emplace_back(float(i), 0.f, 0.f);
    c.is_dense = false;
    c.sensor_origin_[0] = 1.5f;
    Indices idx = {7, 2, 9};
    PointCloud<PointXYZ> s(c, idx);
    CHECK(s.size() == 3 && s.width == 3 && s.height == 1 && !s.is_dense && s[0].x == 7.f && s[2].x == 9.f);
    CHECK(s.sensor_origin_[0] == 1.5f && s.sensor_orientation_ == Eigen::Quaternionf::Identity());
    PointCloud<PointXYZ> t = grid();
    t.swap(s);
    CHECK(t.size() == 3 && !t.is_dense && s.isOrganized() && s.width == 640 && s.is_dense);
    PointCloud<PointXYZ> z(4, 3, PointXYZ(5, 5, 5));
    CHECK(z.isOrganized() && z.size() == 12 && z.back().x == 5.f);
    std::size_t cnt = 0;
    for (auto it = z.crbegin(); it != z.crend(); ++it) ++cnt;
    CHECK(cnt == 12 && z.max_size() > 0);
  }
  {  // VoxelGrid's host-side grid geometry and leaf layout (b200::voxel_layout; impl/voxel_grid.hpp:612-650, 757-776)
    std::vector<PointXYZ> pts;
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    for (int i = 0; i < 5000; ++i) pts.emplace_back(rnd() * 2.f - 0.7f, rnd() * 1.5f - 1.1f, rnd() * 0.9f + 0.3f);
    pts[17].x = std::numeric_limits<float>::quiet_NaN();
    const float leaf = 0.1f, inv[3] = {1.0f / leaf, 1.0f / leaf, 1.0f / leaf};
    b200::VoxelLayout L;
    CHECK(b200::voxel_layout(pts, nullptr, pts.size(), false, inv, 0u, true, L));
    // brute force: cell of every finite point, distinct cells in ascending linear order
    std::map<int, int> count;
    for (std::size_t i = 0; i < pts.size(); ++i) {
      if (i == 17) continue;
      const int c0 = (int)std::floor(pts[i].x * inv[0]) - L.min_b[0], c1 = (int)std::floor(pts[i].y * inv[1]) - L.min_b[1],
                c2 = (int)std::floor(pts[i].z * inv[2]) - L.min_b[2];
      CHECK(c0 >= 0 && c0 < L.div_b[0] && c1 >= 0 && c1 < L.div_b[1] && c2 >= 0 && c2 < L.div_b[2]);
      ++count[c0 + c1 * L.div_b[0] + c2 * L.div_b[0] * L.div_b[1]];
    }
    CHECK(L.cells_kept == count.size());
    CHECK(L.leaf_layout.size() == (std::size_t)L.div_b[0] * L.div_b[1] * L.div_b[2]);
    CHECK(L.divb_mul[0] == 1 && L.divb_mul[1] == L.div_b[0] && L.divb_mul[2] == L.div_b[0] * L.div_b[1]);
    int rank = 0, bad = 0, filled = 0;
    for (const auto& kv : count) bad += L.leaf_layout[kv.first] != rank++;
    for (int v : L.leaf_layout) filled += v >= 0;
    CHECK(bad == 0 && filled == (int)count.size());
    // cells with fewer than min_points points are dropped and the survivors renumbered
    b200::VoxelLayout M;
    CHECK(b200::voxel_layout(pts, nullptr, pts.size(), false, inv, 3u, true, M));
    rank = 0; bad = 0;
    for (const auto& kv : count) {
      if (kv.second >= 3) bad += M.leaf_layout[kv.first] != rank++;
      else bad += M.leaf_layout[kv.first] != -1;
    }
    CHECK(bad == 0 && M.cells_kept == (std::size_t)rank && M.cells_kept < L.cells_kept);
    // an index subset; geometry only
    Indices sub;
    for (int i = 100; i < 900; i += 3) sub.push_back(i);
    b200::VoxelLayout S;
    CHECK(b200::voxel_layout(pts, sub.data(), sub.size(), true, inv, 0u, false, S));
    CHECK(S.leaf_layout.empty() && S.div_b[0] >= 1 && S.div_b[0] <= L.div_b[0] && S.min_b[0] >= L.min_b[0] && S.max_b[2] <= L.max_b[2]);
    // nothing finite -> no grid
    std::vector<PointXYZ> nan1(3);
    for (auto& p : nan1) p.x = std::numeric_limits<float>::infinity();
    CHECK(!b200::voxel_layout(nan1, nullptr, nan1.size(), false, inv, 0u, true, S));
    // the neighbour offset tables of voxel_grid.h:52-100
    const Eigen::MatrixXi half = getHalfNeighborCellIndices(), all = getAllNeighborCellIndices();
    CHECK(half.rows() == 3 && half.cols() == 13 && all.cols() == 27);
    CHECK(all(0, 13) == 0 && all(1, 13) == 0 && all(2, 13) == 0);
    std::set<int> seen;
    for (int j = 0; j < 27; ++j) seen.insert((all(0, j) + 1) + 3 * (all(1, j) + 1) + 9 * (all(2, j) + 1));
    CHECK(seen.size() == 27);
    for (int j = 0; j < 13; ++j) CHECK(all(0, 14 + j) == -half(0, j) && all(2, 14 + j) == -half(2, j));
  }
  {  // moments (common/centroid.h): centroid, one-pass covariance, demeaning; non-finite points of a non-dense cloud skipped
    PointCloud<PointXYZ> c;
    c.emplace_back(1.f, 2.f, 3.f);
    c.emplace_back(3.f, 2.f, 1.f);
    c.emplace_back(2.f, 5.f, 2.f);
    c.emplace_back(2.f, -1.f, 6.f);
    Eigen::Vector4f cen;
    CHECK(compute3DCentroid(c, cen) == 4 && cen[0] == 2.f && cen[1] == 2.f && cen[2] == 3.f && cen[3] == 1.f);
    Eigen::Matrix3f cov;
    Eigen::Vector4f cen2;
    CHECK(computeMeanAndCovarianceMatrix(c, cov, cen2) == 4);
    CHECK(std::fabs(cen2[0] - 2.f) < 1e-6f && std::fabs(cen2[1] - 2.f) < 1e-6f && std::fabs(cen2[2] - 3.f) < 1e-6f);
    // population covariance by hand: x: {-1,1,0,0} -> 0.5 ; y: {0,0,3,-3} -> 4.5 ; z: {0,-2,-1,3} -> 3.5 ; xz: (0-2+0+0)/4 = -0.5
    CHECK(std::fabs(cov(0, 0) - 0.5f) < 1e-6f && std::fabs(cov(1, 1) - 4.5f) < 1e-6f && std::fabs(cov(2, 2) - 3.5f) < 1e-6f);
    CHECK(std::fabs(cov(0, 2) + 0.5f) < 1e-6f && cov(2, 0) == cov(0, 2) && std::fabs(cov(1, 2) + 3.f) < 1e-6f && std::fabs(cov(0, 1)) < 1e-6f);
    PointCloud<PointXYZ> nd = c;
    nd.is_dense = false;
    nd.emplace_back(std::numeric_limits<float>::quiet_NaN(), 0.f, 0.f);
    Eigen::Vector4f cen3;
    CHECK(compute3DCentroid(nd, cen3) == 4 && cen3[1] == 2.f);
    Indices sub = {0, 1};
    Eigen::Matrix<double, 4, 1> cd;
    CHECK(compute3DCentroid(c, sub, cd) == 2 && cd[0] == 2.0 && cd[2] == 2.0);
    PointCloud<PointXYZ> dm;
    demeanPointCloud(c, cen, dm);
    CHECK(dm.size() == 4 && dm[0].x == -1.f && dm[3].z == 3.f);
    demeanPointCloud(c, sub, cen, dm);
    CHECK(dm.size() == 2 && dm.width == 2 && dm[1].x == 1.f && dm[1].z == -2.f);
    // a plane z = 0.25: normal +-(0,0,1), curvature 0, d = -+0.25; fewer than three points -> NaN
    PointCloud<PointXYZ> pl;
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 5; ++j) pl.emplace_back(0.1f * i, 0.2f * j, 0.25f);
    Eigen::Vector4f plane;
    float curv = 1.f;
    CHECK(computePointNormal(pl, plane, curv));
    CHECK(std::fabs(std::fabs(plane[2]) - 1.f) < 1e-6f && std::fabs(plane[0]) < 1e-6f && curv < 1e-6f && std::fabs(std::fabs(plane[3]) - 0.25f) < 1e-6f);
    flipNormalTowardsViewpoint(pl[0], 0.f, 0.f, 10.f, plane);
    CHECK(plane[2] > 0.99f && std::fabs(plane[3] + 0.25f) < 1e-6f);
    Eigen::Vector3f n3;
    n3[0] = 0.f; n3[1] = 0.f; n3[2] = 1.f;
    flipNormalTowardsViewpoint(pl[0], 0.f, 0.f, -10.f, n3);
    CHECK(n3[2] == -1.f);
    Indices two = {0, 1};
    CHECK(!computePointNormal(pl, two, plane, curv) && std::isnan(plane[0]) && std::isnan(curv));
  }
  {  // transformPointCloud[WithNormals] / transformPoint, getMinMax3D, removeNaNFromPointCloud, copyPointCloud
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    const float a = 0.3f;
    T(0, 0) = std::cos(a); T(0, 1) = -std::sin(a); T(1, 0) = std::sin(a); T(1, 1) = std::cos(a);
    T(0, 3) = 1.5f; T(1, 3) = -2.f; T(2, 3) = 0.25f;
    PointCloud<PointNormal> c;
    for (int i = 0; i < 12; ++i) c.emplace_back(0.1f * i, 1.f - 0.2f * i, 0.05f * i * i, 0.f, 0.6f, 0.8f, 0.01f * i);
    c.resize(4, 3);
    c.is_dense = false;
    c[5].x = std::numeric_limits<float>::quiet_NaN();
    c.header.frame_id = "lidar";
    PointCloud<PointNormal> t, tn;
    transformPointCloud(c, t, T);
    transformPointCloudWithNormals(c, tn, T);
    CHECK(t.width == 4 && t.height == 3 && !t.is_dense && t.header.frame_id == "lidar" && t.size() == 12);
    double worst = 0;
    for (int i = 0; i < 12; ++i) {
      if (i == 5) continue;
      const double x = c[i].x, y = c[i].y, z = c[i].z;
      worst = std::max(worst, std::fabs(t[i].x - (std::cos((double)a) * x - std::sin((double)a) * y + 1.5)));
      worst = std::max(worst, std::fabs(t[i].y - (std::sin((double)a) * x + std::cos((double)a) * y - 2.0)));
      worst = std::max(worst, std::fabs(t[i].z - (z + 0.25)));
      CHECK(t[i].normal_y == 0.6f && t[i].curvature == c[i].curvature);      // other fields copied, not rotated
      CHECK(std::fabs(tn[i].normal_x - (-std::sin(a) * 0.6f)) < 1e-6f && std::fabs(tn[i].normal_y - std::cos(a) * 0.6f) < 1e-6f && tn[i].normal_z == 0.8f);
      CHECK(tn[i].x == t[i].x && t[i].data[3] == 1.f && tn[i].data_n[3] == 0.f);
    }
    CHECK(worst < 1e-6);
    CHECK(std::isnan(t[5].x) && t[5].y == c[5].y);                            // a non-finite point is left as it was
    const PointNormal one = transformPoint(c[3], T);
    CHECK(one.x == t[3].x && one.z == t[3].z && one.normal_y == 0.6f);
    CHECK(transformPointWithNormal(c[3], T).normal_x == tn[3].normal_x);
    Eigen::Matrix4d Td = T.cast<double>();
    PointCloud<PointNormal> td;
    transformPointCloud(c, td, Td, false);
    CHECK(std::fabs(td[7].x - t[7].x) < 1e-6f && td[7].normal_y == 0.f);      // copy_all_fields = false: only the coordinates
    transformPointCloud(c, c, T);                                             // in place
    CHECK(c[3].x == t[3].x && c[11].z == t[11].z && c[3].normal_y == 0.6f);
    PointCloud<PointNormal> sub;
    transformPointCloud(t, Indices{2, 9}, sub, Eigen::Matrix4f::Identity());
    CHECK(sub.size() == 2 && sub.width == 2 && sub.height == 1 && sub[1].x == t[9].x && sub[1].curvature == t[9].curvature);
    Eigen::Vector4f mn, mx;
    getMinMax3D(t, mn, mx);
    float lo = 1e9f, hi = -1e9f;
    for (int i = 0; i < 12; ++i)
      if (i != 5) { lo = std::min(lo, t[i].y); hi = std::max(hi, t[i].y); }
    CHECK(mn[1] == lo && mx[1] == hi && mn[0] <= mx[0]);
    getMinMax3D(t, Indices{0, 1}, mn, mx);
    CHECK(mn[2] == std::min(t[0].z, t[1].z) && mx[2] == std::max(t[0].z, t[1].z));
    PointNormal pmn, pmx;
    getMinMax3D(t, pmn, pmx);
    CHECK(pmn.y == lo && pmx.y == hi);
    PointCloud<PointNormal> clean;
    Indices kept;
    removeNaNFromPointCloud(t, clean, kept);
    CHECK(clean.size() == 11 && clean.is_dense && clean.height == 1 && clean.width == 11 && kept.size() == 11 && kept[5] == 6 && clean[5].x == t[6].x);
    removeNaNFromPointCloud(clean, clean, kept);                             // dense: identity
    CHECK(clean.size() == 11 && kept[10] == 10);
    removeNaNFromPointCloud(t, t, kept);                                     // in place
    CHECK(t.size() == 11 && t.is_dense);
    PointCloud<PointXYZ> xyz;
    copyPointCloud(clean, xyz);
    CHECK(xyz.size() == 11 && xyz[4].y == clean[4].y && xyz.width == 11);
    PointCloud<PointNormal> picked;
    copyPointCloud(clean, Indices{10, 0}, picked);
    CHECK(picked.size() == 2 && picked[0].curvature == clean[10].curvature && picked.height == 1);
  }
  {  // pcl::search::Search: the index / cloud+index / batch / other-point-type forms through the two pure virtuals
    PointCloud<PointXYZ>::Ptr c(new PointCloud<PointXYZ>);
    for (int i = 0; i < 50; ++i) c->emplace_back(float(i), float(i % 7), float(i % 3));
    BruteForce bf;
    search::Search<PointXYZ>& sr = bf;
    CHECK(sr.getName() == "BruteForce" && sr.getSortedResults());
    sr.setInputCloud(c);
    Indices ki, ki2;
    std::vector<float> kd, kd2;
    CHECK(sr.nearestKSearch((*c)[10], 4, ki, kd) == 4 && ki[0] == 10 && kd[0] == 0.f);
    CHECK(sr.nearestKSearch(*c, 10, 4, ki2, kd2) == 4 && ki2 == ki);
    CHECK(sr.nearestKSearch(10, 4, ki2, kd2) == 4 && ki2 == ki);
    PointNormal pn;
    pn.x = 10.f; pn.y = 3.f; pn.z = 1.f;
    CHECK(sr.nearestKSearchT(pn, 4, ki2, kd2) == 4 && ki2 == ki && kd2 == kd);
    std::vector<Indices> bi;
    std::vector<std::vector<float>> bd;
    sr.nearestKSearch(*c, Indices{10, 20}, 4, bi, bd);
    CHECK(bi.size() == 2 && bi[0] == ki && bi[1][0] == 20);
    sr.nearestKSearch(*c, Indices(), 1, bi, bd);
    CHECK(bi.size() == 50 && bi[49][0] == 49);
    CHECK(sr.radiusSearch((*c)[10], 2.5, ki, kd) >= 1 && ki[0] == 10);
    CHECK(sr.radiusSearch(*c, 10, 2.5, ki2, kd2) == (int)ki.size() && ki2 == ki);
    CHECK(sr.radiusSearch(10, 2.5, ki2, kd2, 1) == 1 && ki2[0] == 10);
    CHECK(sr.radiusSearchT(pn, 2.5, ki2, kd2) == (int)ki.size() && ki2 == ki);
    sr.radiusSearch(*c, Indices{10}, 2.5, bi, bd);
    CHECK(bi.size() == 1 && bi[0] == ki);
    // with an index list the plain index form addresses the list (impl/search.hpp:93-108)
    IndicesPtr sub(new Indices{40, 41, 42, 43});
    sr.setInputCloud(c, sub);
    CHECK(sr.nearestKSearch(1, 1, ki, kd) == 1 && ki[0] == 41);
    Indices ui = {5, 6, 7};
    std::vector<float> ud = {3.f, 1.f, 2.f};
    BruteForce::sortResults(ui, ud);
    CHECK((ui == Indices{6, 7, 5}) && ud[0] == 1.f && ud[2] == 3.f);
    sr.setNumberOfThreads(0);
    CHECK(sr.getNumberOfThreads() == 1);
    // a consumer handed a searcher without a device side refuses it (no CPU fallback); a null pointer stays null
    search::Search<PointXYZ>::Ptr foreign(new BruteForce);
    CHECK(!search::deviceSearcher<PointXYZ>(foreign, "test"));
    CHECK(!search::deviceSearcher<PointXYZ>(search::Search<PointXYZ>::Ptr(), "test"));
  }
  {  // pcl::getRejectedQueryIndices / isBetterCorrespondence / operator<< (common/src/correspondence.cpp:46-94)
    Correspondences before, after;
    for (int q : {9, 2, 7, 4, 0}) before.emplace_back(q, q + 100, 0.5f * q);
    for (int q : {7, 0}) after.emplace_back(q, q + 100, 0.5f * q);
    Indices rej;
    getRejectedQueryIndices(before, after, rej);
    CHECK((rej == Indices{2, 4, 9}));
    getRejectedQueryIndices(before, Correspondences(), rej);
    CHECK((rej == Indices{9, 2, 7, 4, 0}));          // nothing survived: the query indices in their original order
    getRejectedQueryIndices(Correspondences(), after, rej);
    CHECK(rej.empty());
    CHECK(isBetterCorrespondence(before[0], before[1]) && !isBetterCorrespondence(before[4], before[1]));
    std::ostringstream os;
    os << before[1];
    CHECK(os.str() == "2 102 1");
  }
  std::printf("%d checks, %d failures\n%s\n", g_checks, g_fail, g_fail ? "FAILED" : "PASSED");
  return g_fail ? 1 : 0;
}
