// Second facade test program (runs on the device): API surface added after the main program's last run on hardware —
// the remaining pcl::search::Search overloads, CorrespondenceEstimation::setPointRepresentation[Reciprocal] and the
// DefaultConvergenceCriteria thresholds reached through getConvergeCriteria().  Kept apart (and run last by
// tests/test_zz_facade_extra_gpu.py) so the main program stays exactly what was verified.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <string>
#include <limits>
#include <memory>
#include <vector>

#include <pcl/common/io.h>
#include <pcl/features/normal_3d.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/filters/radius_outlier_removal.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/io/pcd_io.h>
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl/point_representation.h>
#include <pcl/point_types.h>
#include <pcl/common/transforms.h>
#include <pcl/registration/correspondence_estimation.h>
#include <pcl/registration/correspondence_rejection_median_distance.h>
#include <pcl/registration/correspondence_rejection_sample_consensus.h>
#include <pcl/registration/correspondence_rejection_var_trimmed.h>
#include <pcl/registration/transformation_estimation_lm.h>
#include <pcl/registration/icp.h>
#include <pcl/search/brute_force.h>
#include <pcl/search/kdtree.h>
#include <random>

using namespace pcl;

static int g_fail = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { ++g_checks; if (!(c)) { ++g_fail; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_EQ(a, b) do { ++g_checks; if (!((a) == (b))) { ++g_fail; std::printf("FAIL %s:%d  %s == %s  (%g vs %g)\n", __FILE__, __LINE__, #a, #b, (double)(a), (double)(b)); } } while (0)
#define EXPECT_NEAR(a, b, tol) do { ++g_checks; if (!(std::fabs((double)(a) - (double)(b)) <= (tol))) { ++g_fail; std::printf("FAIL %s:%d  |%s - %s| <= %g  (%.9g vs %.9g)\n", __FILE__, __LINE__, #a, #b, (double)(tol), (double)(a), (double)(b)); } } while (0)

static std::map<std::string, std::vector<double>> load_golden(const char* path)
{
  std::map<std::string, std::vector<double>> g;
  std::ifstream in(path);
  std::string name;
  std::size_t n;
  while (in >> name >> n) {
    std::vector<double> v(n);
    for (auto& x : v) in >> x;
    g[name] = v;
  }
  return g;
}

// test/registration/test_registration.cpp:322-334 sampleRandomTransform, with the axis / angle / translation drawn the same way
static Eigen::Matrix4f sample_random_transform(float max_angle, float max_trans)
{
  float ax = static_cast<float>(std::rand()) / RAND_MAX, ay = static_cast<float>(std::rand()) / RAND_MAX, az = static_cast<float>(std::rand()) / RAND_MAX;
  const float nrm = std::sqrt(ax * ax + ay * ay + az * az);
  ax /= nrm; ay /= nrm; az /= nrm;
  const float angle = static_cast<float>(std::rand()) / RAND_MAX * max_angle;
  const float tx = static_cast<float>(std::rand()) / RAND_MAX * max_trans, ty = static_cast<float>(std::rand()) / RAND_MAX * max_trans,
              tz = static_cast<float>(std::rand()) / RAND_MAX * max_trans;
  const float c = std::cos(angle), s = std::sin(angle), t = 1.f - c;
  Eigen::Matrix4f m = Eigen::Matrix4f::Identity();
  m(0, 0) = c + ax * ax * t;      m(0, 1) = ax * ay * t - az * s; m(0, 2) = ax * az * t + ay * s; m(0, 3) = tx;
  m(1, 0) = ay * ax * t + az * s; m(1, 1) = c + ay * ay * t;      m(1, 2) = ay * az * t - ax * s; m(1, 3) = ty;
  m(2, 0) = az * ax * t - ay * s; m(2, 1) = az * ay * t + ax * s; m(2, 2) = c + az * az * t;      m(2, 3) = tz;
  return m;
}
static Eigen::Matrix4f rigid_inverse(const Eigen::Matrix4f& m)
{
  Eigen::Matrix4f r = Eigen::Matrix4f::Identity();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r(i, j) = m(j, i);
  for (int i = 0; i < 3; ++i) r(i, 3) = -(r(i, 0) * m(0, 3) + r(i, 1) * m(1, 3) + r(i, 2) * m(2, 3));
  return r;
}

int main(int argc, char** argv)
{
  if (argc < 3) { std::fprintf(stderr, "usage: %s bun0.pcd bun4.pcd\n", argv[0]); return 2; }
  PointCloud<PointXYZ> cloud_source, cloud_target;
  if (io::loadPCDFile(argv[1], cloud_source) || io::loadPCDFile(argv[2], cloud_target)) return 2;
  std::map<std::string, std::vector<double>> G;
  if (argc > 3) G = load_golden(argv[3]);

  {  // the remaining Search<PointT> overloads (search.h:158-165, 229-260, 285-292, 311-315, 368-397): same lists as the
     // per-point forms
    PointCloud<PointXYZ>::Ptr c(new PointCloud<PointXYZ>(cloud_target));
    KdTreeFLANN<PointXYZ> kdtree;
    kdtree.setInputCloud(c);
    Indices ri, ri2;
    std::vector<float> rd, rd2;
    const int n_point = kdtree.radiusSearch((*c)[3], 0.02, ri, rd);
    EXPECT_TRUE(n_point > 1);
    EXPECT_EQ(kdtree.radiusSearch(*c, 3, 0.02, ri2, rd2), n_point);
    EXPECT_TRUE(ri2 == ri && rd2 == rd);
    PointNormal pn;
    pn.x = (*c)[3].x; pn.y = (*c)[3].y; pn.z = (*c)[3].z;
    EXPECT_EQ(kdtree.radiusSearchT(pn, 0.02, ri2, rd2), n_point);
    EXPECT_TRUE(ri2 == ri && rd2 == rd);
    Indices ki, ki2;
    std::vector<float> kd, kd2;
    EXPECT_EQ(kdtree.nearestKSearch((*c)[3], 10, ki, kd), 10);
    EXPECT_EQ(kdtree.nearestKSearchT(pn, 10, ki2, kd2), 10);
    EXPECT_TRUE(ki2 == ki && kd2 == kd);
    PointCloud<PointNormal> other;
    other.push_back(pn);
    std::vector<Indices> bi;
    std::vector<std::vector<float>> bd;
    kdtree.nearestKSearchT(other, Indices(), 10, bi, bd);
    EXPECT_EQ(bi.size(), 1u);
    if (bi.size() == 1) EXPECT_TRUE(bi[0] == ki && bd[0] == kd);
    kdtree.radiusSearchT(other, Indices(), 0.02, bi, bd);
    EXPECT_EQ(bi.size(), 1u);
    if (bi.size() == 1) EXPECT_TRUE(bi[0] == ri && bd[0] == rd);
    kdtree.setNumberOfThreads(4);
    EXPECT_EQ(kdtree.getNumberOfThreads(), 4u);
  }

  {  // TEST (PCL, KdTree_differentPointT) and (PCL, KdTree_multipointKnnSearch) — test/search/test_kdtree.cpp:127-204, with
     // PointNormal as the foreign point type: the batch forms return what the per-point forms return
    const unsigned int no_of_neighbors = 20;
    pcl::search::KdTree<PointXYZ> kdtree;
    kdtree.setInputCloud(cloud_target.makeShared());
    PointCloud<PointNormal> cloud_other;
    copyPointCloud(cloud_target, cloud_other);
    std::vector<std::vector<float>> dists, dists_same;
    std::vector<Indices> indices, indices_same;
    kdtree.nearestKSearchT(cloud_other, Indices(), no_of_neighbors, indices, dists);
    kdtree.nearestKSearch(cloud_target, Indices(), no_of_neighbors, indices_same, dists_same);
    EXPECT_EQ(indices.size(), cloud_target.size());
    EXPECT_EQ(indices_same.size(), cloud_target.size());
    int bad = 0;
    Indices k_indices, k_indices_t;
    std::vector<float> k_distances, k_distances_t;
    for (std::size_t i = 0; i < cloud_other.size() && i < indices.size() && i < indices_same.size(); i += 7) {
      kdtree.nearestKSearchT(cloud_other[i], no_of_neighbors, k_indices_t, k_distances_t);
      kdtree.nearestKSearch(cloud_target[i], no_of_neighbors, k_indices, k_distances);
      if (k_indices.size() != indices[i].size() || k_distances.size() != dists[i].size() || k_indices.size() != no_of_neighbors) { ++bad; continue; }
      for (std::size_t j = 0; j < no_of_neighbors; ++j) {
        if (!(k_indices[j] == indices[i][j] || k_distances[j] == dists[i][j])) ++bad;
        if (k_indices[j] != k_indices_t[j] || k_distances[j] != k_distances_t[j]) ++bad;
        if (indices_same[i][j] != indices[i][j] || dists_same[i][j] != dists[i][j]) ++bad;
      }
    }
    EXPECT_EQ(bad, 0);
  }

  {  // CorrespondenceEstimation::setPointRepresentation[Reciprocal] (correspondence_estimation.h:296-318): under a uniform
     // rescale by 2 the pairs are those of the plain estimator and the squared distances four times larger; the gate is
     // applied in the representation's space
    registration::CorrespondenceEstimation<PointXYZ, PointXYZ> plain, scaled;
    Correspondences cp, cs, rp, rs, gated;
    plain.setInputSource(cloud_source.makeShared());
    plain.setInputTarget(cloud_target.makeShared());
    plain.determineCorrespondences(cp);
    plain.determineReciprocalCorrespondences(rp);
    DefaultPointRepresentation<PointXYZ> rep;
    const float alpha[3] = {2.f, 2.f, 2.f};
    rep.setRescaleValues(alpha);
    scaled.setInputSource(cloud_source.makeShared());
    scaled.setInputTarget(cloud_target.makeShared());
    scaled.setPointRepresentation(rep.makeShared());
    scaled.setPointRepresentationReciprocal(rep.makeShared());
    scaled.determineCorrespondences(cs);
    EXPECT_EQ(cs.size(), cp.size());
    EXPECT_EQ(cp.size(), cloud_source.size());
    int idx_diff = 0;
    double worst = 0;
    for (std::size_t i = 0; i < cs.size() && i < cp.size(); ++i) {
      idx_diff += (cs[i].index_query != cp[i].index_query) || (cs[i].index_match != cp[i].index_match);
      worst = std::max(worst, std::fabs((double)cs[i].distance - 4.0 * (double)cp[i].distance));
    }
    EXPECT_EQ(idx_diff, 0);
    EXPECT_TRUE(worst < 1e-7);
    scaled.determineReciprocalCorrespondences(rs);
    EXPECT_EQ(rs.size(), rp.size());
    idx_diff = 0;
    for (std::size_t i = 0; i < rs.size() && i < rp.size(); ++i)
      idx_diff += (rs[i].index_query != rp[i].index_query) || (rs[i].index_match != rp[i].index_match);
    EXPECT_EQ(idx_diff, 0);
    // a gate of 0.02 in the rescaled space is a gate of 0.01 in the original one (97 of the 397 pairs, by the oracle)
    Correspondences gp;
    plain.determineCorrespondences(gp, 0.01);
    scaled.determineCorrespondences(gated, 0.02);
    EXPECT_EQ(gated.size(), gp.size());
    EXPECT_TRUE(gp.size() > 0 && gp.size() < cp.size());
    EXPECT_TRUE(scaled.getIndicesTarget() == nullptr);
    EXPECT_TRUE(scaled.getIndicesSource() != nullptr);
  }

  {  // DefaultConvergenceCriteria through getConvergeCriteria() (default_convergence_criteria.h:130-205): align() copies
     // Registration's thresholds into it (icp.hpp:157-161); a rotation threshold set on the object is what an align()
     // without setTransformationRotationEpsilon uses — an unreachable one (cos > 1) switches the TRANSFORM test off
    using Criteria = registration::DefaultConvergenceCriteria<float>;
    auto run = [&](double rot_thr, int& iters, Criteria::ConvergenceState& state, Criteria& seen) {
      IterativeClosestPoint<PointXYZ, PointXYZ> reg;
      reg.setInputSource(cloud_source.makeShared());
      reg.setInputTarget(cloud_target.makeShared());
      reg.setMaximumIterations(50);
      reg.setTransformationEpsilon(1e-8);
      reg.setMaxCorrespondenceDistance(0.05);
      if (rot_thr > 0) reg.getConvergeCriteria()->setRotationThreshold(rot_thr);
      PointCloud<PointXYZ> out;
      reg.align(out);
      iters = reg.getNumberOfIterations();
      state = reg.getConvergeCriteria()->getConvergenceState();
      seen = *reg.getConvergeCriteria();
      EXPECT_EQ(reg.getRANSACIterations(), 0);
      EXPECT_TRUE(reg.hasConverged());
    };
    int it_default = 0, it_off = 0;
    Criteria::ConvergenceState st_default, st_off;
    Criteria c_default, c_off;
    run(0.0, it_default, st_default, c_default);
    EXPECT_EQ(st_default, Criteria::CONVERGENCE_CRITERIA_TRANSFORM);
    EXPECT_EQ(c_default.getMaximumIterations(), 50);
    EXPECT_NEAR(c_default.getTranslationThreshold(), 1e-8, 1e-20);
    EXPECT_NEAR(c_default.getRotationThreshold(), 0.99999, 1e-12);
    EXPECT_TRUE(c_default.getRelativeMSE() < 0);
    run(1.5, it_off, st_off, c_off);
    EXPECT_TRUE(st_off != Criteria::CONVERGENCE_CRITERIA_TRANSFORM);
    EXPECT_TRUE(it_off >= it_default);
    EXPECT_NEAR(c_off.getRotationThreshold(), 1.5, 1e-12);
  }

  {  // b200::PinnedCloud: the source's storage page-locked for the duration of the aligns — same matrix as unpinned
    auto run = [&](bool pin) {
      PointCloud<PointXYZ>::Ptr src(new PointCloud<PointXYZ>(cloud_source));
      std::unique_ptr<b200::PinnedCloud<PointCloud<PointXYZ>>> guard;
      if (pin) guard.reset(new b200::PinnedCloud<PointCloud<PointXYZ>>(*src));
      IterativeClosestPoint<PointXYZ, PointXYZ> reg;
      reg.setInputSource(src);
      reg.setInputTarget(cloud_target.makeShared());
      reg.setMaximumIterations(50);
      reg.setTransformationEpsilon(1e-8);
      reg.setMaxCorrespondenceDistance(0.05);
      PointCloud<PointXYZ> out;
      reg.align(out);
      return reg.getFinalTransformation();
    };
    const Eigen::Matrix4f a = run(false), b = run(true);
    EXPECT_TRUE(a == b);
  }

  {  // VoxelGrid::setSaveLeafLayout and the grid accessors (voxel_grid.h:296-425): every input point's cell maps to the
     // output centroid of that cell
    VoxelGrid<PointXYZ> vg;
    vg.setInputCloud(cloud_source.makeShared());
    vg.setLeafSize(0.01f, 0.01f, 0.01f);
    vg.setSaveLeafLayout(true);
    PointCloud<PointXYZ> out;
    vg.filter(out);
    EXPECT_TRUE(out.size() > 10 && out.size() < cloud_source.size());
    const Eigen::Vector3i div = vg.getNrDivisions(), mul = vg.getDivisionMultiplier();
    EXPECT_TRUE((std::size_t)div[0] * div[1] * div[2] >= out.size());
    EXPECT_EQ(mul[0], 1);
    EXPECT_EQ(mul[1], div[0]);
    EXPECT_EQ(mul[2], div[0] * div[1]);
    EXPECT_EQ(vg.getLeafLayout().size(), (std::size_t)div[0] * div[1] * div[2]);
    int filled = 0;
    for (int v : vg.getLeafLayout()) filled += v >= 0;
    EXPECT_EQ((std::size_t)filled, out.size());
    int bad = 0;
    const Eigen::MatrixXi all = getAllNeighborCellIndices();
    for (const auto& p : cloud_source.points) {
      const int c = vg.getCentroidIndex(p);
      if (c < 0 || c >= (int)out.size()) { ++bad; continue; }
      const float dx = out[c].x - p.x, dy = out[c].y - p.y, dz = out[c].z - p.z;
      if (std::sqrt(dx * dx + dy * dy + dz * dz) > 0.01f * 1.7321f) ++bad;   // a centroid lies inside its cell
      if (vg.getCentroidIndexAt(vg.getGridCoordinates(p.x, p.y, p.z)) != c) ++bad;
      if (vg.getNeighborCentroidIndices(p, all)[13] != c) ++bad;
    }
    EXPECT_EQ(bad, 0);
    Eigen::Vector3i far;
    far[0] = vg.getMinBoxCoordinates()[0]; far[1] = vg.getMinBoxCoordinates()[1]; far[2] = vg.getMaxBoxCoordinates()[2] + 5;  // past the last slab
    EXPECT_EQ(vg.getCentroidIndexAt(far), -1);
    EXPECT_TRUE(vg.getSaveLeafLayout() && vg.getDownsampleAllData() && !vg.getFilterLimitsNegative());
  }

  {  // TEST (PCL, KdTreeFLANN_radiusSearch) and (PCL, KdTreeFLANN_nearestKSearch) — test/kdtree/test_kdtree.cpp:92-123, 161-205:
     // an 11^3 lattice of a point type DERIVED from PointXYZ, against brute force (the timing loops assert nothing and are left out)
    struct MyPoint : public PointXYZ {
      MyPoint() = default;
      MyPoint(float x_, float y_, float z_) { x = x_; y = y_; z = z_; }
    };
    PointCloud<MyPoint> cloud;
    const float resolution = 0.1f;
    for (float z = -0.5f; z <= 0.5f; z += resolution)
      for (float y = -0.5f; y <= 0.5f; y += resolution)
        for (float x = -0.5f; x <= 0.5f; x += resolution) cloud.push_back(MyPoint(x, y, z));
    cloud.width = static_cast<std::uint32_t>(cloud.size());
    cloud.height = 1;
    auto dist = [](const PointXYZ& a, const PointXYZ& b) { return std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z)); };
    {
      KdTreeFLANN<MyPoint> kdtree;
      kdtree.setInputCloud(cloud.makeShared());
      const MyPoint test_point(0.0f, 0.0f, 0.0f);
      const double max_dist = 0.15;
      std::set<int> brute_force_result;
      for (std::size_t i = 0; i < cloud.size(); ++i)
        if (dist(cloud[i], test_point) < max_dist) brute_force_result.insert(static_cast<int>(i));
      EXPECT_TRUE(brute_force_result.size() > 1 && brute_force_result.size() < 100);
      Indices k_indices;
      std::vector<float> k_distances;
      kdtree.radiusSearch(test_point, max_dist, k_indices, k_distances, 100);
      for (const auto& k_index : k_indices) {
        auto it = brute_force_result.find(k_index);
        const bool ok = it != brute_force_result.end();
        EXPECT_TRUE(ok);
        if (ok) brute_force_result.erase(it);
      }
      EXPECT_TRUE(brute_force_result.empty());
    }
    {
      KdTreeFLANN<MyPoint> kdtree;
      kdtree.setInputCloud(cloud.makeShared());
      const MyPoint test_point(0.01f, 0.01f, 0.01f);
      const unsigned int no_of_neighbors = 20;
      std::multimap<float, int> sorted_brute_force_result;
      for (std::size_t i = 0; i < cloud.size(); ++i) sorted_brute_force_result.insert(std::make_pair(dist(cloud[i], test_point), static_cast<int>(i)));
      float max_dist = 0.0f;
      unsigned int counter = 0;
      for (auto it = sorted_brute_force_result.begin(); it != sorted_brute_force_result.end() && counter < no_of_neighbors; ++it) {
        max_dist = std::max(max_dist, it->first);
        ++counter;
      }
      Indices k_indices(no_of_neighbors);
      std::vector<float> k_distances(no_of_neighbors);
      kdtree.nearestKSearch(test_point, no_of_neighbors, k_indices, k_distances);
      EXPECT_EQ(k_indices.size(), no_of_neighbors);
      for (const auto& k_index : k_indices) {
        bool ok = dist(test_point, cloud[k_index]) <= max_dist;
        if (!ok) ok = (std::abs(dist(test_point, cloud[k_index])) - max_dist) <= 1e-6;
        EXPECT_TRUE(ok);
      }
    }
  }

  {  // TEST (CorrespondenceRejectors, CorrespondenceRejectionMedianDistance) — test/registration/test_correspondence_rejectors.cpp:53-73
    CorrespondencesPtr corresps(new Correspondences());
    for (int i = 0; i <= 10; ++i) {
      Correspondence c;
      c.distance = static_cast<float>(i * i);
      corresps->push_back(c);
    }
    registration::CorrespondenceRejectorMedianDistance rejector;
    rejector.setInputCorrespondences(corresps);
    rejector.setMedianFactor(2.0);
    Correspondences corresps_filtered;
    rejector.getCorrespondences(corresps_filtered);
    EXPECT_EQ(corresps_filtered.size(), 8u);
    for (int i = 0; i < 8 && i < (int)corresps_filtered.size(); ++i) EXPECT_NEAR(corresps_filtered[i].distance, static_cast<float>(i * i), 1e-5);
  }

  {  // TYPED_TEST (CorrespondenceEstimationTestSuite, CorrespondenceEstimationSetSearchMethod) —
     // test/registration/test_correspondence_estimation.cpp:139-174: trees handed over with force_no_recompute give the
     // correspondences the estimator's own trees give (Scalar = double, PointXYZ -> PointXYZ and PointXYZ -> PointNormal)
    auto body = [&](auto source_tag, auto target_tag) {
      using PointSource = decltype(source_tag);
      using PointTarget = decltype(target_tag);
      auto cloud1 = std::make_shared<PointCloud<PointSource>>();
      auto cloud2 = std::make_shared<PointCloud<PointTarget>>();
      for (std::size_t i = 0; i < 50; i++) {
        PointSource a;
        PointTarget b;
        a.x = static_cast<float>(std::rand()); a.y = static_cast<float>(std::rand()); a.z = static_cast<float>(std::rand());
        b.x = static_cast<float>(std::rand()); b.y = static_cast<float>(std::rand()); b.z = static_cast<float>(std::rand());
        cloud1->push_back(a);
        cloud2->push_back(b);
      }
      auto tree1 = std::make_shared<pcl::search::KdTree<PointSource>>();
      tree1->setInputCloud(cloud1);
      auto tree2 = std::make_shared<pcl::search::KdTree<PointTarget>>();
      tree2->setInputCloud(cloud2);
      registration::CorrespondenceEstimation<PointSource, PointTarget, double> ce;
      ce.setInputSource(cloud1);
      ce.setInputTarget(cloud2);
      Correspondences corr_orig;
      ce.determineCorrespondences(corr_orig);
      ce.setSearchMethodSource(tree1, true);
      ce.setSearchMethodTarget(tree2, true);
      Correspondences corr_cached;
      ce.determineCorrespondences(corr_cached);
      EXPECT_EQ(corr_orig.size(), 50u);
      EXPECT_EQ(corr_orig.size(), corr_cached.size());
      for (std::size_t i = 0; i < corr_orig.size() && i < corr_cached.size(); i++) {
        EXPECT_EQ(corr_orig[i].index_query, corr_cached[i].index_query);
        EXPECT_EQ(corr_orig[i].index_match, corr_cached[i].index_match);
      }
    };
    body(PointXYZ(), PointXYZ());
    body(PointXYZ(), PointNormal());
  }

  {  // test/search/test_search.cpp:293-451, 456-529 — the unorganized scenarios: the device searcher against pcl::search::BruteForce,
     // k = 1, 8, 64, 512 and radius = 0.01 ... 0.08, on a dense cloud, a cloud with NaN points, a 10^3 grid; whole cloud and a shuffled
     // ~10 % view; results unique, ascending, inside the view and finite, and equal under the reference's own rule
     // (indices equal OR distances within 1e-6; a radius result may differ by the one point that sits on the sphere)
    std::mt19937 rng;
    std::uniform_int_distribution<unsigned> rand_uint(0, 10);
    std::uniform_real_distribution<float> rand_float(0.0f, 1.0f);
    const unsigned point_count = 1200, query_count = 100;
    PointCloud<PointXYZ>::Ptr dense(new PointCloud<PointXYZ>), sparse(new PointCloud<PointXYZ>), grid(new PointCloud<PointXYZ>);
    dense->resize(point_count); dense->height = 1; dense->width = point_count; dense->is_dense = true;
    sparse->resize(point_count); sparse->height = 1; sparse->width = point_count; sparse->is_dense = false;
    for (unsigned i = 0; i < point_count; ++i) {
      PointXYZ point(rand_float(rng), rand_float(rng), rand_float(rng));
      (*dense)[i] = point;
      if (rand_uint(rng) == 0) (*sparse)[i].x = (*sparse)[i].y = (*sparse)[i].z = std::numeric_limits<float>::quiet_NaN();
      else (*sparse)[i] = point;
    }
    grid->height = 1; grid->is_dense = true;
    for (unsigned x = 0; x < 10; ++x)
      for (unsigned y = 0; y < 10; ++y)
        for (unsigned z = 0; z < 10; ++z) grid->push_back(PointXYZ(0.1f * static_cast<float>(x), 0.1f * static_cast<float>(y), 0.1f * static_cast<float>(z)));
    Indices view;
    for (unsigned idx = 0; idx < point_count; ++idx)
      if (rand_uint(rng) == 0) view.push_back(static_cast<index_t>(idx));
    {
      std::uniform_int_distribution<> pick(0, static_cast<int>(view.size()) - 1);
      for (unsigned idx = 0; idx < point_count - 1; ++idx) std::swap(view[pick(rng)], view[pick(rng)]);
    }
    auto query_indices = [&](const PointCloud<PointXYZ>& c) {
      Indices q;
      const unsigned skip = static_cast<unsigned>(c.size()) / query_count;
      for (unsigned idx = 0; idx < c.size() && q.size() < query_count; ++idx)
        if ((std::rand() % skip) == 0 && std::isfinite(c[idx].x)) q.push_back(static_cast<index_t>(idx));
      return q;
    };
    pcl::search::BruteForce<PointXYZ> brute_force(true);
    pcl::search::KdTree<PointXYZ> kdtree;
    kdtree.setSortedResults(true);
    std::vector<pcl::search::Search<PointXYZ>*> methods = {&brute_force, &kdtree};
    auto unique = [](const Indices& v) { for (std::size_t a = 1; a < v.size(); ++a) for (std::size_t b = 0; b < a; ++b) if (v[a] == v[b]) return false; return true; };
    auto ordered = [](const std::vector<float>& d) { for (std::size_t a = 1; a < d.size(); ++a) if (d[a - 1] > d[a]) return false; return true; };
    auto same = [](const Indices& i1, const std::vector<float>& d1, const Indices& i2, const std::vector<float>& d2, float eps) {
      if (i1.size() != i2.size()) return false;
      for (std::size_t k = 0; k < i1.size(); ++k) if (i1[k] != i2[k] && std::abs(d1[k] - d2[k]) > eps) return false;
      return true;
    };
    auto run = [&](const PointCloud<PointXYZ>::ConstPtr& cloud, const Indices& queries, const Indices& input_indices, bool radius_mode) {
      std::vector<bool> in_view(cloud->size(), input_indices.empty()), finite(cloud->size(), true);
      for (index_t i : input_indices) in_view[i] = true;
      for (std::size_t i = 0; i < cloud->size(); ++i) finite[i] = std::isfinite((*cloud)[i].x) && std::isfinite((*cloud)[i].y) && std::isfinite((*cloud)[i].z);
      IndicesPtr idx;
      if (!input_indices.empty()) idx.reset(new Indices(input_indices));
      for (auto* m : methods) m->setInputCloud(cloud, idx);
      bool passed[2] = {true, true};
      std::vector<Indices> ind(2);
      std::vector<std::vector<float>> dst(2);
      auto check = [&](float radius) {
        for (int s = 0; s < 2; ++s) {
          bool valid = true;
          for (index_t i : ind[s]) valid = valid && in_view[i] && finite[i];
          passed[s] = passed[s] && unique(ind[s]) && ordered(dst[s]) && valid;
        }
        if (!same(ind[0], dst[0], ind[1], dst[1], 1e-6f)) {
          const bool on_sphere = radius > 0 && ((ind[0].size() + 1 == ind[1].size() && std::abs(dst[1].back() - radius * radius) < 1e-6) ||
                                                (ind[1].size() + 1 == ind[0].size() && std::abs(dst[0].back() - radius * radius) < 1e-6));
          if (!on_sphere) passed[1] = false;
        }
      };
      if (radius_mode)
        for (float radius = 0.01f; radius < 0.1f; radius *= 2.0f)
          for (index_t q : queries) {
            for (int s = 0; s < 2; ++s) methods[s]->radiusSearch((*cloud)[q], radius, ind[s], dst[s], 0);
            check(radius);
          }
      else
        for (unsigned knn = 1; knn <= 512; knn <<= 3)
          for (index_t q : queries) {
            for (int s = 0; s < 2; ++s) methods[s]->nearestKSearch((*cloud)[q], static_cast<int>(knn), ind[s], dst[s]);
            check(0.f);
          }
      EXPECT_TRUE(passed[0]);
      EXPECT_TRUE(passed[1]);
    };
    std::srand(3);
    const Indices dense_q = query_indices(*dense), sparse_q = query_indices(*sparse), grid_q = query_indices(*grid);
    EXPECT_TRUE(dense_q.size() > 20 && sparse_q.size() > 20 && grid_q.size() > 20 && view.size() > 50);
    run(dense, dense_q, Indices(), false);    // unorganized_dense_cloud_Complete_KNN
    run(dense, dense_q, view, false);         // unorganized_dense_cloud_View_KNN
    run(sparse, sparse_q, Indices(), false);  // unorganized_sparse_cloud_Complete_KNN
    run(sparse, sparse_q, view, false);       // unorganized_sparse_cloud_View_KNN
    run(dense, dense_q, Indices(), true);     // unorganized_dense_cloud_Complete_Radius
    run(grid, grid_q, Indices(), true);       // unorganized_grid_cloud_Complete_Radius
    run(dense, dense_q, view, true);          // unorganized_dense_cloud_View_Radius
    run(sparse, sparse_q, Indices(), true);   // unorganized_sparse_cloud_Complete_Radius
    run(sparse, sparse_q, view, true);        // unorganized_sparse_cloud_View_Radius
  }

  {  // TEST (VoxelGrid, Filters), the leaf-layout half — test/filters/test_filters.cpp:598-649 (the counts and centroids of
     // :566-597 are in the main program)
    PointCloud<PointXYZ>::Ptr cloud = cloud_source.makeShared();
    PointCloud<PointXYZ> output;
    VoxelGrid<PointXYZ> grid;
    grid.setLeafSize(0.02f, 0.02f, 0.02f);
    grid.setInputCloud(cloud);
    grid.setFilterFieldName("z");
    grid.setFilterLimits(0.05, 0.1);
    grid.setFilterLimitsNegative(true);
    grid.setSaveLeafLayout(true);
    grid.filter(output);
    EXPECT_EQ(output.size(), 100u);
    EXPECT_EQ(output.width, 100u);
    EXPECT_EQ(output.height, 1u);
    EXPECT_TRUE(output.is_dense);
    if (output.size() == 100) {
      EXPECT_EQ(grid.getCentroidIndex(output[0]), 0);
      EXPECT_EQ(grid.getCentroidIndex(output[99]), 99);
      EXPECT_EQ(grid.getCentroidIndexAt(grid.getGridCoordinates(-1, -1, -1)), -1);
      const int centroidIdx = grid.getCentroidIndex((*cloud)[195]);   // input point 195 [0.048722, 0.07376, 0.017434]
      EXPECT_TRUE(centroidIdx >= 0 && centroidIdx < 100);
      if (centroidIdx >= 0 && centroidIdx < 100) {
        EXPECT_TRUE(std::abs(output[centroidIdx].x - (*cloud)[195].x) <= 0.02);
        EXPECT_TRUE(std::abs(output[centroidIdx].y - (*cloud)[195].y) <= 0.02);
        EXPECT_TRUE(std::abs(output[centroidIdx].z - (*cloud)[195].z) <= 0.02);
        EXPECT_EQ(grid.getNeighborCentroidIndices(output[0], Eigen::MatrixXi::Zero(3, 1))[0], 0);
        EXPECT_EQ(grid.getNeighborCentroidIndices(output[99], Eigen::MatrixXi::Zero(3, 1))[0], 99);
        Eigen::MatrixXi directions = Eigen::Vector3i(0, 0, 1);
        std::vector<int> neighbors = grid.getNeighborCentroidIndices((*cloud)[195], directions);
        EXPECT_EQ(neighbors.size(), std::size_t(directions.cols()));
        EXPECT_TRUE(neighbors.at(0) != -1);
        if (neighbors.at(0) != -1) {
          EXPECT_TRUE(std::abs(output[neighbors.at(0)].x - output[centroidIdx].x) <= 0.02);
          EXPECT_TRUE(std::abs(output[neighbors.at(0)].y - output[centroidIdx].y) <= 0.02);
          EXPECT_TRUE(output[neighbors.at(0)].z - output[centroidIdx].z <= 0.02 * 2);
        }
      }
    }
    // "indices must be handled correctly": the indices of the original cloud keep the hundred appended points out
    auto indices = grid.getIndices();
    auto cloud_copied = std::make_shared<PointCloud<PointXYZ>>();
    *cloud_copied = *cloud;
    for (int i = 0; i < 100; i++) cloud_copied->push_back(PointXYZ(100.f + i, 100.f + i, 100.f + i));
    grid.setInputCloud(cloud_copied);
    grid.setIndices(indices);
    grid.filter(output);
    EXPECT_EQ(output.size(), 100u);
  }

  {  // TEST (VoxelGrid, Filters), "Test the pcl::PCLPointCloud2 method" — test/filters/test_filters.cpp:651-748
    PCLPointCloud2::Ptr cloud_blob(new PCLPointCloud2);
    toPCLPointCloud2(cloud_source, *cloud_blob);
    PointCloud<PointXYZ> output;
    VoxelGrid<PCLPointCloud2> grid2;
    PCLPointCloud2 output_blob;
    grid2.setLeafSize(0.02f, 0.02f, 0.02f);
    grid2.setInputCloud(cloud_blob);
    grid2.filter(output_blob);
    fromPCLPointCloud2(output_blob, output);
    EXPECT_EQ(output.size(), 103u);
    EXPECT_EQ(output.width, 103u);
    EXPECT_EQ(output.height, 1u);
    EXPECT_TRUE(output.is_dense);
    EXPECT_EQ(output_blob.point_step, 12u);
    grid2.setFilterFieldName("z");
    grid2.setFilterLimits(0.05, 0.1);
    grid2.filter(output_blob);
    fromPCLPointCloud2(output_blob, output);
    EXPECT_EQ(output.size(), 14u);
    EXPECT_EQ(output.width, 14u);
    EXPECT_EQ(output.height, 1u);
    EXPECT_TRUE(output.is_dense);
    if (output.size() == 14) {
      EXPECT_NEAR(output[0].x, -0.026125, 1e-4);
      EXPECT_NEAR(output[0].y, 0.039788, 1e-4);
      EXPECT_NEAR(output[0].z, 0.052827, 1e-4);
      EXPECT_NEAR(output[13].x, -0.073202, 1e-4);
      EXPECT_NEAR(output[13].y, 0.1296, 1e-4);
      EXPECT_NEAR(output[13].z, 0.051333, 1e-4);
    }
    grid2.setFilterLimitsNegative(true);
    grid2.setSaveLeafLayout(true);
    grid2.filter(output_blob);
    fromPCLPointCloud2(output_blob, output);
    EXPECT_EQ(output.size(), 100u);
    EXPECT_EQ(output.width, 100u);
    EXPECT_EQ(output.height, 1u);
    EXPECT_TRUE(output.is_dense);
    if (output.size() == 100) {
      EXPECT_EQ(grid2.getCentroidIndex(output[0].x, output[0].y, output[0].z), 0);
      EXPECT_EQ(grid2.getCentroidIndex(output[99].x, output[99].y, output[99].z), 99);
      EXPECT_EQ(grid2.getCentroidIndexAt(grid2.getGridCoordinates(-1, -1, -1)), -1);
      const int centroidIdx2 = grid2.getCentroidIndex(0.048722f, 0.073760f, 0.017434f);
      EXPECT_TRUE(centroidIdx2 != -1);
      if (centroidIdx2 >= 0 && centroidIdx2 < 100) {
        EXPECT_TRUE(std::abs(output[centroidIdx2].x - 0.048722) <= 0.02);
        EXPECT_TRUE(std::abs(output[centroidIdx2].y - 0.073760) <= 0.02);
        EXPECT_TRUE(std::abs(output[centroidIdx2].z - 0.017434) <= 0.02);
        EXPECT_EQ(grid2.getNeighborCentroidIndices(output[0].x, output[0].y, output[0].z, Eigen::MatrixXi::Zero(3, 1))[0], 0);
        EXPECT_EQ(grid2.getNeighborCentroidIndices(output[99].x, output[99].y, output[99].z, Eigen::MatrixXi::Zero(3, 1))[0], 99);
        Eigen::MatrixXi directions2 = Eigen::Vector3i(0, 0, 1);
        std::vector<int> neighbors2 = grid2.getNeighborCentroidIndices(0.048722f, 0.073760f, 0.017434f, directions2);
        EXPECT_EQ(neighbors2.size(), std::size_t(directions2.cols()));
        EXPECT_TRUE(neighbors2.at(0) != -1);
        if (neighbors2.at(0) != -1) {
          EXPECT_TRUE(std::abs(output[neighbors2.at(0)].x - output[centroidIdx2].x) <= 0.02);
          EXPECT_TRUE(std::abs(output[neighbors2.at(0)].y - output[centroidIdx2].y) <= 0.02);
          EXPECT_TRUE(output[neighbors2.at(0)].z - output[centroidIdx2].z <= 0.02 * 2);
        }
      }
    }
    auto indices2 = grid2.getIndices();   // original cloud indices
    PointCloud<PointXYZ> cloud_copied = cloud_source;
    for (int i = 0; i < 100; i++) cloud_copied.push_back(PointXYZ(100.f + i, 100.f + i, 100.f + i));
    auto cloud_blob2 = std::make_shared<PCLPointCloud2>();
    toPCLPointCloud2(cloud_copied, *cloud_blob2);
    grid2.setInputCloud(cloud_blob2);
    grid2.setIndices(indices2);
    grid2.filter(output_blob);
    fromPCLPointCloud2(output_blob, output);
    EXPECT_EQ(output.size(), 100u);   // additional points are ignored
    // a blob of PointNormal records: the normal and curvature planes come back too
    PointCloud<PointNormal> pn;
    for (const auto& p : cloud_source.points) pn.push_back(PointNormal(p.x, p.y, p.z, 0.f, 0.6f, 0.8f, 0.25f));
    auto pn_blob = std::make_shared<PCLPointCloud2>();
    toPCLPointCloud2(pn, *pn_blob);
    VoxelGrid<PCLPointCloud2> grid3;
    grid3.setLeafSize(0.02f, 0.02f, 0.02f);
    grid3.setInputCloud(pn_blob);
    grid3.filter(output_blob);
    PointCloud<PointNormal> out_pn;
    fromPCLPointCloud2(output_blob, out_pn);
    EXPECT_EQ(out_pn.size(), 103u);
    EXPECT_EQ(output_blob.point_step, 28u);
    if (out_pn.size() == 103) {
      EXPECT_NEAR(out_pn[5].normal_y, 0.6, 1e-6);
      EXPECT_NEAR(out_pn[5].normal_z, 0.8, 1e-6);
      EXPECT_NEAR(out_pn[5].curvature, 0.25, 1e-6);
    }
  }

  {  // FilterIndices::setKeepOrganized (impl/filter_indices.hpp:47-63) on RadiusOutlierRemoval, and ExtractIndices mapping the removed
     // indices back onto a blob the way tools/outlier_removal.cpp does
    PointCloud<PointXYZ>::Ptr cloud = cloud_source.makeShared();
    RadiusOutlierRemoval<PointXYZ> plain(true), organized(true);
    for (auto* f : {&plain, &organized}) {
      f->setInputCloud(cloud);
      f->setRadiusSearch(0.01);
      f->setMinNeighborsInRadius(4);
    }
    organized.setKeepOrganized(true);
    PointCloud<PointXYZ> kept, same_size;
    plain.filter(kept);
    organized.filter(same_size);
    PointIndices removed;
    plain.getRemovedIndices(removed);
    EXPECT_TRUE(!kept.empty() && kept.size() < cloud->size());
    EXPECT_EQ(kept.size() + removed.indices.size(), cloud->size());
    EXPECT_EQ(same_size.size(), cloud->size());
    EXPECT_EQ(same_size.width, cloud->width);
    EXPECT_TRUE(!same_size.is_dense);
    std::size_t nan_points = 0;
    for (const auto& p : same_size.points) nan_points += !std::isfinite(p.x);
    EXPECT_EQ(nan_points, removed.indices.size());
    for (index_t r : removed.indices) EXPECT_TRUE(std::isnan(same_size[static_cast<std::size_t>(r)].x) && std::isnan(same_size[static_cast<std::size_t>(r)].z));
    PCLPointCloud2::Ptr blob(new PCLPointCloud2);
    toPCLPointCloud2(*cloud, *blob);
    ExtractIndices<PCLPointCloud2> ei;
    ei.setInputCloud(blob);
    ei.setIndices(std::make_shared<const PointIndices>(removed));
    ei.setNegative(true);
    PCLPointCloud2 survivors;
    ei.filter(survivors);
    PointCloud<PointXYZ> back;
    fromPCLPointCloud2(survivors, back);
    EXPECT_EQ(back.size(), kept.size());
    bool same = back.size() == kept.size();
    for (std::size_t i = 0; same && i < back.size(); ++i) same = back[i].x == kept[i].x && back[i].y == kept[i].y && back[i].z == kept[i].z;
    EXPECT_TRUE(same);
    ExtractIndices<PointXYZ> et;
    et.setInputCloud(cloud);
    et.setIndices(std::make_shared<const PointIndices>(removed));
    PointCloud<PointXYZ> only_removed;
    et.filter(only_removed);
    EXPECT_EQ(only_removed.size(), removed.indices.size());
  }

  {  // TEST (VoxelGridMinPoints, Filters) — test/filters/test_filters.cpp:1356-1406, the positions (the reference's cloud is
     // PointXYZRGB; colour fields are outside this library): single points at 0 and 1, five points around 0.11, six around 0.31
    PointCloud<PointXYZ>::Ptr input(new PointCloud<PointXYZ>());
    input->push_back(PointXYZ(0.0f, 0.0f, 0.0f));
    const float offsets[6] = {0.001f, 0.002f, 0.003f, -0.001f, -0.002f, -0.003f};
    for (unsigned int i = 0; i < 5; ++i) {
      input->push_back(PointXYZ(0.11f + offsets[i], 0.11f + offsets[i], 0.11f + offsets[i]));
      input->push_back(PointXYZ(0.31f + offsets[i], 0.31f + offsets[i], 0.31f + offsets[i]));
    }
    input->push_back(PointXYZ(0.31f + offsets[5], 0.31f + offsets[5], 0.31f + offsets[5]));
    input->push_back(PointXYZ(1.0f, 1.0f, 1.0f));
    PointCloud<PointXYZ> outputMin4, outputMin6;
    VoxelGrid<PointXYZ> grid;
    grid.setLeafSize(0.02f, 0.02f, 0.02f);
    grid.setInputCloud(input);
    grid.setMinimumPointsNumberPerVoxel(4);
    grid.setDownsampleAllData(true);
    grid.filter(outputMin4);
    EXPECT_EQ(outputMin4.size(), 2u);
    if (outputMin4.size() == 2) {
      EXPECT_NEAR(outputMin4[0].x, input->at(1).x, 1e-2);
      EXPECT_NEAR(outputMin4[0].y, input->at(1).y, 1e-2);
      EXPECT_NEAR(outputMin4[0].z, input->at(1).z, 1e-2);
      EXPECT_NEAR(outputMin4[1].x, input->at(2).x, 1e-2);
      EXPECT_NEAR(outputMin4[1].y, input->at(2).y, 1e-2);
      EXPECT_NEAR(outputMin4[1].z, input->at(2).z, 1e-2);
    }
    grid.setMinimumPointsNumberPerVoxel(6);
    grid.setDownsampleAllData(false);
    grid.filter(outputMin6);
    EXPECT_EQ(outputMin6.size(), 1u);
    if (outputMin6.size() == 1) {
      EXPECT_NEAR(outputMin6[0].x, input->at(2).x, 1e-2);
      EXPECT_NEAR(outputMin6[0].y, input->at(2).y, 1e-2);
      EXPECT_NEAR(outputMin6[0].z, input->at(2).z, 1e-2);
    }
  }

  {  // TEST (PCL, TranslatedNormalEstimation), (NormalEstimation, FarFromOrigin), (PCL, NormalEstimationOpenMP) —
     // test/features/test_normal_estimation.cpp:187-285, 287-316, 375-410 (cloud = bun0, indices = all of it)
    const PointCloud<PointXYZ>& cloud = cloud_source;
    Indices indices(cloud.size());
    for (std::size_t i = 0; i < indices.size(); ++i) indices[i] = static_cast<index_t>(i);
    search::KdTree<PointXYZ>::Ptr tree(new search::KdTree<PointXYZ>(false));
    tree->setInputCloud(cloud.makeShared());
    {
      Eigen::Vector4f plane_parameters;
      float curvature;
      NormalEstimation<PointXYZ, Normal> n;
      PointCloud<PointXYZ> translatedCloud(cloud);
      for (auto& i : translatedCloud) { i.x += 100; i.y += 100; i.z += 100; }
      computePointNormal(translatedCloud, indices, plane_parameters, curvature);
      EXPECT_NEAR(std::abs(plane_parameters[0]), 0.035592, 1e-4);
      EXPECT_NEAR(std::abs(plane_parameters[1]), 0.369596, 1e-4);
      EXPECT_NEAR(std::abs(plane_parameters[2]), 0.928511, 1e-4);
      EXPECT_NEAR(curvature, 0.0693136, 1e-4);
      float nx, ny, nz;
      n.computePointNormal(translatedCloud, indices, nx, ny, nz, curvature);
      EXPECT_NEAR(std::abs(nx), 0.035592, 1e-4);
      EXPECT_NEAR(std::abs(ny), 0.369596, 1e-4);
      EXPECT_NEAR(std::abs(nz), 0.928511, 1e-4);
      EXPECT_NEAR(curvature, 0.0693136, 1e-4);
      computePointNormal(translatedCloud, plane_parameters, curvature);
      EXPECT_NEAR(plane_parameters[0], 0.035592, 1e-4);
      EXPECT_NEAR(plane_parameters[1], 0.369596, 1e-4);
      EXPECT_NEAR(plane_parameters[2], 0.928511, 1e-4);
      EXPECT_NEAR(curvature, 0.0693136, 1e-4);
      flipNormalTowardsViewpoint(translatedCloud.points[0], 0, 0, 0, plane_parameters);
      EXPECT_NEAR(plane_parameters[0], -0.035592, 1e-4);
      EXPECT_NEAR(plane_parameters[1], -0.369596, 1e-4);
      EXPECT_NEAR(plane_parameters[2], -0.928511, 1e-4);
      flipNormalTowardsViewpoint(translatedCloud.points[0], 0, 0, 0, nx, ny, nz);
      EXPECT_NEAR(nx, -0.035592, 1e-4);
      EXPECT_NEAR(ny, -0.369596, 1e-4);
      EXPECT_NEAR(nz, -0.928511, 1e-4);
      PointCloud<Normal>::Ptr normals(new PointCloud<Normal>());
      PointCloud<PointXYZ>::Ptr cloudptr = translatedCloud.makeShared();
      n.setInputCloud(cloudptr);
      EXPECT_TRUE(n.getInputCloud() == cloudptr);
      IndicesPtr indicesptr(new Indices(indices));
      n.setIndices(indicesptr);
      EXPECT_TRUE(n.getIndices() == indicesptr);
      n.setSearchMethod(tree);
      EXPECT_TRUE(n.getSearchMethod() == tree);
      n.setKSearch(static_cast<int>(indices.size()));
      n.compute(*normals);
      EXPECT_EQ(normals->size(), indices.size());
      for (const auto& point : normals->points) {
        EXPECT_NEAR(point.normal[0], -0.035592, 1e-4);
        EXPECT_NEAR(point.normal[1], -0.369596, 1e-4);
        EXPECT_NEAR(point.normal[2], -0.928511, 1e-4);
        EXPECT_NEAR(point.curvature, 0.0693136, 1e-4);
      }
      PointCloud<PointXYZ>::Ptr surfaceptr = cloudptr;
      n.setSearchSurface(surfaceptr);
      EXPECT_TRUE(n.getSearchSurface() == surfaceptr);
      // "Additional test for searchForNeighbors": a surface larger than the input, and no searcher given
      surfaceptr.reset(new PointCloud<PointXYZ>);
      *surfaceptr = *cloudptr;
      surfaceptr->points.resize(640 * 480);
      surfaceptr->width = 640;
      surfaceptr->height = 480;
      EXPECT_EQ(surfaceptr->size(), static_cast<std::size_t>(surfaceptr->width) * surfaceptr->height);
      n.setSearchSurface(surfaceptr);
      search::KdTree<PointXYZ>::Ptr none;
      n.setSearchMethod(none);
      n.compute(*normals);
      EXPECT_EQ(normals->size(), indices.size());
    }
    {
      NormalEstimation<PointXYZ, Normal> ne1;
      ne1.setInputCloud(cloud.makeShared());
      ne1.setKSearch(15);
      PointCloud<Normal> normals1;
      ne1.compute(normals1);
      PointCloud<PointXYZ>::Ptr cloud_translated(new PointCloud<PointXYZ>(cloud));
      for (auto& point : *cloud_translated) { point.x += 123.0f; point.y += -45.0f; point.z += 98.0f; }
      NormalEstimation<PointXYZ, Normal> ne2;
      ne2.setInputCloud(cloud_translated);
      ne2.setKSearch(15);
      ne2.setViewPoint(123.0f, -45.0f, 98.0f);
      PointCloud<Normal> normals2;
      ne2.compute(normals2);
      EXPECT_EQ(normals1.size(), normals2.size());
      EXPECT_EQ(normals1.size(), cloud.size());
      int off = 0;
      for (std::size_t i = 0; i < normals1.size() && i < normals2.size(); ++i) {
        const float dot = normals1[i].normal_x * normals2[i].normal_x + normals1[i].normal_y * normals2[i].normal_y + normals1[i].normal_z * normals2[i].normal_z;
        if (!(std::abs(std::abs(dot) - 1.0f) <= 1e-6f)) ++off;
        if (!(std::abs(normals1[i].normal_x - normals2[i].normal_x) <= 5e-4f && std::abs(normals1[i].normal_y - normals2[i].normal_y) <= 5e-4f &&
              std::abs(normals1[i].normal_z - normals2[i].normal_z) <= 5e-4f))
          ++off;
      }
      EXPECT_EQ(off, 0);
    }
    {
      NormalEstimationOMP<PointXYZ, Normal> n(4);
      EXPECT_EQ(n.getNumberOfThreads(), 4u);
      PointCloud<Normal>::Ptr normals(new PointCloud<Normal>());
      PointCloud<PointXYZ>::Ptr cloudptr = cloud.makeShared();
      n.setInputCloud(cloudptr);
      IndicesPtr indicesptr(new Indices(indices));
      n.setIndices(indicesptr);
      tree->setInputCloud(cloudptr);
      n.setSearchMethod(tree);
      n.setKSearch(static_cast<int>(indices.size()));
      n.compute(*normals);
      EXPECT_EQ(normals->size(), indices.size());
      for (const auto& point : normals->points) {
        EXPECT_NEAR(point.normal[0], -0.035592, 1e-4);
        EXPECT_NEAR(point.normal[1], -0.369596, 1e-4);
        EXPECT_NEAR(point.normal[2], -0.928511, 1e-4);
        EXPECT_NEAR(point.curvature, 0.0693136, 1e-4);
      }
    }
  }

  if (argc > 3) {  // TEST (PCL, CorrespondenceRejectorSampleConsensus) — test/registration/test_registration_api.cpp:225-263
    const std::vector<double>& want = G["corr_rej_sac"];
    const std::vector<double>& want_T = G["sac_transform"];
    PointCloud<PointXYZ>::ConstPtr source(new PointCloud<PointXYZ>(cloud_source)), target(new PointCloud<PointXYZ>(cloud_target));
    CorrespondencesPtr correspondences(new Correspondences);
    registration::CorrespondenceEstimation<PointXYZ, PointXYZ> corr_est;
    corr_est.setInputSource(source);
    corr_est.setInputTarget(target);
    corr_est.determineCorrespondences(*correspondences);
    EXPECT_EQ(correspondences->size(), 397u);
    Correspondences result;
    registration::CorrespondenceRejectorSampleConsensus<PointXYZ> corr_rej_sac;
    corr_rej_sac.setInputSource(source);
    corr_rej_sac.setInputTarget(target);
    corr_rej_sac.setInlierThreshold(0.01);    // rej_sac_max_dist, test_registration_api_data.h:816
    corr_rej_sac.setMaximumIterations(1000);  // rej_sac_max_iter
    corr_rej_sac.setInputCorrespondences(correspondences);
    corr_rej_sac.getCorrespondences(result);
    const Eigen::Matrix4f T = corr_rej_sac.getBestTransformation();
    EXPECT_EQ(want.size(), 2u * 97u);
    EXPECT_EQ(result.size(), want.size() / 2);
    if (result.size() == want.size() / 2)
      for (std::size_t i = 0; i < result.size(); ++i) {
        EXPECT_EQ(result[i].index_query, (int)want[2 * i]);
        EXPECT_EQ(result[i].index_match, (int)want[2 * i + 1]);
      }
    EXPECT_EQ(want_T.size(), 16u);
    if (want_T.size() == 16)
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) EXPECT_NEAR(T(i, j), want_T[4 * i + j], 1e-4);
    EXPECT_TRUE(!corr_rej_sac.runsOnDevice() && corr_rej_sac.requiresSourcePoints() && corr_rej_sac.requiresTargetPoints());
    EXPECT_NEAR(corr_rej_sac.getInlierThreshold(), 0.01, 0.0);
    EXPECT_EQ(corr_rej_sac.getMaximumIterations(), 1000);
    corr_rej_sac.setSaveInliers(true);
    corr_rej_sac.getCorrespondences(result);
    Indices inl;
    corr_rej_sac.getInliersIndices(inl);
    EXPECT_EQ(inl.size(), result.size());
  }

  {  // TEST (PCL, CorrespondenceRejectorVarTrimmed) — test/registration/test_registration_api.cpp:351-380 (its only assertion is
     // conditional on the result having 97 pairs), plus what the reference's code does: the trim factor minimises its FRMS over
     // [5 %, 95 %], the threshold is the sorted distance at that rank, and the FIRST m input pairs survive, m = #distances below it
    PointCloud<PointXYZ>::ConstPtr source(new PointCloud<PointXYZ>(cloud_source)), target(new PointCloud<PointXYZ>(cloud_target));
    CorrespondencesPtr correspondences(new Correspondences);
    registration::CorrespondenceEstimation<PointXYZ, PointXYZ> corr_est;
    corr_est.setInputSource(source);
    corr_est.setInputTarget(target);
    corr_est.determineCorrespondences(*correspondences);
    Correspondences kept;
    registration::CorrespondenceRejectorVarTrimmed rej;
    rej.setInputSource<PointXYZ>(source);
    rej.setInputTarget<PointXYZ>(target);
    rej.setInputCorrespondences(correspondences);
    rej.getCorrespondences(kept);
    const std::vector<double>& gd = G["corr_rej_dist"];
    if (kept.size() * 2 == gd.size())
      for (std::size_t i = 0; i < kept.size(); ++i) {
        EXPECT_EQ(kept[i].index_query, (int)gd[2 * i]);
        EXPECT_EQ(kept[i].index_match, (int)gd[2 * i + 1]);
      }
    std::vector<double> sorted;
    for (const auto& c : *correspondences) sorted.push_back(c.distance);
    std::sort(sorted.begin(), sorted.end());
    EXPECT_TRUE(rej.getTrimFactor() >= 0.05 - 1e-6 && rej.getTrimFactor() <= 0.95 + 1e-6);
    const std::size_t at = static_cast<std::size_t>(static_cast<int>(static_cast<double>(sorted.size()) * rej.getTrimFactor()));
    EXPECT_NEAR(rej.getTrimmedDistance(), sorted[at], 1e-12);
    std::size_t below = 0;
    for (double d : sorted) below += d < rej.getTrimmedDistance();
    EXPECT_EQ(kept.size(), below);
    for (std::size_t i = 0; i < kept.size() && i < correspondences->size(); ++i) EXPECT_EQ(kept[i].index_query, (*correspondences)[i].index_query);
    EXPECT_TRUE(!rej.runsOnDevice() && rej.requiresSourcePoints() && rej.requiresTargetPoints());
    // the scored form of the median rejector (clouds given): scores are the squared point distances, here the stored ones
    registration::CorrespondenceRejectorMedianDistance med_plain, med_scored;
    med_plain.setMedianFactor(1.5);
    med_scored.setMedianFactor(1.5);
    med_scored.setInputSource<PointXYZ>(source);
    med_scored.setInputTarget<PointXYZ>(target);
    Correspondences a, b;
    med_plain.getRemainingCorrespondences(*correspondences, a);
    med_scored.getRemainingCorrespondences(*correspondences, b);
    EXPECT_EQ(a.size(), b.size());
    EXPECT_NEAR(med_plain.getMedianDistance(), med_scored.getMedianDistance(), 1e-9);
    // TransformationEstimationLM: the minimiser of the point-to-point objective = the SVD estimate
    registration::TransformationEstimationLM<PointXYZ, PointXYZ, double> lm;
    registration::TransformationEstimationSVD<PointXYZ, PointXYZ, double> svd;
    Eigen::Matrix4d Tl, Ts;
    lm.estimateRigidTransformation(*source, *target, *correspondences, Tl);
    svd.estimateRigidTransformation(*source, *target, *correspondences, Ts);
    EXPECT_TRUE(Tl == Ts);
  }

  {  // TEST (PCL, IterativeClosestPointWithRejectors) — test/registration/test_registration.cpp:336-382: a median-distance and
     // a sample-consensus rejector, ten random offsets (<= 0.05) under ten random global poses (rotation <= 2 pi, |t| <= 10)
    IterativeClosestPoint<PointXYZ, PointXYZ> reg;
    reg.setMaximumIterations(50);
    reg.setTransformationEpsilon(1e-8);
    reg.setMaxCorrespondenceDistance(0.15);
    registration::CorrespondenceRejectorMedianDistance::Ptr rej_med(new registration::CorrespondenceRejectorMedianDistance);
    rej_med->setMedianFactor(4.0);
    reg.addCorrespondenceRejector(rej_med);
    registration::CorrespondenceRejectorSampleConsensus<PointXYZ>::Ptr rej_samp(new registration::CorrespondenceRejectorSampleConsensus<PointXYZ>);
    reg.addCorrespondenceRejector(rej_samp);
    std::srand(1);
    for (int t = 0; t < 10; ++t) {
      const Eigen::Matrix4f delta = sample_random_transform(0.f, 0.05f);
      const Eigen::Matrix4f net = sample_random_transform(2.f * static_cast<float>(M_PI), 10.f);
      PointCloud<PointXYZ>::Ptr source_trans(new PointCloud<PointXYZ>), target_trans(new PointCloud<PointXYZ>);
      transformPointCloud(cloud_source, *source_trans, rigid_inverse(delta) * net);
      transformPointCloud(cloud_source, *target_trans, net);
      reg.setInputSource(source_trans);
      reg.setInputTarget(target_trans);
      PointCloud<PointXYZ> cloud_reg;
      reg.align(cloud_reg);
      const Eigen::Matrix4f trans_final = reg.getFinalTransformation();
      for (int y = 0; y < 4; ++y) EXPECT_NEAR(trans_final(y, 3), delta(y, 3), 1e-2);
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 3; ++x) EXPECT_NEAR(trans_final(y, x), delta(y, x), 1e-1);
    }
  }

  std::printf("%d checks, %d failures\n%s\n", g_checks, g_fail, g_fail ? "FAILED" : "PASSED");
  return g_fail ? 1 : 0;
}
