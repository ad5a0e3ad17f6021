// test_facade.cpp — the reference's own registration / kdtree / filters / features tests, re-expressed against the
// facade (same class names, same calls, same expectations).  Each block cites the PCL test it mirrors.
// usage: test_facade <bun0.pcd> <bun4.pcd> <golden.txt>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include <pcl/features/normal_3d.h>
#include <pcl/filters/radius_outlier_removal.h>
#include <pcl/filters/statistical_outlier_removal.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/io/pcd_io.h>
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl/point_cloud.h>
#include <thread>

#include <pcl/conversions.h>
#include <pcl/point_representation.h>
#include <pcl/registration/transformation_validation_euclidean.h>
#include <pcl/point_types.h>
#include <pcl/registration/correspondence_estimation.h>
#include <pcl/registration/correspondence_estimation_backprojection.h>
#include <pcl/registration/correspondence_estimation_normal_shooting.h>
#include <pcl/registration/correspondence_rejection_surface_normal.h>
#include <pcl/registration/correspondence_rejection_distance.h>
#include <pcl/registration/correspondence_rejection_median_distance.h>
#include <pcl/registration/correspondence_rejection_one_to_one.h>
#include <pcl/registration/correspondence_rejection_trimmed.h>
#include <pcl/registration/icp.h>
#include <pcl/segmentation/extract_clusters.h>
#include <pcl/registration/transformation_estimation_point_to_plane_lls.h>
#include <pcl/registration/transformation_estimation_svd.h>
#include <pcl/registration/transformation_estimation_symmetric_point_to_plane_lls.h>

using namespace pcl;

static int g_fail = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { ++g_checks; if (!(c)) { ++g_fail; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_EQ(a, b) do { ++g_checks; if (!((a) == (b))) { ++g_fail; std::printf("FAIL %s:%d  %s == %s  (%g vs %g)\n", __FILE__, __LINE__, #a, #b, (double)(a), (double)(b)); } } while (0)
#define EXPECT_NEAR(a, b, tol) do { ++g_checks; if (!(std::fabs((double)(a) - (double)(b)) <= (tol))) { ++g_fail; std::printf("FAIL %s:%d  |%s - %s| <= %g  (%.9g vs %.9g)\n", __FILE__, __LINE__, #a, #b, (double)(tol), (double)(a), (double)(b)); } } while (0)
#define EXPECT_LT(a, b) do { ++g_checks; if (!((a) < (b))) { ++g_fail; std::printf("FAIL %s:%d  %s < %s  (%g vs %g)\n", __FILE__, __LINE__, #a, #b, (double)(a), (double)(b)); } } while (0)

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

int main(int argc, char** argv)
{
  if (argc < 4) { std::fprintf(stderr, "usage: %s bun0.pcd bun4.pcd golden.txt\n", argv[0]); return 2; }
  PointCloud<PointXYZ> cloud_source, cloud_target, cloud_reg;
  if (io::loadPCDFile(argv[1], cloud_source) || io::loadPCDFile(argv[2], cloud_target)) return 2;
  auto G = load_golden(argv[3]);
  EXPECT_EQ(cloud_source.size(), 397u);
  EXPECT_EQ(cloud_target.size(), 361u);

  {  // TEST (PCL, CorrespondenceEstimation) + Reciprocal — test/registration/test_registration_api.cpp:83-128
    CorrespondencesPtr corr(new Correspondences);
    registration::CorrespondenceEstimation<PointXYZ, PointXYZ> corr_est;
    corr_est.setInputSource(cloud_source.makeShared());
    corr_est.setInputTarget(cloud_target.makeShared());
    corr_est.determineCorrespondences(*corr);
    const auto& g = G["corr_original"];
    EXPECT_EQ(corr->size(), g.size() / 2);
    for (std::size_t i = 0; i < corr->size() && 2 * i + 1 < g.size(); ++i) {
      EXPECT_EQ((*corr)[i].index_query, (int)i);
      EXPECT_EQ((*corr)[i].index_match, (int)g[2 * i + 1]);
    }
    corr_est.determineReciprocalCorrespondences(*corr);
    const auto& r = G["corr_reciprocal"];
    EXPECT_EQ(corr->size(), r.size() / 2);
    for (std::size_t i = 0; i < corr->size() && 2 * i + 1 < r.size(); ++i) {
      EXPECT_EQ((*corr)[i].index_query, (int)r[2 * i]);
      EXPECT_EQ((*corr)[i].index_match, (int)r[2 * i + 1]);
    }
  }

  {  // TEST (PCL, CorrespondenceRejector{Distance,MedianDistance,OneToOne,Trimmed}) — test_registration_api.cpp:131-380
    CorrespondencesPtr corr(new Correspondences);
    registration::CorrespondenceEstimation<PointXYZ, PointXYZ> corr_est;
    corr_est.setInputSource(cloud_source.makeShared());
    corr_est.setInputTarget(cloud_target.makeShared());
    corr_est.determineCorrespondences(*corr);
    auto check = [&](registration::CorrespondenceRejector& rej, const char* key) {
      Correspondences out;
      rej.setInputCorrespondences(corr);
      rej.getCorrespondences(out);
      const auto& g = G[key];
      EXPECT_EQ(out.size(), g.size() / 2);
      for (std::size_t i = 0; i < out.size() && 2 * i + 1 < g.size(); ++i) {
        EXPECT_EQ(out[i].index_query, (int)g[2 * i]);
        EXPECT_EQ(out[i].index_match, (int)g[2 * i + 1]);
      }
    };
    registration::CorrespondenceRejectorDistance rd;
    rd.setMaximumDistance(0.01f);
    check(rd, "corr_rej_dist");
    registration::CorrespondenceRejectorMedianDistance rm;
    rm.setMedianFactor(0.5);
    check(rm, "corr_rej_median");
    EXPECT_NEAR(rm.getMedianDistance(), 0.000465391, 1e-4);
    registration::CorrespondenceRejectorOneToOne ro;
    check(ro, "corr_rej_one_to_one");
    registration::CorrespondenceRejectorTrimmed rt;
    rt.setOverlapRatio(0.5f);
    check(rt, "corr_rej_trimmed");
    // ICP with a rejector chain (test_registration.cpp:336-382 shape): converges to the same neighbourhood
    IterativeClosestPoint<PointXYZ, PointXYZ> reg;
    reg.setInputSource(cloud_source.makeShared());
    reg.setInputTarget(cloud_target.makeShared());
    reg.setMaximumIterations(50);
    reg.setTransformationEpsilon(1e-8);
    reg.setMaxCorrespondenceDistance(0.15);
    registration::CorrespondenceRejectorMedianDistance::Ptr rmp(new registration::CorrespondenceRejectorMedianDistance);
    rmp->setMedianFactor(4.0);
    reg.addCorrespondenceRejector(rmp);
    registration::CorrespondenceRejectorOneToOne::Ptr rop(new registration::CorrespondenceRejectorOneToOne);
    reg.addCorrespondenceRejector(rop);
    PointCloud<PointXYZ> aligned;
    reg.align(aligned);
    EXPECT_TRUE(reg.hasConverged());
    EXPECT_EQ(reg.getCorrespondenceRejectors().size(), 2u);
    EXPECT_LT(reg.getFitnessScore(), 1e-3);
  }

  {  // TEST (PCL, KdTreeFLANN_setPointRepresentation), default representation — test/kdtree/test_kdtree.cpp:226-262
    PointCloud<PointXYZ>::Ptr c(new PointCloud<PointXYZ>);
    const float pts[10][3] = {{86.6f, 42.1f, 92.4f}, {63.1f, 18.4f, 22.3f}, {35.5f, 72.5f, 37.3f}, {99.7f, 37.0f, 8.7f},
                              {22.4f, 84.1f, 64.0f}, {65.2f, 73.4f, 18.0f}, {60.4f, 57.1f, 4.5f},  {38.7f, 17.6f, 72.3f},
                              {14.2f, 95.7f, 34.7f}, {2.5f, 26.5f, 66.0f}};
    for (auto& p : pts) c->emplace_back(p[0], p[1], p[2]);
    KdTreeFLANN<PointXYZ> kdtree;
    kdtree.setInputCloud(c);
    Indices ki(10);
    std::vector<float> kd(10);
    EXPECT_EQ(kdtree.nearestKSearch(PointXYZ(50.f, 50.f, 50.f), 10, ki, kd), 10);
    const int gt_i[10] = {2, 7, 5, 1, 4, 6, 9, 0, 8, 3};
    const float gt_d[10] = {877.8f, 1674.7f, 1802.6f, 1937.5f, 2120.6f, 2228.8f, 3064.5f, 3199.7f, 3604.2f, 4344.8f};
    for (int i = 0; i < 10; ++i) { EXPECT_EQ(ki[i], gt_i[i]); EXPECT_NEAR(kd[i], gt_d[i], 0.1); }
    EXPECT_EQ(kdtree.nearestKSearch(PointXYZ(50.f, 50.f, 50.f), 25, ki, kd), 10);  // k clamped (kdtree_flann.hpp:241-245)
    EXPECT_EQ(ki.size(), 10u);
    // radius search: brute-force ball, sorted (test/kdtree/test_kdtree.cpp:92-123 shape)
    Indices ri;
    std::vector<float> rd;
    int nfound = kdtree.radiusSearch(PointXYZ(50.f, 50.f, 50.f), 45.0, ri, rd);
    int expect = 0;
    for (int i = 0; i < 10; ++i) expect += gt_d[i] < 45.0f * 45.0f ? 1 : 0;
    EXPECT_EQ(nfound, expect);
    for (int i = 0; i < nfound; ++i) EXPECT_EQ(ri[i], gt_i[i]);
    // the same test's second and third parts (test/kdtree/test_kdtree.cpp:255-289): a 2-D (x, y) representation, then the
    // default one with rescale values {1, 2, 3} — the reference's own ground truth
    {
      CustomPointRepresentation<PointXYZ>::Ptr xy(new CustomPointRepresentation<PointXYZ>(2, 0));
      kdtree.setPointRepresentation(xy);
      EXPECT_EQ(kdtree.nearestKSearch(PointXYZ(50.f, 50.f, 50.f), 10, ki, kd), 10);
      const int g2[10] = {6, 2, 5, 1, 7, 0, 4, 3, 9, 8};
      const float d2[10] = {158.6f, 716.5f, 778.6f, 1170.2f, 1177.5f, 1402.0f, 1924.6f, 2639.1f, 2808.5f, 3370.1f};
      for (int i = 0; i < 10; ++i) { EXPECT_EQ(ki[i], g2[i]); EXPECT_NEAR(kd[i], d2[i], 0.1); }
      DefaultPointRepresentation<PointXYZ> point_rep;
      const float alpha[3] = {1.0f, 2.0f, 3.0f};
      point_rep.setRescaleValues(alpha);
      kdtree.setPointRepresentation(point_rep.makeShared());
      EXPECT_EQ(kdtree.nearestKSearch(PointXYZ(50.f, 50.f, 50.f), 10, ki, kd), 10);
      const int g3[10] = {2, 9, 4, 7, 1, 5, 8, 0, 3, 6};
      const float d3[10] = {3686.9f, 6769.2f, 7177.0f, 8802.3f, 11071.5f, 11637.3f, 11742.4f, 17769.0f, 18497.3f, 18942.0f};
      for (int i = 0; i < 10; ++i) { EXPECT_EQ(ki[i], g3[i]); EXPECT_NEAR(kd[i], d3[i], 0.5); }
    }
  }

  {  // Registration::setPointRepresentation (registration.h:419-425): the searcher indexes rescaled coordinates, the
     // transform is estimated on the real ones.  A uniform rescale leaves every nearest neighbour where it was, so the
     // staged loop must land on the fused loop's transform; an anisotropic one still has to register the scans.
    PointCloud<PointXYZ>::Ptr src(new PointCloud<PointXYZ>(cloud_source)), tgt(new PointCloud<PointXYZ>(cloud_target));
    IterativeClosestPoint<PointXYZ, PointXYZ> plain, uniform, aniso;
    PointCloud<PointXYZ> out;
    for (auto* reg : {&plain, &uniform, &aniso}) {
      reg->setInputSource(src);
      reg->setInputTarget(tgt);
      reg->setMaximumIterations(50);
      reg->setTransformationEpsilon(1e-8);
    }
    plain.setMaxCorrespondenceDistance(0.05);
    plain.align(out);
    DefaultPointRepresentation<PointXYZ> rep2, rep3;
    const float a2[3] = {2.f, 2.f, 2.f}, a3[3] = {1.f, 1.5f, 0.75f};
    rep2.setRescaleValues(a2);
    rep3.setRescaleValues(a3);
    uniform.setPointRepresentation(rep2.makeShared());
    uniform.setMaxCorrespondenceDistance(0.1);  // distances are measured in the representation's space: 2 x 0.05
    uniform.align(out);
    EXPECT_TRUE(plain.hasConverged() && uniform.hasConverged());
    EXPECT_EQ(plain.getNumberOfIterations(), uniform.getNumberOfIterations());
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) EXPECT_NEAR(plain.getFinalTransformation()(r, c), uniform.getFinalTransformation()(r, c), 2e-5);
    aniso.setPointRepresentation(rep3.makeShared());
    aniso.setMaxCorrespondenceDistance(0.05);
    aniso.align(out);
    EXPECT_TRUE(aniso.hasConverged());
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) EXPECT_NEAR(plain.getFinalTransformation()(r, c), aniso.getFinalTransformation()(r, c), 2e-2);
  }

  {  // PCLPointCloud2 blobs (conversions.h:166-330): round trip by field NAME, and the blob form of setSourceNormals
    PointCloud<PointNormal> pn;
    for (int i = 0; i < 50; ++i) pn.emplace_back(0.1f * i, 1.f - 0.02f * i, 0.5f * i, 0.f, 0.6f, 0.8f, 0.01f * i);
    PCLPointCloud2 blob;
    toPCLPointCloud2(pn, blob);
    EXPECT_EQ(blob.point_step, 48u);
    EXPECT_EQ(blob.fields.size(), 7u);
    EXPECT_EQ(blob.width * blob.height, 50u);
    PointCloud<PointNormal> back;
    fromPCLPointCloud2(blob, back);
    EXPECT_EQ(back.size(), 50u);
    PointCloud<Normal> only_normals;   // a different point type picks its fields out of the same blob
    fromPCLPointCloud2(blob, only_normals);
    PointCloud<PointXYZ> only_xyz;
    fromPCLPointCloud2(blob, only_xyz);
    for (int i = 0; i < 50; i += 7) {
      EXPECT_EQ(back[i].x, pn[i].x); EXPECT_EQ(back[i].normal_z, pn[i].normal_z); EXPECT_EQ(back[i].curvature, pn[i].curvature);
      EXPECT_EQ(only_normals[i].normal_y, pn[i].normal_y); EXPECT_EQ(only_normals[i].curvature, pn[i].curvature);
      EXPECT_EQ(only_xyz[i].z, pn[i].z);
    }
    registration::CorrespondenceEstimationNormalShooting<PointXYZ, PointXYZ, Normal> ns;
    PCLPointCloud2::Ptr nb(new PCLPointCloud2(blob));
    ns.setSourceNormals(PCLPointCloud2::ConstPtr(nb));
    EXPECT_TRUE(ns.getSourceNormals() && ns.getSourceNormals()->size() == 50u);
    if (ns.getSourceNormals()) EXPECT_EQ((*ns.getSourceNormals())[10].normal_z, 0.8f);
  }

  {  // TransformationValidationEuclidean (transformation_validation_euclidean.h:77-263): identity on a cloud against itself
     // scores 0; the ICP result scores far below the identity on the bunny pair; isValid needs a threshold
    PointCloud<PointXYZ>::Ptr src(new PointCloud<PointXYZ>(cloud_source)), tgt(new PointCloud<PointXYZ>(cloud_target));
    registration::TransformationValidationEuclidean<PointXYZ, PointXYZ> tve;
    EXPECT_NEAR(tve.validateTransformation(tgt, tgt, Eigen::Matrix4f::Identity()), 0.0, 1e-12);
    IterativeClosestPoint<PointXYZ, PointXYZ> reg;
    PointCloud<PointXYZ> out;
    reg.setInputSource(src);
    reg.setInputTarget(tgt);
    reg.setMaximumIterations(50);
    reg.setTransformationEpsilon(1e-8);
    reg.setMaxCorrespondenceDistance(0.05);
    reg.align(out);
    const double s_id = tve.validateTransformation(src, tgt, Eigen::Matrix4f::Identity());
    const double s_icp = tve.validateTransformation(src, tgt, reg.getFinalTransformation());
    EXPECT_LT(s_icp, 0.25 * s_id);
    EXPECT_NEAR(s_icp, reg.getFitnessScore(), 1e-7);   // the same quantity by Registration::getFitnessScore
    EXPECT_TRUE(!tve.isValid(src, tgt, reg.getFinalTransformation()));  // threshold not set (:176-181)
    tve.setThreshold(2.0 * s_icp);
    EXPECT_TRUE(tve.isValid(src, tgt, reg.getFinalTransformation()));
    EXPECT_TRUE(!tve.isValid(src, tgt, Eigen::Matrix4f::Identity()));
    tve.setMaxRange(1e-12);
    EXPECT_EQ(tve.validateTransformation(src, tgt, reg.getFinalTransformation()), std::numeric_limits<double>::max());
  }

  Eigen::Matrix4f T_ref;
  {
    const auto& t = G["svd_Tref"];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T_ref(r, c) = (float)t[4 * r + c];
  }
  {  // TEST (PCL, TransformationEstimationSVD) — test_registration_api.cpp:383-423
    PointCloud<PointXYZ> source = cloud_target, target = cloud_target;
    for (auto& p : target) {  // pcl::transformPointCloud (SSE order, transforms.hpp:119-133)
      float x = p.x, y = p.y, z = p.z;
      p.x = T_ref(0, 0) * x + (T_ref(0, 1) * y + (T_ref(0, 2) * z + T_ref(0, 3)));
      p.y = T_ref(1, 0) * x + (T_ref(1, 1) * y + (T_ref(1, 2) * z + T_ref(1, 3)));
      p.z = T_ref(2, 0) * x + (T_ref(2, 1) * y + (T_ref(2, 2) * z + T_ref(2, 3)));
    }
    Eigen::Matrix4f T1, T2;
    const registration::TransformationEstimationSVD<PointXYZ, PointXYZ> est;
    est.estimateRigidTransformation(source, target, T1);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) EXPECT_NEAR(T1(r, c), T_ref(r, c), 2e-6);
    Correspondences corr;
    for (std::size_t i = 0; i < source.size(); ++i) corr.emplace_back((index_t)i, (index_t)i, 0.f);
    est.estimateRigidTransformation(source, target, corr, T2);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) EXPECT_EQ(T1(r, c), T2(r, c));
  }

  {  // TEST (PCL, TransformationEstimationPointToPlaneLLS) — test_registration_api.cpp:469-518
    registration::TransformationEstimationPointToPlaneLLS<PointNormal, PointNormal> est;
    PointCloud<PointNormal>::Ptr src(new PointCloud<PointNormal>), tgt(new PointCloud<PointNormal>);
    for (float x = -5.0f; x <= 5.0f; x += 0.5f)
      for (float y = -5.0f; y <= 5.0f; y += 0.5f) {
        PointNormal p;
        p.x = x; p.y = y;
        p.z = 0.1f * powf(x, 2.0f) + 0.2f * p.x * p.y - 0.3f * y + 1.0f;
        float nx = -0.2f * p.x - 0.2f, ny = 0.6f * p.y - 0.2f, nz = 1.0f;
        float m = std::sqrt(nx * nx + ny * ny + nz * nz);
        p.normal_x = nx / m; p.normal_y = ny / m; p.normal_z = nz / m;
        src->push_back(p);
      }
    Eigen::Matrix4f gt = Eigen::Matrix4f::Identity();
    const float rows[3][4] = {{0.9938f, 0.0988f, 0.0517f, 0.1f}, {-0.0997f, 0.9949f, 0.0149f, -0.2f}, {-0.05f, -0.02f, 0.9986f, 0.3f}};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) gt(r, c) = rows[r][c];
    *tgt = *src;
    for (auto& p : *tgt) {  // pcl::transformPointCloudWithNormals
      float x = p.x, y = p.y, z = p.z, a = p.normal_x, b = p.normal_y, c = p.normal_z;
      p.x = gt(0, 0) * x + (gt(0, 1) * y + (gt(0, 2) * z + gt(0, 3)));
      p.y = gt(1, 0) * x + (gt(1, 1) * y + (gt(1, 2) * z + gt(1, 3)));
      p.z = gt(2, 0) * x + (gt(2, 1) * y + (gt(2, 2) * z + gt(2, 3)));
      p.normal_x = gt(0, 0) * a + (gt(0, 1) * b + gt(0, 2) * c);
      p.normal_y = gt(1, 0) * a + (gt(1, 1) * b + gt(1, 2) * c);
      p.normal_z = gt(2, 0) * a + (gt(2, 1) * b + gt(2, 2) * c);
    }
    Eigen::Matrix4f est_T;
    est.estimateRigidTransformation(*src, *tgt, est_T);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) EXPECT_NEAR(est_T(i, j), gt(i, j), 1e-2);
    // TEST (PCL, TransformationEstimationSymmetricPointToPlaneLLS) — test_registration_api.cpp:663-713
    registration::TransformationEstimationSymmetricPointToPlaneLLS<PointNormal, PointNormal> sym;
    Eigen::Matrix4f sym_T;
    sym.estimateRigidTransformation(*src, *tgt, sym_T);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) EXPECT_NEAR(sym_T(i, j), gt(i, j), 1e-2);
  }

  {  // TEST (PCL, IterativeClosestPoint) — test/registration/test_registration.cpp:236-270
    IterativeClosestPoint<PointXYZ, PointXYZ> reg;
    reg.setInputSource(cloud_source.makeShared());
    reg.setInputTarget(cloud_target.makeShared());
    reg.setMaximumIterations(50);
    reg.setTransformationEpsilon(1e-8);
    reg.setMaxCorrespondenceDistance(0.05);
    reg.align(cloud_reg);
    EXPECT_EQ(cloud_reg.size(), cloud_source.size());
    Eigen::Matrix4f T = reg.getFinalTransformation();
    const auto& g = G["icp_bun0_bun4"];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) EXPECT_NEAR(T(r, c), g[4 * r + c], (r == 0 && c == 1) ? 1e-2 : 1e-3);
    EXPECT_EQ(T(3, 0), 0); EXPECT_EQ(T(3, 1), 0); EXPECT_EQ(T(3, 2), 0); EXPECT_EQ(T(3, 3), 1);
    EXPECT_TRUE(reg.hasConverged());
    {  // two threads, each inside a b200::Context::ThreadScope (own stream): concurrent aligns that share the target
       // tree of the process-wide context give the single-threaded matrix bit for bit
      pcl::search::KdTree<PointXYZ>::Ptr shared_tree(new pcl::search::KdTree<PointXYZ>);
      shared_tree->setInputCloud(cloud_target.makeShared());
      Eigen::Matrix4f Tt[2];
      bool conv[2] = {false, false};
      auto work = [&](int k) {
        b200::Context::ThreadScope scope;
        for (int rep = 0; rep < 3; ++rep) {
          IterativeClosestPoint<PointXYZ, PointXYZ> r2;
          r2.setInputSource(cloud_source.makeShared());
          r2.setInputTarget(cloud_target.makeShared());
          r2.setSearchMethodTarget(shared_tree, true);
          r2.setMaximumIterations(50);
          r2.setTransformationEpsilon(1e-8);
          r2.setMaxCorrespondenceDistance(0.05);
          PointCloud<PointXYZ> out2;
          r2.align(out2);
          Tt[k] = r2.getFinalTransformation();
          conv[k] = r2.hasConverged() && out2.size() == cloud_source.size();
        }
      };
      std::thread ta(work, 0), tb(work, 1);
      ta.join();
      tb.join();
      for (int k = 0; k < 2; ++k) {
        EXPECT_TRUE(conv[k]);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) EXPECT_EQ(Tt[k](r, c), T(r, c));
      }
    }
    {  // visualisation callback (registration.h:431-442): same result, one call per iteration with valid indices
      IterativeClosestPoint<PointXYZ, PointXYZ> regv;
      regv.setInputSource(cloud_source.makeShared());
      regv.setInputTarget(cloud_target.makeShared());
      regv.setMaximumIterations(50);
      regv.setTransformationEpsilon(1e-8);
      regv.setMaxCorrespondenceDistance(0.05);
      int calls = 0;
      bool ok_idx = true;
      std::function<void(const PointCloud<PointXYZ>&, const Indices&, const PointCloud<PointXYZ>&, const Indices&)> cb =
          [&](const PointCloud<PointXYZ>& s, const Indices& si, const PointCloud<PointXYZ>& t, const Indices& ti) {
            ++calls;
            ok_idx = ok_idx && si.size() == ti.size() && s.size() == cloud_source.size();
            for (std::size_t i = 0; i < si.size(); ++i) ok_idx = ok_idx && si[i] >= 0 && ti[i] >= 0 && (std::size_t)ti[i] < t.size();
          };
      regv.registerVisualizationCallback(cb);
      PointCloud<PointXYZ> outv;
      regv.align(outv);
      EXPECT_EQ(calls, regv.getNumberOfIterations());
      EXPECT_TRUE(ok_idx);
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) EXPECT_EQ(regv.getFinalTransformation()(r, c), T(r, c));
    }
    // the aligned cloud is the source under the final transform
    const PointXYZ& s = cloud_source[7];
    EXPECT_NEAR(cloud_reg[7].x, T(0, 0) * s.x + T(0, 1) * s.y + T(0, 2) * s.z + T(0, 3), 1e-6);
  }

  {  // TEST(PCL, ICP_translated) — test_registration.cpp:161-195
    PointCloud<PointXYZ>::Ptr in(new PointCloud<PointXYZ>(cloud_source)), out(new PointCloud<PointXYZ>(cloud_source));
    for (auto& p : *out) p.z += 0.2f;
    IterativeClosestPoint<PointXYZ, PointXYZ> icp;
    icp.setInputSource(in);
    icp.setInputTarget(out);
    icp.setMaximumIterations(50);
    PointCloud<PointXYZ> Final;
    icp.align(Final);
    EXPECT_TRUE(icp.hasConverged());
    EXPECT_LT(icp.getFitnessScore(), 1e-6);
    EXPECT_NEAR(icp.getFinalTransformation()(0, 0), 1.0, 2e-3);
    EXPECT_NEAR(icp.getFinalTransformation()(1, 1), 1.0, 2e-3);
    EXPECT_NEAR(icp.getFinalTransformation()(2, 2), 1.0, 2e-3);
    EXPECT_NEAR(icp.getFinalTransformation()(0, 3), 0.0, 2e-3);
    EXPECT_NEAR(icp.getFinalTransformation()(1, 3), 0.0, 2e-3);
    EXPECT_NEAR(icp.getFinalTransformation()(2, 3), 0.2, 2e-3);
  }

  {  // TEST(PCL, Registration_getFitnessScore_Indices) — test_registration.cpp:198-233
    PointCloud<PointXYZ>::Ptr in(new PointCloud<PointXYZ>), out(new PointCloud<PointXYZ>);
    in->push_back(PointXYZ(0, 0, 0)); in->push_back(PointXYZ(0, 1, 0)); in->push_back(PointXYZ(0, 0, 1)); in->push_back(PointXYZ(10, 0, 0));
    out->push_back(PointXYZ(0, 0, 0)); out->push_back(PointXYZ(0, 1, 0)); out->push_back(PointXYZ(0, 0, 1)); out->push_back(PointXYZ(10, 0, 0.5));
    IterativeClosestPoint<PointXYZ, PointXYZ> reg;
    reg.setInputSource(in);
    reg.setInputTarget(out);
    IndicesPtr ind(new Indices{0, 1, 2});
    reg.setIndices(ind);
    PointCloud<PointXYZ> fin;
    reg.align(fin);
    EXPECT_NEAR(reg.getFitnessScore(1.0, false), 0.0625, 1e-4);
    EXPECT_NEAR(reg.getFitnessScore(1.0, true), 0.0, 1e-4);
  }

  {  // TEST (PCL, IterativeClosestPointWithNormals) float and double — test_registration.cpp:272-323
    PointCloud<PointNormal>::Ptr src(new PointCloud<PointNormal>), tgt(new PointCloud<PointNormal>);
    NormalEstimation<PointXYZ, Normal> ne;
    ne.setKSearch(10);
    PointCloud<Normal> ns, nt;
    ne.setInputCloud(cloud_source.makeShared());
    ne.compute(ns);
    ne.setInputCloud(cloud_target.makeShared());
    ne.compute(nt);
    EXPECT_EQ(ns.size(), cloud_source.size());
    for (std::size_t i = 0; i < cloud_source.size(); ++i)
      src->push_back(PointNormal(cloud_source[i].x, cloud_source[i].y, cloud_source[i].z, ns[i].normal_x, ns[i].normal_y, ns[i].normal_z, ns[i].curvature));
    for (std::size_t i = 0; i < cloud_target.size(); ++i)
      tgt->push_back(PointNormal(cloud_target[i].x, cloud_target[i].y, cloud_target[i].z, nt[i].normal_x, nt[i].normal_y, nt[i].normal_z, nt[i].curvature));
    IterativeClosestPointWithNormals<PointNormal, PointNormal, float> regf;
    regf.setInputSource(src);
    regf.setInputTarget(tgt);
    regf.setMaximumIterations(50);
    regf.setTransformationEpsilon(1e-8);
    regf.setMaxCorrespondenceDistance(0.05);
    PointCloud<PointNormal> outf;
    regf.align(outf);
    EXPECT_EQ(outf.size(), src->size());
    EXPECT_TRUE(regf.hasConverged());
    EXPECT_LT(regf.getFitnessScore(), 0.001);
    EXPECT_NEAR(outf[0].curvature, (*src)[0].curvature, 0.0);  // fields other than xyz/normal are carried through
    IterativeClosestPointWithNormals<PointNormal, PointNormal, double> regd;
    regd.setInputSource(src);
    regd.setInputTarget(tgt);
    regd.setMaximumIterations(50);
    regd.setTransformationEpsilon(1e-8);
    regd.setMaxCorrespondenceDistance(0.05);
    PointCloud<PointNormal> outd;
    regd.align(outd);
    EXPECT_TRUE(regd.hasConverged());
    EXPECT_LT(regd.getFitnessScore(), 0.001);
    Eigen::Matrix4d Td = regd.getFinalTransformation();
    Eigen::Matrix4f Tf = regf.getFinalTransformation();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) EXPECT_NEAR(Td(r, c), Tf(r, c), 5e-3);
    // symmetric objective (icp.h:381-393): same neighbourhood of the solution
    IterativeClosestPointWithNormals<PointNormal, PointNormal, float> regs;
    regs.setUseSymmetricObjective(true);
    regs.setInputSource(src);
    regs.setInputTarget(tgt);
    regs.setMaximumIterations(50);
    regs.setTransformationEpsilon(1e-8);
    regs.setMaxCorrespondenceDistance(0.05);
    PointCloud<PointNormal> outs;
    regs.align(outs);
    EXPECT_TRUE(regs.hasConverged());
    EXPECT_TRUE(regs.getUseSymmetricObjective());
    EXPECT_LT(regs.getFitnessScore(), 0.001);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) EXPECT_NEAR(regs.getFinalTransformation()(r, c), Tf(r, c), 0.1);  // a different objective: same basin only
    // cached target tree: setSearchMethodTarget(tree, force_no_recompute) — test_registration.cpp:513-600 shape
    search::KdTree<PointNormal>::Ptr tree(new search::KdTree<PointNormal>);
    tree->setInputCloud(tgt);
    IterativeClosestPointWithNormals<PointNormal, PointNormal, float> regc;
    regc.setInputSource(src);
    regc.setInputTarget(tgt);
    regc.setSearchMethodTarget(tree, true);
    regc.setMaximumIterations(50);
    regc.setTransformationEpsilon(1e-8);
    regc.setMaxCorrespondenceDistance(0.05);
    PointCloud<PointNormal> outc;
    regc.align(outc);
    EXPECT_LT(regc.getFitnessScore(), 0.005);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) EXPECT_EQ(regc.getFinalTransformation()(r, c), Tf(r, c));
  }

  {  // TEST (VoxelGrid, Filters) — test/filters/test_filters.cpp:566-580
    PointCloud<PointXYZ> output;
    VoxelGrid<PointXYZ> grid;
    grid.setLeafSize(0.02f, 0.02f, 0.02f);
    grid.setInputCloud(cloud_source.makeShared());
    grid.filter(output);
    EXPECT_EQ(output.size(), 103u);
    EXPECT_EQ(output.width, 103u);
    EXPECT_EQ(output.height, 1u);
    EXPECT_TRUE(output.is_dense);
    grid.setFilterFieldName("z");           // test_filters.cpp:582-603
    grid.setFilterLimits(0.05, 0.1);
    grid.filter(output);
    EXPECT_EQ(output.size(), 14u);
    EXPECT_EQ(output.width, 14u);
    EXPECT_NEAR(output[0].x, -0.026125, 1e-4);
    EXPECT_NEAR(output[0].y, 0.039788, 1e-4);
    EXPECT_NEAR(output[0].z, 0.052827, 1e-4);
    EXPECT_NEAR(output[13].x, -0.073202, 1e-4);
    EXPECT_NEAR(output[13].y, 0.1296, 1e-4);
    EXPECT_NEAR(output[13].z, 0.051333, 1e-4);
    grid.setFilterLimitsNegative(true);
    grid.filter(output);
    EXPECT_EQ(output.size(), 100u);
    grid.setFilterFieldName("");
    grid.setLeafSize(1e-5f, 1e-5f, 1e-5f);  // overflow guard: input returned unfiltered (voxel_grid.hpp:620-629)
    grid.filter(output);
    EXPECT_EQ(output.size(), cloud_source.size());
  }

  {  // TEST (RadiusOutlierRemoval, Filters) — test/filters/test_filters.cpp:1494-1515
    PointCloud<PointXYZ> out, out_neg;
    RadiusOutlierRemoval<PointXYZ> outrem;
    outrem.setInputCloud(cloud_source.makeShared());
    outrem.setRadiusSearch(0.02);
    outrem.setMinNeighborsInRadius(14);
    outrem.filter(out);
    EXPECT_EQ(out.size(), 307u);
    EXPECT_EQ(out.width, 307u);
    EXPECT_TRUE(out.is_dense);
    outrem.setNegative(true);
    outrem.filter(out_neg);
    EXPECT_EQ(out_neg.size(), 90u);
  }
  {  // TEST (StatisticalOutlierRemoval, Filters) — test/filters/test_filters.cpp:1587-1613
    PointCloud<PointXYZ> output;
    StatisticalOutlierRemoval<PointXYZ> outrem(true);
    outrem.setInputCloud(cloud_source.makeShared());
    outrem.setMeanK(50);
    outrem.setStddevMulThresh(1.0);
    outrem.filter(output);
    EXPECT_EQ(output.size(), 352u);
    EXPECT_EQ(output.width, 352u);
    EXPECT_TRUE(output.is_dense);
    EXPECT_NEAR(output[output.size() - 1].x, -0.034667, 1e-4);
    EXPECT_NEAR(output[output.size() - 1].y, 0.15131, 1e-4);
    EXPECT_NEAR(output[output.size() - 1].z, -0.00071029, 1e-4);
    EXPECT_EQ(outrem.getRemovedIndices()->size(), cloud_source.size() - 352);
    outrem.setNegative(true);
    outrem.filter(output);
    EXPECT_EQ(output.size(), cloud_source.size() - 352);
    EXPECT_NEAR(output[output.size() - 1].x, -0.07793, 1e-4);
    EXPECT_NEAR(output[output.size() - 1].y, 0.17516, 1e-4);
    EXPECT_NEAR(output[output.size() - 1].z, -0.0444, 1e-4);
  }

  {  // TEST (PCL, NormalEstimation) — test/features/test_normal_estimation.cpp:128-163: k = all points
    NormalEstimation<PointXYZ, Normal> n;
    PointCloud<Normal> normals;
    search::KdTree<PointXYZ>::Ptr tree(new search::KdTree<PointXYZ>);
    auto cptr = cloud_source.makeShared();
    n.setInputCloud(cptr);
    n.setSearchMethod(tree);
    n.setKSearch(static_cast<int>(cloud_source.size()));
    n.compute(normals);
    EXPECT_EQ(normals.size(), cloud_source.size());
    const auto& g = G["normal_bun0"];
    for (const auto& p : normals.points) {
      EXPECT_NEAR(p.normal_x, -g[0], 1e-4);
      EXPECT_NEAR(p.normal_y, -g[1], 1e-4);
      EXPECT_NEAR(p.normal_z, -g[2], 1e-4);
      EXPECT_NEAR(p.curvature, g[4], 1e-4);
    }
  }


  {  // radius-search normals: a radius that holds the whole cloud is the k = all-points case above
    NormalEstimation<PointXYZ, Normal> n;
    PointCloud<Normal> normals;
    n.setInputCloud(cloud_source.makeShared());
    n.setRadiusSearch(10.0);
    n.compute(normals);
    EXPECT_EQ(normals.size(), cloud_source.size());
    EXPECT_TRUE(normals.is_dense);
    const auto& g = G["normal_bun0"];
    for (const auto& p : normals.points) {
      EXPECT_NEAR(p.normal_x, -g[0], 1e-4);
      EXPECT_NEAR(p.normal_y, -g[1], 1e-4);
      EXPECT_NEAR(p.normal_z, -g[2], 1e-4);
      EXPECT_NEAR(p.curvature, g[4], 1e-4);
    }
    n.setRadiusSearch(1e-4);  // nobody has 3 neighbours: NaN normals, is_dense = false (normal_3d.hpp:62-69)
    n.compute(normals);
    EXPECT_EQ(normals.size(), cloud_source.size());
    EXPECT_TRUE(!normals.is_dense);
    EXPECT_TRUE(std::isnan(normals[0].normal_x) && std::isnan(normals[0].curvature));
    n.setKSearch(5);  // both set: error, empty output (feature.hpp:135-141)
    n.compute(normals);
    EXPECT_EQ(normals.size(), 0u);
  }

  {  // TYPED_TEST (CorrespondenceEstimationTestSuite, CorrespondenceEstimationNormalShooting) —
     // test/registration/test_correspondence_estimation.cpp:95-137
    auto cloud1 = std::make_shared<PointCloud<PointXYZ>>();
    auto cloud2 = std::make_shared<PointCloud<PointXYZ>>();
    for (std::size_t i = 0; i < 50; ++i)
      for (std::size_t j = 0; j < 25; ++j) {
        cloud1->push_back(PointXYZ(i * 0.2f, 0.f, j * 0.2f));
        cloud2->push_back(PointXYZ(i * 0.2f, 2.f, j * 0.2f));
      }
    NormalEstimation<PointXYZ, Normal> ne;
    ne.setInputCloud(cloud1);
    ne.setSearchMethod(std::make_shared<search::KdTree<PointXYZ>>());
    auto cloud1_normals = std::make_shared<PointCloud<Normal>>();
    ne.setKSearch(5);
    ne.compute(*cloud1_normals);
    auto corr = std::make_shared<Correspondences>();
    registration::CorrespondenceEstimationNormalShooting<PointXYZ, PointXYZ, Normal> ce;
    ce.setInputSource(cloud1);
    ce.setKSearch(10);
    ce.setSourceNormals(cloud1_normals);
    ce.setInputTarget(cloud2);
    ce.determineCorrespondences(*corr);
    EXPECT_EQ(corr->size(), cloud1->size());
    for (std::size_t i = 0; i < corr->size(); i++) EXPECT_EQ((*corr)[i].index_query, (*corr)[i].index_match);
    registration::CorrespondenceEstimationBackProjection<PointXYZ, PointXYZ, Normal> cb;
    cb.setInputSource(cloud1);
    cb.setSourceNormals(cloud1_normals);
    cb.setTargetNormals(cloud1_normals);
    cb.setInputTarget(cloud2);
    corr->clear();
    cb.determineCorrespondences(*corr);
    EXPECT_EQ(corr->size(), cloud1->size());
    for (std::size_t i = 0; i < corr->size(); i++) EXPECT_EQ((*corr)[i].index_query, (*corr)[i].index_match);
    auto cl = ce.clone();
    EXPECT_TRUE(cl->requiresSourceNormals() && !cl->requiresTargetNormals() && cb.requiresTargetNormals());
  }

  {  // TEST (PCL, CorrespondenceRejectorSurfaceNormal) — test_registration_api.cpp:266-317, and
     // TEST (PCL, IterativeClosestPoint_PointToPlane)'s estimator / rejector combination — test_registration.cpp:511-560
    auto src = std::make_shared<PointCloud<PointNormal>>();
    auto tgt = std::make_shared<PointCloud<PointNormal>>();
    for (const auto& p : cloud_source.points) src->push_back(PointNormal(p.x, p.y, p.z));
    for (const auto& p : cloud_target.points) tgt->push_back(PointNormal(p.x, p.y, p.z));
    for (auto* c : {&src, &tgt}) {
      NormalEstimation<PointNormal, PointNormal> norm_est;
      norm_est.setSearchMethod(std::make_shared<search::KdTree<PointNormal>>());
      norm_est.setKSearch(10);
      norm_est.setInputCloud(*c);
      PointCloud<PointNormal> nrm;
      norm_est.compute(nrm);
      EXPECT_EQ(nrm.size(), (*c)->size());
      for (std::size_t i = 0; i < nrm.size(); ++i) {
        (**c)[i].normal_x = nrm[i].normal_x;
        (**c)[i].normal_y = nrm[i].normal_y;
        (**c)[i].normal_z = nrm[i].normal_z;
        (**c)[i].curvature = nrm[i].curvature;
      }
    }
    auto correspondences = std::make_shared<Correspondences>();
    registration::CorrespondenceEstimation<PointNormal, PointNormal> corr_est;
    corr_est.setInputSource(src);
    corr_est.setInputTarget(tgt);
    corr_est.determineCorrespondences(*correspondences);
    EXPECT_EQ(correspondences->size(), 397u);
    registration::CorrespondenceRejectorSurfaceNormal rej;
    rej.initializeDataContainer<PointNormal, PointNormal>();
    rej.setInputSource<PointNormal>(src);
    rej.setInputTarget<PointNormal>(tgt);
    rej.setInputNormals<PointNormal, PointNormal>(src);
    rej.setTargetNormals<PointNormal, PointNormal>(tgt);
    rej.setInputCorrespondences(correspondences);
    rej.setThreshold(0.5);
    Correspondences kept;
    rej.getCorrespondences(kept);
    EXPECT_TRUE(kept.size() > 0 && kept.size() < correspondences->size());
    std::size_t expect = 0;  // the same float dot product on the host
    for (const auto& c : *correspondences) {
      const auto& a = (*src)[c.index_query];
      const auto& b = (*tgt)[c.index_match];
      const float dot = (a.normal_x * b.normal_x) + (a.normal_y * b.normal_y) + (a.normal_z * b.normal_z);
      if (static_cast<double>(dot) > 0.5) {
        if (expect < kept.size()) {
          EXPECT_EQ(kept[expect].index_query, c.index_query);
          EXPECT_EQ(kept[expect].index_match, c.index_match);
        }
        ++expect;
      }
    }
    EXPECT_EQ(kept.size(), expect);

    IterativeClosestPoint<PointNormal, PointNormal> reg;
    reg.setTransformationEstimation(std::make_shared<registration::TransformationEstimationPointToPlaneLLS<PointNormal, PointNormal>>());
    reg.setInputSource(src);
    reg.setInputTarget(tgt);
    reg.setMaximumIterations(50);
    reg.setTransformationEpsilon(1e-8);
    auto ce = std::make_shared<registration::CorrespondenceEstimationNormalShooting<PointNormal, PointNormal, PointNormal>>();
    reg.setCorrespondenceEstimation(ce);
    auto rej2 = std::make_shared<registration::CorrespondenceRejectorSurfaceNormal>();
    rej2->setThreshold(0);
    reg.addCorrespondenceRejector(rej2);
    PointCloud<PointNormal> output;
    reg.align(output);
    EXPECT_EQ(output.size(), cloud_source.size());
    EXPECT_TRUE(reg.hasConverged());
    EXPECT_LT(reg.getFitnessScore(), 0.005);
    for (int iter = 0; iter < 4; iter++) {  // "Check again, for all possible caching schemes" (:543-559)
      const bool force_cache = static_cast<bool>(iter / 2);
      const bool force_cache_reciprocal = static_cast<bool>(iter % 2);
      auto tree = std::make_shared<search::KdTree<PointNormal>>();
      if (force_cache) tree->setInputCloud(tgt);
      reg.setSearchMethodTarget(tree, force_cache);
      auto tree_recip = std::make_shared<search::KdTree<PointNormal>>();
      if (force_cache_reciprocal) tree_recip->setInputCloud(src);
      reg.setSearchMethodSource(tree_recip, force_cache_reciprocal);
      reg.align(output);
      EXPECT_EQ(output.size(), cloud_source.size());
      EXPECT_LT(reg.getFitnessScore(), 0.005);
    }
    auto cbp = std::make_shared<registration::CorrespondenceEstimationBackProjection<PointNormal, PointNormal, PointNormal>>();
    reg.setCorrespondenceEstimation(cbp);
    reg.align(output);
    EXPECT_TRUE(reg.hasConverged());
    EXPECT_LT(reg.getFitnessScore(), 0.005);
  }


  {  // EuclideanClusterExtraction (segmentation/impl/extract_clusters.hpp:225-252): three well separated blobs + stragglers
    auto cloud = std::make_shared<PointCloud<PointXYZ>>();
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.f / 16777216.f); };
    const float centres[3][3] = {{0, 0, 0}, {5, 0, 0}, {0, 5, 0}};
    const int sizes[3] = {300, 200, 100};
    for (int b = 0; b < 3; ++b)
      for (int i = 0; i < sizes[b]; ++i)
        cloud->push_back(PointXYZ(centres[b][0] + 0.3f * rnd(), centres[b][1] + 0.3f * rnd(), centres[b][2] + 0.3f * rnd()));
    for (int i = 0; i < 7; ++i) cloud->push_back(PointXYZ(20.f + 3.f * i, 20.f, 20.f));  // isolated points
    EuclideanClusterExtraction<PointXYZ> ec;
    ec.setClusterTolerance(0.2);
    ec.setMinClusterSize(50);
    ec.setMaxClusterSize(25000);
    ec.setSearchMethod(std::make_shared<search::KdTree<PointXYZ>>());
    ec.setInputCloud(cloud);
    std::vector<PointIndices> clusters;
    ec.extract(clusters);
    EXPECT_EQ(clusters.size(), 3u);
    if (clusters.size() == 3) {
      EXPECT_EQ(clusters[0].indices.size(), 300u);  // largest first
      EXPECT_EQ(clusters[1].indices.size(), 200u);
      EXPECT_EQ(clusters[2].indices.size(), 100u);
      EXPECT_EQ(clusters[0].indices.front(), 0);
      EXPECT_EQ(clusters[0].indices.back(), 299);   // indices ascending inside a cluster
      EXPECT_EQ(clusters[1].indices.front(), 300);
      EXPECT_EQ(clusters[2].indices.back(), 599);
    }
    ec.setMinClusterSize(1);
    ec.extract(clusters);
    EXPECT_EQ(clusters.size(), 10u);  // + the 7 singletons
    auto idx = std::make_shared<Indices>();
    for (int i = 0; i < 300; i += 2) idx->push_back(i);  // every other point of the first blob only
    ec.setIndices(idx);
    ec.setClusterTolerance(0.3);
    ec.extract(clusters);
    EXPECT_EQ(clusters.size(), 1u);
    if (!clusters.empty()) {
      EXPECT_EQ(clusters[0].indices.size(), 150u);
      EXPECT_EQ(clusters[0].indices[1], 2);
    }
  }

  std::printf("%s: %d checks, %d failures\n", g_fail ? "FAILED" : "PASSED", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
