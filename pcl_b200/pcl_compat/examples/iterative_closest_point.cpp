// A command-line registration in the shape of the reference's tools/iterative_closest_point.cpp:87-226 — load two PCD files
// as type-erased clouds, convert to PointNormal, run pcl::IterativeClosestPoint<PointNormal, PointNormal, double> with an
// injected correspondence estimator, transformation estimator and a one-to-one rejector, put the aligned coordinates back
// beside the source's remaining fields and write the result — compiled against the facade, so every stage runs on the
// device through libpclb200.  (The reference tool's Levenberg–Marquardt estimator is outside this path; the SVD
// estimator it lists as the alternative is used.)
//   iterative_closest_point source.pcd target.pcd output.pcd [max_iterations] [max_correspondence_distance]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <pcl/common/io.h>
#include <pcl/conversions.h>
#include <pcl/io/pcd_io.h>
#include <pcl/point_types.h>
#include <pcl/registration/correspondence_estimation.h>
#include <pcl/registration/correspondence_rejection_one_to_one.h>
#include <pcl/registration/icp.h>
#include <pcl/registration/transformation_estimation_svd.h>

using namespace pcl;
using namespace pcl::registration;

static Eigen::Vector4f translation;
static Eigen::Quaternionf orientation;

static bool loadCloud(const std::string& filename, PCLPointCloud2& cloud)
{
  const auto t0 = std::chrono::steady_clock::now();
  if (io::loadPCDFile(filename, cloud, translation, orientation) < 0) return false;
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::printf("Loading %s [done, %g ms : %u points]\nAvailable dimensions: %s\n", filename.c_str(), ms, cloud.width * cloud.height,
              getFieldsList(cloud).c_str());
  return true;
}

static bool compute(const PCLPointCloud2::ConstPtr& source, const PCLPointCloud2::ConstPtr& target, PCLPointCloud2& transformed_source,
                    int max_iterations, double max_distance)
{
  PointCloud<PointNormal>::Ptr src(new PointCloud<PointNormal>), tgt(new PointCloud<PointNormal>);
  fromPCLPointCloud2(*source, *src);
  fromPCLPointCloud2(*target, *tgt);
  const auto t0 = std::chrono::steady_clock::now();
  using Scalar = double;
  TransformationEstimationSVD<PointNormal, PointNormal, Scalar>::Ptr te(new TransformationEstimationSVD<PointNormal, PointNormal, Scalar>);
  CorrespondenceEstimation<PointNormal, PointNormal, Scalar>::Ptr cens(new CorrespondenceEstimation<PointNormal, PointNormal, Scalar>);
  cens->setInputSource(src);
  cens->setInputTarget(tgt);
  CorrespondenceRejectorOneToOne::Ptr cor_rej_o2o(new CorrespondenceRejectorOneToOne);
  IterativeClosestPoint<PointNormal, PointNormal, Scalar> icp;
  icp.setCorrespondenceEstimation(cens);
  icp.setTransformationEstimation(te);
  icp.addCorrespondenceRejector(cor_rej_o2o);
  icp.setInputSource(src);
  icp.setInputTarget(tgt);
  icp.setMaximumIterations(max_iterations);
  icp.setTransformationEpsilon(1e-10);
  if (max_distance > 0) icp.setMaxCorrespondenceDistance(max_distance);
  PointCloud<PointNormal> output;
  icp.align(output);
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::printf("Computing [done, %g ms : %u points], has converged: %d after %d iterations with score: %f\n", ms, output.width * output.height,
              icp.hasConverged() ? 1 : 0, icp.getNumberOfIterations(), icp.getFitnessScore());
  const Eigen::Matrix4d T = icp.getFinalTransformation();
  std::printf("Transformation is:\n");
  for (int r = 0; r < 4; ++r) std::printf("\t%9.6f\t%9.6f\t%9.6f\t%9.6f\n", T(r, 0), T(r, 1), T(r, 2), T(r, 3));
  PCLPointCloud2 output_source;
  toPCLPointCloud2(output, output_source);
  return concatenateFields(*source, output_source, transformed_source) && icp.hasConverged();
}

int main(int argc, char** argv)
{
  std::printf("Estimate a rigid transformation using IterativeClosestPoint.\n");
  if (argc < 4) {
    std::fprintf(stderr, "Syntax is: %s input_source.pcd input_target.pcd output.pcd [max_iterations] [max_correspondence_distance]\n", argv[0]);
    return -1;
  }
  const int max_iterations = argc > 4 ? std::atoi(argv[4]) : 1000;
  const double max_distance = argc > 5 ? std::atof(argv[5]) : 0.0;
  PCLPointCloud2::Ptr src(new PCLPointCloud2), tgt(new PCLPointCloud2);
  if (!loadCloud(argv[1], *src) || !loadCloud(argv[2], *tgt)) return -1;
  PCLPointCloud2 output;
  const bool ok = compute(src, tgt, output, max_iterations, max_distance);
  PCDWriter w;
  if (w.writeASCII(argv[3], output, translation, orientation) != 0) return -1;
  std::printf("Saving %s [done : %u points, fields: %s]\n", argv[3], output.width * output.height, getFieldsList(output).c_str());
  return ok ? 0 : 1;
}
