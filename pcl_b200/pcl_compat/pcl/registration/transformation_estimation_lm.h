// pcl/registration/transformation_estimation_lm.h — TransformationEstimationLM<PointSource, PointTarget, Scalar>
// (registration/include/pcl/registration/transformation_estimation_lm.h:55-330, impl/transformation_estimation_lm.hpp).
// The reference minimises sum |T p_i - q_i|^2 over the six rigid parameters with Levenberg-Marquardt, starting from the
// identity.  That objective has a closed-form minimiser — the one TransformationEstimationSVD computes — and that is what
// this class returns, through the same device estimator: the LM ITERATIONS are not reproduced (they would run on the host,
// one Jacobian per pair per step), the value they converge to is.  A custom warp function (setWarpFunction) other than the
// rigid one is therefore not supported and is refused loudly.
#pragma once
#include <cstdio>

#include "transformation_estimation.h"

namespace pcl {
namespace registration {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationEstimationLM : public TransformationEstimationSVD<PointSource, PointTarget, Scalar> {
public:
  using Ptr = std::shared_ptr<TransformationEstimationLM<PointSource, PointTarget, Scalar>>;
  using ConstPtr = std::shared_ptr<const TransformationEstimationLM<PointSource, PointTarget, Scalar>>;
  TransformationEstimationLM() : TransformationEstimationSVD<PointSource, PointTarget, Scalar>(true) {}
  // transformation_estimation_lm.h:140-146: only the default (rigid, 6 parameters) warp has a closed form
  template <typename WarpFunctionPtr>
  void setWarpFunction(const WarpFunctionPtr&)
  {
    std::fprintf(stderr, "[pcl::registration::TransformationEstimationLM::setWarpFunction] only the rigid warp is supported: the estimate is the "
                         "closed-form minimiser, not an LM iteration\n");
  }
};
}  // namespace registration
}  // namespace pcl
