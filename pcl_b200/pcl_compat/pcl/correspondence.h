// pcl/correspondence.h — common/include/pcl/correspondence.h:60-91 (12-byte POD) == pclb200_corr
#pragma once
#include <memory>
#include <vector>

#include "types.h"
#include "../../../include/pclb200.h"
namespace pcl {
struct Correspondence {
  index_t index_query = 0;
  index_t index_match = -1;
  union { float distance; float weight; };
  Correspondence() : distance(0.f) {}
  Correspondence(index_t q, index_t m, float d) : index_query(q), index_match(m), distance(d) {}
};
static_assert(sizeof(Correspondence) == sizeof(pclb200_corr), "Correspondence must match the C-ABI record");
using Correspondences = std::vector<Correspondence>;
using CorrespondencesPtr = std::shared_ptr<Correspondences>;
using CorrespondencesConstPtr = std::shared_ptr<const Correspondences>;
}  // namespace pcl
