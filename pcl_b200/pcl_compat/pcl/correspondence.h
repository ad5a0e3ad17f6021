// pcl/correspondence.h — common/include/pcl/correspondence.h:60-91 (12-byte POD) == pclb200_corr
#pragma once
#include <algorithm>
#include <iterator>
#include <memory>
#include <ostream>
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

inline std::ostream& operator<<(std::ostream& os, const Correspondence& c)   // common/src/correspondence.cpp:88-94
{
  os << c.index_query << " " << c.index_match << " " << c.distance;
  return os;
}

// Query indices that a rejection step removed (common/src/correspondence.cpp:46-84): the sorted set difference of the
// index_query values before and after; pass presorting_required = false when both lists are already ordered by query
inline void getRejectedQueryIndices(const pcl::Correspondences& correspondences_before,
                                    const pcl::Correspondences& correspondences_after, Indices& indices,
                                    bool presorting_required = true)
{
  indices.clear();
  if (correspondences_before.empty()) return;
  Indices before(correspondences_before.size()), after(correspondences_after.size());
  for (std::size_t i = 0; i < before.size(); ++i) before[i] = correspondences_before[i].index_query;
  for (std::size_t i = 0; i < after.size(); ++i) after[i] = correspondences_after[i].index_query;
  if (after.empty()) {
    indices = before;
    return;
  }
  if (presorting_required) {
    std::sort(before.begin(), before.end());
    std::sort(after.begin(), after.end());
  }
  std::set_difference(before.begin(), before.end(), after.begin(), after.end(), std::back_inserter(indices));
}

// ordering used when the strongest correspondences are wanted first (correspondence.h:140-144)
inline bool isBetterCorrespondence(const Correspondence& pc1, const Correspondence& pc2) { return pc1.distance > pc2.distance; }
}  // namespace pcl
