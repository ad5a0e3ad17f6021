// pcl/types.h — index types (common/include/pcl/types.h:97-133)
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
using index_t = std::int32_t;
using uindex_t = std::uint32_t;
using Indices = std::vector<index_t>;
using IndicesPtr = std::shared_ptr<Indices>;
using IndicesConstPtr = std::shared_ptr<const Indices>;
}  // namespace pcl
