// pcl/memory.h — the smart-pointer spellings PCL code uses (common/include/pcl/memory.h:50-140): aliases of the std ones
#pragma once
#include <memory>
#include <utility>

namespace pcl {
using std::dynamic_pointer_cast;
using std::shared_ptr;
using std::static_pointer_cast;
using std::weak_ptr;
template <typename T, typename... Args>
shared_ptr<T> make_shared(Args&&... args)
{
  return std::make_shared<T>(std::forward<Args>(args)...);
}
}  // namespace pcl
