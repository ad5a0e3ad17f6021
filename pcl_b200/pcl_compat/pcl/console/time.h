// pcl/console/time.h — pcl::console::TicToc (common/include/pcl/console/time.h:48-86): wall-clock stopwatch of the tools
#pragma once
#include <chrono>

#include "print.h"

namespace pcl {
namespace console {
class TicToc {
public:
  void tic() { tictic_ = std::chrono::steady_clock::now(); }
  // milliseconds since the last tic()
  double toc() const { return std::chrono::duration<double, std::ratio<1, 1000>>(std::chrono::steady_clock::now() - tictic_).count(); }
  void toc_print() const
  {
    const double milliseconds = toc();
    print_value("%g", milliseconds);
    print_info(" ms\n");
  }

private:
  std::chrono::time_point<std::chrono::steady_clock> tictic_ = std::chrono::steady_clock::now();
};
}  // namespace console
}  // namespace pcl
