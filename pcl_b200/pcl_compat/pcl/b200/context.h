// pcl/b200/context.h — process-wide handle on libpclb200 for the facade classes.
// PCL's classes take no device argument, so the facade keeps one lazily created context per process
// (device 0, or $PCLB200_DEVICE).  There is no CPU fallback: without a CUDA device construction throws.
#pragma once
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>

#include "../../../../include/pclb200.h"

namespace pcl {
namespace b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc, const char* where)
{
  if (rc != PCLB200_OK)
    throw Error(rc, std::string(where) + ": " + pclb200_last_error());
}

class Context {
public:
  static pclb200_ctx* get()
  {
    static Context c;
    return c.h_;
  }
  Context(const Context&) = delete;

private:
  Context()
  {
    const char* d = std::getenv("PCLB200_DEVICE");
    check(pclb200_create(d ? std::atoi(d) : 0, &h_), "pclb200_create");
  }
  ~Context() { pclb200_destroy(h_); }
  pclb200_ctx* h_ = nullptr;
};

// shared ownership of a device index (the LBVH) so trees can be handed between objects like PCL's KdTreePtr
struct IndexHandle {
  pclb200_index* h = nullptr;
  explicit IndexHandle(pclb200_index* p) : h(p) {}
  ~IndexHandle() { pclb200_index_destroy(h); }
  IndexHandle(const IndexHandle&) = delete;
};

}  // namespace b200
}  // namespace pcl
