// pcl/b200/context.h — process-wide handle on libpclb200 for the facade classes.
// PCL's classes take no device argument, so the facade keeps one lazily created context per process
// (device 0, or $PCLB200_DEVICE); a thread can open a context of its own with Context::ThreadScope.  There is no CPU fallback: without a CUDA device construction throws.
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
    if (pclb200_ctx* own = threadSlot())
      return own;
    static Context c;
    return c.h_;
  }
  Context(const Context&) = delete;

  // While a ThreadScope lives, the facade calls made by THIS thread run on a context of their own (its own CUDA stream
  // and allocator cache) instead of the process-wide one.  A context serialises its calls; two threads that each hold a
  // scope overlap on the device — one align()'s PCIe copies under the other's kernels (bench.py's e2e leg: 39.8 -> 30.3
  // ms per 10 M-point align).  Device objects created inside a scope (trees, registration objects) must be destroyed
  // before it ends; objects of the process-wide context (a shared target tree) may be used inside any scope.
  class ThreadScope {
  public:
    ThreadScope() : prev_(threadSlot())
    {
      check(pclb200_create(device(), &own_), "pclb200_create");
      threadSlot() = own_;
    }
    ~ThreadScope()
    {
      threadSlot() = prev_;
      pclb200_destroy(own_);
    }
    ThreadScope(const ThreadScope&) = delete;

  private:
    pclb200_ctx* prev_;
    pclb200_ctx* own_ = nullptr;
  };

private:
  static pclb200_ctx*& threadSlot()
  {
    static thread_local pclb200_ctx* p = nullptr;
    return p;
  }
  static int device()
  {
    const char* d = std::getenv("PCLB200_DEVICE");
    return d ? std::atoi(d) : 0;
  }
  Context() { check(pclb200_create(device(), &h_), "pclb200_create"); }
  ~Context() { pclb200_destroy(h_); }
  pclb200_ctx* h_ = nullptr;
};

// Page-locks the storage of a cloud for as long as the object lives (the pinned reader — load the file,
// pin once, and every setInputSource / setInputCloud / align() on that cloud moves it by DMA at PCIe rate instead of
// through the pageable staging path; at 200 M points that is the difference between ~0.25 s and ~0.06 s per upload).
// The cloud must not be resized or destroyed while pinned.
template <typename CloudT>
class PinnedCloud {
public:
  explicit PinnedCloud(CloudT& cloud) : p_(cloud.points.empty() ? nullptr : cloud.points.data())
  {
    if (p_) check(pclb200_host_register(Context::get(), p_, cloud.points.size() * sizeof(cloud.points[0])), "pclb200_host_register");
  }
  ~PinnedCloud()
  {
    if (p_) pclb200_host_unregister(Context::get(), p_);
  }
  PinnedCloud(const PinnedCloud&) = delete;
  PinnedCloud& operator=(const PinnedCloud&) = delete;

private:
  void* p_;
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
