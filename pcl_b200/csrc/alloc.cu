// alloc.cu — per-stream block cache behind DevBuf (see internal.cuh).
#include <map>
#include <mutex>
#include <unordered_map>

#include "internal.cuh"

namespace pclb200 {

namespace {
struct BlockCache {
  std::multimap<size_t, void*> free_blocks;       // size -> block
  std::unordered_map<void*, size_t> size_of;      // every block this cache owns (free or in use)
};
std::mutex g_mu;
std::unordered_map<cudaStream_t, BlockCache> g_caches;

inline size_t round_up(size_t b)
{
  const size_t g = b < (1u << 20) ? 512 : (1u << 16);  // 512 B granules for small, 64 KiB for large requests
  return (b + g - 1) / g * g;
}
}  // namespace

void* cached_alloc(cudaStream_t s, size_t bytes)
{
  const size_t want = round_up(bytes);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    BlockCache& c = g_caches[s];
    auto it = c.free_blocks.lower_bound(want);
    // accept a cached block unless it would waste more than half of itself (and more than 1 MiB)
    if (it != c.free_blocks.end() && (it->first <= 2 * want || it->first - want <= (1u << 20))) {
      void* p = it->second;
      c.free_blocks.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    // out of memory: give cached blocks back to the driver and retry once
    cudaGetLastError();
    cudaStreamSynchronize(s);
    {
      std::lock_guard<std::mutex> lk(g_mu);
      BlockCache& c = g_caches[s];
      for (auto& kv : c.free_blocks) {
        cudaFree(kv.second);
        c.size_of.erase(kv.second);
      }
      c.free_blocks.clear();
    }
    e = cudaMalloc(&p, want);
    if (e != cudaSuccess)
      throw Error(PCLB200_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(want) + " bytes): " + cudaGetErrorString(e));
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_caches[s].size_of[p] = want;
  return p;
}

void cached_free(cudaStream_t s, void* p)
{
  if (!p)
    return;
  std::lock_guard<std::mutex> lk(g_mu);
  auto ci = g_caches.find(s);
  if (ci == g_caches.end())
    return;  // cache already torn down with its context: the block went with it
  auto it = ci->second.size_of.find(p);
  if (it == ci->second.size_of.end())
    return;
  ci->second.free_blocks.emplace(it->second, p);
}

void cached_release_all(cudaStream_t s)
{
  std::lock_guard<std::mutex> lk(g_mu);
  auto ci = g_caches.find(s);
  if (ci == g_caches.end())
    return;
  for (auto& kv : ci->second.size_of)
    cudaFree(kv.first);
  g_caches.erase(ci);
}

}  // namespace pclb200
