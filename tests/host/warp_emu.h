// warp_emu.h — a lock-step emulation of ONE CUDA thread block (up to 8 warps) for host-compiled device code: one fiber
// (ucontext) per thread, run round-robin; every warp-synchronous primitive (__shfl*_sync, __ballot_sync, __match_any_sync, __syncwarp) is a
// rendezvous of all 32 lanes: each lane deposits its operand, yields until the last lane has arrived, then reads what it
// needs.  Deterministic and single-threaded.  The emulated code must execute the same sequence of primitives on every
// lane (warp-uniform control flow around them), which is what the *_sync primitives with a full mask require on the
// device as well.  Included through PCLB_HOST_EXTRA_SHIMS, before the device headers.  Test infrastructure.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <vector>

namespace warp_emu {
constexpr int kLanes = 32;
constexpr int kMaxThreads = 256;
struct Block {
  ucontext_t main_ctx, ctx[kMaxThreads];
  std::vector<char> stacks[kMaxThreads];
  bool finished[kMaxThreads];
  int nthreads = 32;
  int current = -1;
  // one rendezvous state per warp, one for the whole block
  unsigned warp_generation[kMaxThreads / kLanes] = {0};
  int warp_arrived[kMaxThreads / kLanes] = {0};
  unsigned block_generation = 0;
  int block_arrived = 0;
  std::uint64_t slot[kMaxThreads];    // operands of the warp primitive in flight (per thread)
  std::uint64_t slot2[kMaxThreads];
  long primitives = 0;
  std::function<void()> body;
};
inline Block& W() { static Block w; return w; }

// The order in which the threads of a block get to run between two synchronisation points is a free choice of the hardware; the
// emulation makes one: PCLB_EMU_ORDER unset = ascending thread index, "reverse" = descending, a number = a stride co-prime with the
// block size (a fixed pseudo-random permutation).  Code that is correct only under one order — a missing __syncthreads /
// __syncwarp between a write and another thread's read — gives different results under the others (tools/dev/racecheck_host_kernels.sh).
inline int schedule_stride(int nthreads)
{
  static const char* e = std::getenv("PCLB_EMU_ORDER");
  if (!e || !*e) return 1;
  if (e[0] == 'r') return nthreads - 1;                       // me - 1 (mod n)
  long k = std::strtol(e, nullptr, 10);
  int s = static_cast<int>((k * 2 + 1) % nthreads);            // odd: co-prime with the power-of-two-multiple block sizes used here
  while (s > 1 && std::__gcd(s, nthreads) != 1) s -= 2;
  return s < 1 ? 1 : s;
}
inline void yield_lane()
{
  Block& w = W();
  const int me = w.current;
  const int stride = schedule_stride(w.nthreads);
  for (int step = 1; step <= w.nthreads; ++step) {
    const int next = static_cast<int>((me + static_cast<long>(step) * stride) % w.nthreads);
    if (next == me) return;
    if (!w.finished[next]) {
      w.current = next;
      swapcontext(&w.ctx[me], &w.ctx[next]);
      return;
    }
  }
}
// the 32 lanes of the calling thread's warp meet here
inline void rendezvous()
{
  Block& w = W();
  const int wid = w.current / kLanes;
  const unsigned gen = w.warp_generation[wid];
  if (++w.warp_arrived[wid] == kLanes) {
    w.warp_arrived[wid] = 0;
    ++w.warp_generation[wid];
    ++w.primitives;
  }
  while (w.warp_generation[wid] == gen) yield_lane();
}
// all threads of the block meet here (__syncthreads)
inline void block_rendezvous()
{
  Block& w = W();
  const unsigned gen = w.block_generation;
  if (++w.block_arrived == w.nthreads) {
    w.block_arrived = 0;
    ++w.block_generation;
  }
  while (w.block_generation == gen) yield_lane();
}
inline void trampoline()
{
  Block& w = W();
  const int me = w.current;
  w.body();
  w.finished[me] = true;
  const int stride = schedule_stride(w.nthreads);
  for (int step = 1; step < w.nthreads; ++step) {   // hand over to a thread that still runs, or back to the caller
    const int next = static_cast<int>((me + static_cast<long>(step) * stride) % w.nthreads);
    if (!w.finished[next]) {
      w.current = next;
      setcontext(&w.ctx[next]);
    }
  }
  setcontext(&w.main_ctx);
}
}  // namespace warp_emu

// ---- what the device code sees -----------------------------------------------------------------------------------------
struct EmuDim { unsigned x = 1, y = 1, z = 1; };
struct EmuThreadIdx { operator int() const = delete; };
namespace warp_emu { inline unsigned lane_id() { return static_cast<unsigned>(W().current); } }   // = threadIdx.x
struct EmuThreadIdxX { operator unsigned() const { return warp_emu::lane_id(); } };
struct EmuThreadIdxT { EmuThreadIdxX x; };
static EmuThreadIdxT threadIdx;
static EmuDim blockIdx_storage{0, 0, 0};
#define blockIdx blockIdx_storage
static EmuDim blockDim{32, 1, 1}, gridDim{1, 1, 1};   // a test sets blockDim.x to the size it passes to run_block

#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __shared__
#define __shared__ static

template <typename T> static inline std::uint64_t emu_pack(T v) { std::uint64_t u = 0; std::memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T emu_unpack(std::uint64_t u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }
template <typename T> static inline T emu_exchange(T v, int src_lane)
{
  auto& w = warp_emu::W();
  const int base = w.current & ~31;
  w.slot[w.current] = emu_pack(v);
  warp_emu::rendezvous();
  const T r = emu_unpack<T>(w.slot[base + (src_lane & 31)]);
  warp_emu::rendezvous();   // nobody overwrites a slot before every lane has read
  return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (warp_emu::W().current & 31) ^ m); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
  const int me = warp_emu::W().current & 31;
  return emu_exchange(v, me >= static_cast<int>(d) ? me - static_cast<int>(d) : me);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
  const int me = warp_emu::W().current & 31;
  return emu_exchange(v, me + static_cast<int>(d) < 32 ? me + static_cast<int>(d) : me);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
  auto& w = warp_emu::W();
  const int base = w.current & ~31;
  w.slot[w.current] = pred ? 1u : 0u;
  warp_emu::rendezvous();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= static_cast<unsigned>(w.slot[base + l] & 1u) << l;
  warp_emu::rendezvous();
  return m;
}
template <typename T> static inline unsigned __match_any_sync(unsigned, T v)
{
  auto& w = warp_emu::W();
  const int base = w.current & ~31;
  w.slot[w.current] = emu_pack(v);
  warp_emu::rendezvous();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (w.slot[base + l] == w.slot[w.current] ? 1u : 0u) << l;
  warp_emu::rendezvous();
  return m;
}
// mma.sync.aligned.m8n8k4.row.col.f64: lane l feeds A[l / 4][l % 4] and B[l % 4][l / 4] and owns C[l / 4][2 (l % 4)], [.. + 1]
static inline void emu_dmma884(double& c0, double& c1, double a, double b)
{
  auto& w = warp_emu::W();
  const int base = w.current & ~31, lane = w.current & 31;
  w.slot[w.current] = emu_pack(a);
  w.slot2[w.current] = emu_pack(b);
  warp_emu::rendezvous();
  const int g = lane >> 2, t = lane & 3;
  for (int k = 0; k < 4; ++k) {
    const double A = emu_unpack<double>(w.slot[base + g * 4 + k]);
    c0 = std::fma(A, emu_unpack<double>(w.slot2[base + (2 * t) * 4 + k]), c0);
    c1 = std::fma(A, emu_unpack<double>(w.slot2[base + (2 * t + 1) * 4 + k]), c1);
  }
  warp_emu::rendezvous();
}
static inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::rendezvous(); }
static inline void __syncthreads() { warp_emu::block_rendezvous(); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline long long clock64() { return 0; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }          // fibers never run concurrently
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
static inline int atomicMin(int* p, int v) { const int o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
static inline int atomicCAS(int* p, int cmp, int v) { const int o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { const unsigned long long o = *p; if (o == cmp) *p = v; return o; }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll(static_cast<unsigned long long>(x)); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __fdiv_rn(float a, float b) { volatile float x = a, y = b; volatile float r = x / y; return r; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
using std::isfinite;

namespace warp_emu {
// runs `kernel_body` once on each thread of an emulated block of `nthreads` (a multiple of 32) threads, in lock step
inline long run_block(int nthreads, std::function<void()> kernel_body)
{
  Block& w = W();
  if (nthreads % kLanes != 0 || nthreads > kMaxThreads) { std::printf("warp_emu: unsupported block size %d\n", nthreads); std::exit(3); }
  w.body = std::move(kernel_body);
  w.nthreads = nthreads;
  w.block_generation = 0;
  w.block_arrived = 0;
  w.primitives = 0;
  for (int i = 0; i < kMaxThreads / kLanes; ++i) { w.warp_generation[i] = 0; w.warp_arrived[i] = 0; }
  for (int l = 0; l < nthreads; ++l) {
    w.finished[l] = false;
    if (w.stacks[l].empty()) w.stacks[l].resize(512 << 10);
    getcontext(&w.ctx[l]);
    w.ctx[l].uc_stack.ss_sp = w.stacks[l].data();
    w.ctx[l].uc_stack.ss_size = w.stacks[l].size();
    w.ctx[l].uc_link = nullptr;
    makecontext(&w.ctx[l], reinterpret_cast<void (*)()>(trampoline), 0);
  }
  w.current = schedule_stride(nthreads) == 1 ? 0 : nthreads - 1;   // another first thread under the other orders
  swapcontext(&w.main_ctx, &w.ctx[w.current]);
  for (int l = 0; l < nthreads; ++l)
    if (!w.finished[l]) { std::printf("warp_emu: thread %d did not finish (divergent primitive sequence?)\n", l); std::exit(3); }
  return w.primitives;
}
inline long run_warp(std::function<void()> kernel_body) { return run_block(kLanes, std::move(kernel_body)); }
}  // namespace warp_emu
