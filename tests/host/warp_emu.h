// warp_emu.h — a lock-step emulation of ONE CUDA warp for host-compiled device code: 32 fibers (ucontext), one per lane,
// run round-robin; every warp-synchronous primitive (__shfl*_sync, __ballot_sync, __match_any_sync, __syncwarp) is a
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
#include <functional>
#include <vector>

namespace warp_emu {
constexpr int kLanes = 32;
struct Warp {
  ucontext_t main_ctx, lane_ctx[kLanes];
  std::vector<char> stacks[kLanes];
  bool finished[kLanes];
  int current = -1;
  unsigned generation = 0;
  int arrived = 0;
  std::uint64_t slot[kLanes];   // operands of the primitive in flight
  long primitives = 0;
  std::function<void()> body;
};
inline Warp& W() { static Warp w; return w; }

inline void yield_lane()
{
  Warp& w = W();
  const int me = w.current;
  for (int step = 1; step <= kLanes; ++step) {
    const int next = (me + step) % kLanes;
    if (!w.finished[next] || next == me) {
      if (next == me) return;
      w.current = next;
      swapcontext(&w.lane_ctx[me], &w.lane_ctx[next]);
      return;
    }
  }
}
// all 32 lanes meet here
inline void rendezvous()
{
  Warp& w = W();
  const unsigned gen = w.generation;
  if (++w.arrived == kLanes) {
    w.arrived = 0;
    ++w.generation;
    ++w.primitives;
  }
  while (w.generation == gen) yield_lane();
}
inline void trampoline()
{
  Warp& w = W();
  const int me = w.current;
  w.body();
  w.finished[me] = true;
  // hand over to a lane that still runs, or back to the caller
  for (int step = 1; step < kLanes; ++step) {
    const int next = (me + step) % kLanes;
    if (!w.finished[next]) {
      w.current = next;
      setcontext(&w.lane_ctx[next]);
    }
  }
  setcontext(&w.main_ctx);
}
}  // namespace warp_emu

// ---- what the device code sees -----------------------------------------------------------------------------------------
struct EmuDim { unsigned x = 1, y = 1, z = 1; };
struct EmuThreadIdx { operator int() const = delete; };
namespace warp_emu { inline unsigned lane_id() { return static_cast<unsigned>(W().current); } }
struct EmuThreadIdxX { operator unsigned() const { return warp_emu::lane_id(); } };
struct EmuThreadIdxT { EmuThreadIdxX x; };
static EmuThreadIdxT threadIdx;
static EmuDim blockIdx_storage{0, 0, 0};
#define blockIdx blockIdx_storage
static EmuDim blockDim{32, 1, 1}, gridDim{1, 1, 1};

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
  w.slot[w.current] = emu_pack(v);
  warp_emu::rendezvous();
  const T r = emu_unpack<T>(w.slot[src_lane & 31]);
  warp_emu::rendezvous();   // nobody overwrites a slot before every lane has read
  return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, warp_emu::W().current ^ m); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
  const int me = warp_emu::W().current;
  return emu_exchange(v, me >= static_cast<int>(d) ? me - static_cast<int>(d) : me);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
  const int me = warp_emu::W().current;
  return emu_exchange(v, me + static_cast<int>(d) < 32 ? me + static_cast<int>(d) : me);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
  auto& w = warp_emu::W();
  w.slot[w.current] = pred ? 1u : 0u;
  warp_emu::rendezvous();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= static_cast<unsigned>(w.slot[l] & 1u) << l;
  warp_emu::rendezvous();
  return m;
}
template <typename T> static inline unsigned __match_any_sync(unsigned, T v)
{
  auto& w = warp_emu::W();
  w.slot[w.current] = emu_pack(v);
  warp_emu::rendezvous();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (w.slot[l] == w.slot[w.current] ? 1u : 0u) << l;
  warp_emu::rendezvous();
  return m;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::rendezvous(); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __fdiv_rn(float a, float b) { volatile float x = a, y = b; volatile float r = x / y; return r; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
using std::isfinite;

namespace warp_emu {
// runs `kernel_body` once on each of the 32 lanes of the emulated warp, in lock step
inline long run_warp(std::function<void()> kernel_body)
{
  Warp& w = W();
  w.body = std::move(kernel_body);
  w.generation = 0;
  w.arrived = 0;
  w.primitives = 0;
  for (int l = 0; l < kLanes; ++l) {
    w.finished[l] = false;
    if (w.stacks[l].empty()) w.stacks[l].resize(1 << 20);
    getcontext(&w.lane_ctx[l]);
    w.lane_ctx[l].uc_stack.ss_sp = w.stacks[l].data();
    w.lane_ctx[l].uc_stack.ss_size = w.stacks[l].size();
    w.lane_ctx[l].uc_link = nullptr;
    makecontext(&w.lane_ctx[l], reinterpret_cast<void (*)()>(trampoline), 0);
  }
  w.current = 0;
  swapcontext(&w.main_ctx, &w.lane_ctx[0]);
  for (int l = 0; l < kLanes; ++l)
    if (!w.finished[l]) { std::printf("warp_emu: lane %d did not finish (divergent primitive sequence?)\n", l); std::exit(3); }
  return w.primitives;
}
}  // namespace warp_emu
