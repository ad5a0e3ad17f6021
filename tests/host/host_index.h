// host_index.h — shared by the host-compiled device-code tests (tests/host/*.cpp): the CUDA intrinsics the device
// headers use, supplied for g++ with the same rounding, and an index built on the host to lbvh.cu's invariants
// (Morton-ordered padded leaves, 64-byte nodes with both children's boxes, per-node leaf ranges, the (level, cell) ->
// subtree hash table), plus brute force under the library's own distance expression and tie rule.  Test infrastructure.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

// ---- the CUDA intrinsics traverse.cuh uses, for the host ---------------------------------------------------------
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { volatile float x = a, y = b; volatile float r = x + y; return r; }
static inline float __fsub_rn(float a, float b) { volatile float x = a, y = b; volatile float r = x - y; return r; }
static inline float __fmul_rn(float a, float b) { volatile float x = a, y = b; volatile float r = x * y; return r; }
template <typename F> static inline float directed(int mode, F f)
{
  std::fesetround(mode);
  volatile float r = f();
  std::fesetround(FE_TONEAREST);
  return r;
}
static inline float __fadd_rd(float a, float b) { volatile float x = a, y = b; return directed(FE_DOWNWARD, [&] { return x + y; }); }
static inline float __fadd_ru(float a, float b) { volatile float x = a, y = b; return directed(FE_UPWARD, [&] { return x + y; }); }
static inline float __fsub_rd(float a, float b) { volatile float x = a, y = b; return directed(FE_DOWNWARD, [&] { return x - y; }); }
static inline float __fmul_rd(float a, float b) { volatile float x = a, y = b; return directed(FE_DOWNWARD, [&] { return x * y; }); }
static inline float __fmul_ru(float a, float b) { volatile float x = a, y = b; return directed(FE_UPWARD, [&] { return x * y; }); }
static inline float __fsqrt_rn(float a) { volatile float x = a; volatile float r = std::sqrt(x); return r; }
static inline float __fsqrt_ru(float a) { volatile float x = a; return directed(FE_UPWARD, [&] { return std::sqrt(x); }); }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }   // CUDA's global integer max
static inline double __dadd_rn(double a, double b) { volatile double x = a, y = b; volatile double r = x + y; return r; }
static inline double __dsub_rn(double a, double b) { volatile double x = a, y = b; volatile double r = x - y; return r; }
static inline double __dmul_rn(double a, double b) { volatile double x = a, y = b; volatile double r = x * y; return r; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz(static_cast<unsigned>(x)); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }

#ifdef PCLB_HOST_EXTRA_SHIMS
#include PCLB_HOST_EXTRA_SHIMS
#endif
#include "../../pcl_b200/csrc/traverse.cuh"

using namespace pclb200;

// ---- host-side index, to lbvh.cu's invariants --------------------------------------------------------------------
struct HostIndex {
  std::vector<float4> pts;       // leaves * kLeafSize, padded with +inf / sentinel
  std::vector<BvhNode> nodes;
  std::vector<uint2> slots;
  std::vector<int2> node_leaves;  // per internal node: {first leaf, number of leaves} — a subtree's leaves are consecutive
  int root = kDone;
  float lo[3], hi[3], scale = 1.f, margin = 0.f;
  int bmax = 0;
  unsigned log2_slots = 0;
  std::vector<float> xyz;        // original points (3 per point), for brute force
  TreeView view(bool with_table) const
  {
    TreeView t;
    t.nodes = nodes.data();
    t.pts = pts.data();
    t.root = root;
    if (with_table && log2_slots) {
      t.cells.slots = slots.data();
      t.cells.shift = 32u - log2_slots;
      t.cells.mask = (1u << log2_slots) - 1u;
      t.cells.bmax = bmax;
      for (int d = 0; d < 3; ++d) t.cells.lo[d] = lo[d];
      t.cells.scale = scale;
      t.cells.inv_scale = 1.f / scale;
      t.cells.margin = margin;
    }
    return t;
  }
};

static int host_prefix_len63(unsigned long long a, unsigned long long b)
{
  const unsigned long long x = a ^ b;
  return x == 0 ? 63 : __builtin_clzll(x) - 1;
}
static unsigned host_compact21(unsigned long long v)   // every third bit of v
{
  unsigned r = 0;
  for (int i = 0; i < 21; ++i) r |= static_cast<unsigned>((v >> (3 * i)) & 1ull) << i;
  return r;
}

struct Builder {
  HostIndex& I;
  std::vector<unsigned long long> keys;   // sorted
  std::vector<int> order;                 // sorted position -> original index
  std::vector<int> leaf_of_first;         // leaf id by first sorted position
  struct Box { float lo[3], hi[3]; };
  std::vector<std::pair<unsigned, int>> entries;  // (cell key, reference)

  void insert(int b, unsigned long long key, int ref)
  {
    const unsigned cx = host_compact21(key) >> (21 - b), cy = host_compact21(key >> 1) >> (21 - b), cz = host_compact21(key >> 2) >> (21 - b);
    entries.emplace_back(cell_key(b, cx, cy, cz), ref);
  }
  Box box_of(int first, int last) const
  {
    Box b;
    for (int d = 0; d < 3; ++d) { b.lo[d] = INFINITY; b.hi[d] = -INFINITY; }
    for (int j = first; j <= last; ++j)
      for (int d = 0; d < 3; ++d) {
        const float v = I.xyz[3 * order[j] + d];
        b.lo[d] = std::min(b.lo[d], v);
        b.hi[d] = std::max(b.hi[d], v);
      }
    return b;
  }
  // returns the reference of the subtree over sorted positions [first, last]; l_parent = prefix length of its parent
  int build(int first, int last, int l_parent)
  {
    const int count = last - first + 1;
    const int l_self = host_prefix_len63(keys[first], keys[last]);
    int ref;
    if (count <= kLeafSize) {
      const int leaf = static_cast<int>(I.pts.size() / kLeafSize);
      for (int j = 0; j < kLeafSize; ++j) {
        if (j < count) {
          const int o = order[first + j];
          I.pts.push_back(make_float4(I.xyz[3 * o], I.xyz[3 * o + 1], I.xyz[3 * o + 2], __int_as_float(o)));
        }
        else
          I.pts.push_back(make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(kSentinelIndex)));
      }
      ref = ~leaf;
      for (int b = l_parent / 3 + 1; b <= I.bmax; ++b) {   // 3b > l_parent
        if (3 * b <= l_self) insert(b, keys[first], ref);
        else
          for (int j = first; j <= last; ++j) insert(b, keys[j], ref);  // a leaf that spans several level-b cells
      }
      return ref;
    }
    const int id = static_cast<int>(I.nodes.size());
    I.nodes.emplace_back();
    int split;   // last position of the left child
    if (l_self == 63) split = first + count / 2 - 1;   // a run of equal codes: any cut
    else {
      const unsigned long long bit = 1ull << (62 - l_self);
      split = first;
      while (split + 1 <= last && !(keys[split + 1] & bit)) ++split;
    }
    // prefix_len(parent) < 3b <= prefix_len(this): nodes cut INSIDE a run of equal codes have a parent of prefix 63 and
    // never qualify, the top node of such a run does (lbvh.cu: k_link_cells)
    for (int b = l_parent / 3 + 1; b <= I.bmax && 3 * b <= l_self; ++b) insert(b, keys[first], id);
    const int first_leaf = static_cast<int>(I.pts.size() / kLeafSize);
    const int l = build(first, split, l_self);
    const int r = build(split + 1, last, l_self);
    if (I.node_leaves.size() < I.nodes.size()) I.node_leaves.resize(I.nodes.size());
    I.node_leaves[id] = make_int2(first_leaf, static_cast<int>(I.pts.size() / kLeafSize) - first_leaf);
    const Box a = box_of(first, split), c = box_of(split + 1, last);
    BvhNode nd;
    nd.a = make_float4(a.lo[0], a.lo[1], a.lo[2], a.hi[0]);
    nd.b = make_float4(a.hi[1], a.hi[2], c.lo[0], c.lo[1]);
    nd.c = make_float4(c.lo[2], c.hi[0], c.hi[1], c.hi[2]);
    nd.d = make_int4(l, r, 0, 0);
    I.nodes[id] = nd;
    return id;
  }
};

static void build_index_reference(HostIndex& I, const std::vector<float>& xyz, int bmax);
#ifdef PCLB_TEST_DEVICE_BUILD   // every test in this translation unit searches an index the DEVICE build kernels produced
namespace device_build { inline void build(HostIndex& I, const std::vector<float>& xyz, bool build_cell_table); }
static void build_index(HostIndex& I, const std::vector<float>& xyz, int /*bmax: the build decides*/) { device_build::build(I, xyz, true); }
#else
static void build_index(HostIndex& I, const std::vector<float>& xyz, int bmax) { build_index_reference(I, xyz, bmax); }
#endif
static void build_index_reference(HostIndex& I, const std::vector<float>& xyz, int bmax)
{
  I = HostIndex();
  I.xyz = xyz;
  const int n = static_cast<int>(xyz.size() / 3);
  for (int d = 0; d < 3; ++d) { I.lo[d] = INFINITY; I.hi[d] = -INFINITY; }
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) { I.lo[d] = std::min(I.lo[d], xyz[3 * i + d]); I.hi[d] = std::max(I.hi[d], xyz[3 * i + d]); }
  float ext = 0.f;
  for (int d = 0; d < 3; ++d) ext = std::max(ext, I.hi[d] - I.lo[d]);
  I.scale = ext > 0.f ? 2097152.f / ext : 1.f;
  if (!std::isfinite(I.scale)) I.scale = 1.f;
  Builder B{I, {}, {}, {}, {}};
  std::vector<std::pair<unsigned long long, int>> kv(n);
  for (int i = 0; i < n; ++i) {
    const unsigned long long k = (expand21(morton_cell(xyz[3 * i + 2], I.lo[2], I.scale)) << 2) |
                                 (expand21(morton_cell(xyz[3 * i + 1], I.lo[1], I.scale)) << 1) |
                                 expand21(morton_cell(xyz[3 * i], I.lo[0], I.scale));
    kv[i] = {k, i};
  }
  std::sort(kv.begin(), kv.end());
  B.keys.resize(n);
  B.order.resize(n);
  for (int i = 0; i < n; ++i) { B.keys[i] = kv[i].first; B.order[i] = kv[i].second; }
  I.bmax = bmax;
  float m = 0.f;
  for (int d = 0; d < 3; ++d) m = std::max(m, std::max(std::fabs(I.lo[d]), std::fabs(I.hi[d])));
  m = std::max(m, 2097152.f / I.scale);
  I.margin = 2e-6f * m;
  I.root = B.build(0, n - 1, -1);
  if (bmax >= 1 && !B.entries.empty()) {
    unsigned lg = 6;
    while ((1ull << lg) < 2 * B.entries.size()) ++lg;
    I.log2_slots = lg;
    I.slots.assign(1u << lg, make_uint2(0u, 0u));
    const unsigned shift = 32u - lg, mask = (1u << lg) - 1u;
    for (const auto& e : B.entries) {
      unsigned h = (e.first * 0x9E3779B1u) >> shift;
      for (;;) {
        if (I.slots[h].x == e.first) {
          if (static_cast<int>(I.slots[h].y) != e.second) std::printf("BUILD ERROR: cell %u mapped to two subtrees\n", e.first);
          break;
        }
        if (I.slots[h].x == 0u) { I.slots[h] = make_uint2(e.first, static_cast<unsigned>(e.second)); break; }
        h = (h + 1u) & mask;
      }
    }
  }
}

// ---- brute force under the library's own distance expression and tie rule --------------------------------------------
struct Truth { float d1 = INFINITY, d2nd = INFINITY; int idx = kSentinelIndex; };
static Truth brute(const std::vector<float>& xyz, const float q[3], float gate)
{
  Truth t;
  const int n = static_cast<int>(xyz.size() / 3);
  std::vector<float> all(n);
  for (int i = 0; i < n; ++i) {
    const float d = dist2_rn(q[0], q[1], q[2], xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    all[i] = d;
    if (d <= gate && (d < t.d1 || (d == t.d1 && i < t.idx))) { t.d1 = d; t.idx = i; }
  }
  for (int i = 0; i < n; ++i)
    if (i != t.idx) t.d2nd = std::min(t.d2nd, all[i]);
  return t;
}

#ifdef PCLB_TEST_DEVICE_BUILD
#include "device_build.h"
#endif
