// traverse.cuh — exact nearest-neighbour traversal of the Morton LBVH, one query per thread.
//
// Exactness contract (the CPU oracle under oracle/ states the same rules):
//   d2(q,p)   = ((dx*dx) + dy*dy) + dz*dz, fp32 round-to-nearest, NO fma  (flann::L2_Simple order)
//   box bound = same expression on the per-axis gap max(lo-q, q-hi, 0).  Rounding is monotone, so
//               bound <= d2(q,p) holds IN FP32 for every p inside the box: pruning on
//               `bound > worst` can never drop a true neighbour, and visiting on `bound == worst`
//               lets an equal-distance point with a smaller index win (canonical tie rule).
#pragma once
#include "internal.cuh"

namespace pclb200 {

__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, float px, float py, float pz)
{
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float box_dist2_rn(float qx, float qy, float qz, float lx, float ly, float lz,
                                              float hx, float hy, float hz)
{
  float dx = fmaxf(fmaxf(__fsub_rn(lx, qx), __fsub_rn(qx, hx)), 0.f);
  float dy = fmaxf(fmaxf(__fsub_rn(ly, qy), __fsub_rn(qy, hy)), 0.f);
  float dz = fmaxf(fmaxf(__fsub_rn(lz, qz), __fsub_rn(qz, hz)), 0.f);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// Generic depth-first traversal ("while-while"): the inner loop walks internal nodes only, so the lanes of a
// warp that are still descending are not serialised against lanes scanning a leaf; leaf scans (the long,
// fully unrolled body) run once the warp has left the inner loop.  `Visitor` provides:
//   float bound() const            — current pruning distance (subtrees with box bound > bound() are skipped)
//   void leaf(const float4* leaf_pts, int first_pos) — examine kLeafSize consecutive points
// Nearer child first; the farther one is pushed with its bound and re-tested when popped.
// Returns false on stack overflow (caller raises the device error flag).
constexpr int kDone = 0x7fffffff;  // "no node": internal ids are >= 0 and < 2^31-1, leaves are negative

template <typename Visitor>
__device__ __forceinline__ bool traverse(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts,
                                         int root, float qx, float qy, float qz, Visitor& v)
{
  int stack_node[kStackSize];
  float stack_dist[kStackSize];
  int sp = 0;
  int node = root;
  bool ok = true;
  while (node != kDone) {
    while (node >= 0 && node != kDone) {
      const float4* np = reinterpret_cast<const float4*>(nodes + node);
      const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
      const int4 d = __ldg(reinterpret_cast<const int4*>(np + 3));
      float dl = box_dist2_rn(qx, qy, qz, a.x, a.y, a.z, a.w, b.x, b.y);
      float dr = box_dist2_rn(qx, qy, qz, b.z, b.w, c.x, c.y, c.z, c.w);
      int nl = d.x, nr = d.y;
      if (dr < dl) {
        float t = dl; dl = dr; dr = t;
        int ti = nl; nl = nr; nr = ti;
      }
      const float bnd = v.bound();
      if (dl <= bnd) {
        if (dr <= bnd) {
          if (sp < kStackSize) {
            stack_node[sp] = nr;
            stack_dist[sp] = dr;
            ++sp;
          }
          else
            ok = false;
        }
        else
          v.prune(dr);
        node = nl;
      }
      else {
        v.prune(dl);  // dl <= dr: both subtrees are at least this far
        node = kDone;
        while (sp > 0) {
          --sp;
          if (stack_dist[sp] <= bnd) {
            node = stack_node[sp];
            break;
          }
          v.prune(stack_dist[sp]);
        }
      }
    }
    if (node == kDone)
      break;
    const int leaf = ~node;
    v.leaf(pts + (size_t)leaf * kLeafSize, leaf * kLeafSize);
    node = kDone;
    const float bnd = v.bound();
    while (sp > 0) {
      --sp;
      if (stack_dist[sp] <= bnd) {
        node = stack_node[sp];
        break;
      }
      v.prune(stack_dist[sp]);
    }
  }
  return ok;
}

// ---- Morton cell coordinates (shared by the build, lbvh.cu, and by walks that start below the root) --------------
// 21-bit cell coordinate of x along one axis: the build's quantiser, monotone non-decreasing in x.
__device__ __forceinline__ unsigned morton_cell(float x, float lo, float scale)
{
  return (unsigned)fminf(fmaxf(__fmul_rn(__fsub_rn(x, lo), scale), 0.f), 2097151.f);
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v)
{
  v &= 0x1fffffULL;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

// ---- start node from the prefix tables -------------------------------------------------------------------------------
// Index::top[b - kTopMinBits][P] = the deepest node that holds EVERY indexed point whose 63-bit Morton code starts with
// the 3b-bit prefix P (root where no such point exists).  A walk that only needs the points of the closed box
// [q - r, q + r]^3 may start there if the box lies inside ONE 3b-cell: the quantiser is monotone, so every indexed point
// of the box has a cell coordinate between those of the two corners, i.e. the same prefix — an integer test.  The finest
// table whose cell holds the box is used; if even the coarsest does not, the walk starts at the root.
constexpr int kTopMinBits = 4, kTopMaxBits = 8, kTopLevels = kTopMaxBits - kTopMinBits + 1;

struct TopTables {
  const int* table[kTopLevels];  // nullptr = level not built
  float lo[3];
  float scale;
};

__device__ __forceinline__ int top_start(const TopTables& T, int root, float qx, float qy, float qz, float r)
{
  const unsigned ax = morton_cell(__fsub_rd(qx, r), T.lo[0], T.scale), bx = morton_cell(__fadd_ru(qx, r), T.lo[0], T.scale);
  const unsigned ay = morton_cell(__fsub_rd(qy, r), T.lo[1], T.scale), by = morton_cell(__fadd_ru(qy, r), T.lo[1], T.scale);
  const unsigned az = morton_cell(__fsub_rd(qz, r), T.lo[2], T.scale), bz = morton_cell(__fadd_ru(qz, r), T.lo[2], T.scale);
  const unsigned diff = (ax ^ bx) | (ay ^ by) | (az ^ bz);  // bit k set: some axis' corners differ in cell bit k
  // the corners share their top b bits on every axis  <=>  diff < 2^(21 - b)
  const int common = diff == 0 ? 21 : __clz(diff) - 11;  // number of shared leading bits (of 21)
#pragma unroll
  for (int b = kTopMaxBits; b >= kTopMinBits; --b) {
    const int* tb = T.table[b - kTopMinBits];
    if (tb != nullptr && common >= b) {
      const unsigned long long key = (expand21(az) << 2) | (expand21(ay) << 1) | expand21(ax);
      return __ldg(tb + (size_t)(key >> (63 - 3 * b)));
    }
  }
  return root;
}

// ---- start node of a SEEDED walk ---------------------------------------------------------------------------------
// A walk that already holds a candidate at squared distance `best` (the previous iteration's match, re-measured) only
// has to look at points inside the closed ball B(q, sqrt(best)).  The tree is a radix tree over Morton codes, so a
// SPATIAL cell X (BvhNode::d.z) holds every indexed point whose code starts with X's prefix, i.e. every indexed point
// of an axis-aligned region that contains box(X).  Hence, if the ball lies inside box(X), every point of the ball
// belongs to X's subtree and the walk can start at X instead of the root: the 20-odd levels above X are never touched.
// climb_start walks up from the seed's leaf (parent arrays of the index) until such an X is found, at most kClimbLevels
// levels; otherwise it returns the root.  `factor` (> 1) inflates the radius: 1.00001 covers the fp32 rounding of the
// distances that are compared later (relative error ~3e-7), larger values buy a larger exit distance for the
// temporal-coherence bound.  *exit2 = squared distance from q to the outside of box(X) (rounded down; +inf for the
// root): a lower bound on the squared distance to every point the restricted walk does not see.
constexpr int kClimbLevels = 12;

__device__ __forceinline__ int climb_start(const BvhNode* __restrict__ nodes, const int* __restrict__ node_parent,
                                           const int* __restrict__ leaf_parent, int root, int seed_leaf, float qx,
                                           float qy, float qz, float best, float factor, float* __restrict__ exit2)
{
  *exit2 = __int_as_float(0x7f800000);
  if (node_parent == nullptr || !(best < __int_as_float(0x7f800000)))
    return root;
  const float r = __fmul_ru(__fsqrt_ru(best), factor);
  int child = ~seed_leaf;
  int cur = __ldg(leaf_parent + seed_leaf);
#pragma unroll 1
  for (int lvl = 0; lvl < kClimbLevels && cur >= 0; ++lvl) {
    const float4* np = reinterpret_cast<const float4*>(nodes + cur);
    const int4 d = __ldg(reinterpret_cast<const int4*>(np + 3));
    const bool is_left = d.x == child;
    if ((d.z >> (is_left ? 0 : 1)) & 1) {
      const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
      const float lox = is_left ? a.x : b.z, loy = is_left ? a.y : b.w, loz = is_left ? a.z : c.x;
      const float hix = is_left ? a.w : c.y, hiy = is_left ? b.x : c.z, hiz = is_left ? b.y : c.w;
      // distance from q to the nearest face, rounded down; negative when q is outside the box
      const float e = fminf(fminf(fminf(__fsub_rd(qx, lox), __fsub_rd(hix, qx)), fminf(__fsub_rd(qy, loy), __fsub_rd(hiy, qy))),
                            fminf(__fsub_rd(qz, loz), __fsub_rd(hiz, qz)));
      if (e >= r) {
        *exit2 = __fmul_rd(e, e);
        return child;
      }
    }
    child = cur;
    cur = __ldg(node_parent + cur);
  }
  return root;
}

// ---- 1-NN visitor: lexicographic (d2, original index) minimum --------------------------------
// TRACK = true additionally maintains a LOWER BOUND on the distance to every point other than the best one:
//   m2     = second smallest d2 among the points actually evaluated during the walk
//   pruned = smallest box bound of any subtree the walk skipped (every point in it is at least that far)
// lower_bound2() = min(m2, pruned) is what lets the next ICP iteration prove, by the triangle inequality, that the
// previous match is still the nearest neighbour without walking the tree again (icp.cu, k_search*).
template <bool TRACK>
struct Nearest1T {
  float qx, qy, qz;
  float best;   // current best d2 (initialised to the gate)
  int best_idx; // original index of the best point (kSentinelIndex = none yet)
  int best_pos; // its position in the Morton array
  float m1, m2, pruned_min;  // TRACK only (initialise to +inf)
  __device__ __forceinline__ float bound() const { return best; }
  __device__ __forceinline__ void prune(float d)
  {
    if (TRACK)
      pruned_min = fminf(pruned_min, d);
  }
  __device__ __forceinline__ float lower_bound2() const { return fminf(m2, pruned_min); }
  // scan that only improves (best, best_idx, best_pos): used for the seed leaf, which the walk visits again
  template <bool COUNT>
  __device__ __forceinline__ void scan(const float4* lp, int first_pos)
  {
    // distances first, one min-reduction, and the (rare) index bookkeeping only when the leaf can improve the
    // best: most visited leaves do not, so the common path is 8 flops per point plus one compare per leaf.
#pragma unroll
    for (int j0 = 0; j0 < kLeafSize; j0 += 8) {
      float4 p[8];
      float d[8];
      float m = __int_as_float(0x7f800000);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        p[j] = ldg4(lp + j0 + j);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d[j] = dist2_rn(qx, qy, qz, p[j].x, p[j].y, p[j].z);
        m = fminf(m, d[j]);
        if (TRACK && COUNT) {
          m2 = fminf(m2, fmaxf(m1, d[j]));
          m1 = fminf(m1, d[j]);
        }
      }
      if (m <= best) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (d[j] <= best) {
            const int oi = __float_as_int(p[j].w);
            if (d[j] < best || oi < best_idx) {
              best = d[j];
              best_idx = oi;
              best_pos = first_pos + j0 + j;
            }
          }
        }
      }
    }
  }
  __device__ __forceinline__ void leaf(const float4* lp, int first_pos) { scan<true>(lp, first_pos); }
};
using Nearest1 = Nearest1T<false>;

// ---- packet traversal: one walk of the tree per WARP, shared by its 32 Morton-adjacent queries ---------------
// All lanes follow the same path (no divergence, every node / leaf line is one broadcast load, the stack is one
// per-warp array in shared memory).  A subtree is entered when ANY lane still needs it (box bound <= that lane's
// current best) — so every lane sees a superset of the leaves its own exact search would visit and its result is
// still the exact lexicographic minimum.  The nearer child is the one with the smaller warp-minimum bound; the other
// is pushed with that minimum and re-tested on pop against the warp-maximum best (conservative).
// Lanes without a query pass best = -1 (never want anything).  Must be called by all 32 lanes.
constexpr int kWarpStack = kStackSize;

template <typename V>
__device__ __forceinline__ bool traverse_packet(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts,
                                                int root, float qx, float qy, float qz, V& v,
                                                int* __restrict__ wnode, float* __restrict__ wdist)
{
  const unsigned full = 0xffffffffu;
  const unsigned INF_BITS = 0x7f800000u;
  int sp = 0;
  int node = root;
  bool ok = true;
  while (node != kDone) {
    while (node >= 0 && node != kDone) {
      const float4* np = reinterpret_cast<const float4*>(nodes + node);
      const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
      const int4 d = __ldg(reinterpret_cast<const int4*>(np + 3));
      const float dl = box_dist2_rn(qx, qy, qz, a.x, a.y, a.z, a.w, b.x, b.y);
      const float dr = box_dist2_rn(qx, qy, qz, b.z, b.w, c.x, c.y, c.z, c.w);
      const float bnd = v.best;
      const unsigned ml = __reduce_min_sync(full, dl <= bnd ? __float_as_uint(dl) : INF_BITS);
      const unsigned mr = __reduce_min_sync(full, dr <= bnd ? __float_as_uint(dr) : INF_BITS);
      // Lower-bound bookkeeping (TRACK visitors): a lane that does not need a child notes ITS OWN bound to that
      // box, whether or not the warp enters the subtree for other lanes.  (The warp-minimum stored with a pushed
      // entry is a minimum over the lanes that WANTED it, so it bounds only those lanes when the entry is later
      // discarded; the others are covered here.)
      if (!(dl <= bnd))
        v.prune(dl);
      if (!(dr <= bnd))
        v.prune(dr);
      if (ml == INF_BITS && mr == INF_BITS) {
        node = kDone;
        const unsigned wmax = __reduce_max_sync(full, __float_as_uint(fmaxf(bnd, 0.f)));
        while (sp > 0) {
          --sp;
          if (__float_as_uint(wdist[sp]) <= wmax) {
            node = wnode[sp];
            break;
          }
          v.prune(wdist[sp]);  // warp-minimum bound of a discarded entry: valid (weaker) bound for every lane
        }
      }
      else if (ml != INF_BITS && mr != INF_BITS) {
        const bool left_first = ml <= mr;
        if (sp < kWarpStack) {
          wnode[sp] = left_first ? d.y : d.x;
          wdist[sp] = __uint_as_float(left_first ? mr : ml);
          ++sp;
        }
        else
          ok = false;
        node = left_first ? d.x : d.y;
      }
      else
        node = ml != INF_BITS ? d.x : d.y;
    }
    if (node == kDone)
      break;
    const int leaf = ~node;
    v.leaf(pts + (size_t)leaf * kLeafSize, leaf * kLeafSize);
    node = kDone;
    const unsigned wmax = __reduce_max_sync(full, __float_as_uint(fmaxf(v.best, 0.f)));
    while (sp > 0) {
      --sp;
      if (__float_as_uint(wdist[sp]) <= wmax) {
        node = wnode[sp];
        break;
      }
      v.prune(wdist[sp]);
    }
  }
  return ok;
}

}  // namespace pclb200
