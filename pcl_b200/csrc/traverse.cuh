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

// -DPCLB_STATS (experimental builds only, tools/): per-walk event counts, accumulated per thread and flushed once
#ifdef PCLB_STATS
struct WalkStats { unsigned n[8]; };  // 0 lookups 1 node visits 2 leaf scans 3 pushes 4 home seeds 5 rooted walks 6 cells pushed 7 queries
#define PCLB_STAT(st, k) (++(st).n[k])
#else
struct WalkStats {};
#define PCLB_STAT(st, k) ((void)0)
#endif

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
//   void prune(float d2)           — a subtree / cell whose every point is at least d2 away was skipped
//   void leaf(const float4* leaf_pts, int first_pos) — examine kLeafSize consecutive points
// Nearer child first; the farther one is pushed with its bound and re-tested when popped.
// The stack is the caller's: walks that start from several subtrees at once (cell-table starts, below) pre-fill it
// with (subtree, lower bound) pairs and pass node = kDone; a plain walk passes sp = 0 and node = root.
// skip_a / skip_b: leaf references (~leaf) the caller has already scanned (seed leaves) — not scanned again.
// Returns false on stack overflow (caller raises the device error flag).
constexpr int kDone = 0x7fffffff;  // "no node": internal ids are >= 0 and < 2^31-1, leaves are negative

template <typename Visitor>
__device__ __forceinline__ bool walk(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts,
                                     int* __restrict__ stack_node, float* __restrict__ stack_dist, int sp, int node,
                                     float qx, float qy, float qz, Visitor& v, int skip_a, int skip_b,
                                     WalkStats& ws)
{
  bool ok = true;
  if (node == kDone) {
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
  while (node != kDone) {
    while (node >= 0 && node != kDone) {
      const float4* np = reinterpret_cast<const float4*>(nodes + node);
      const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
      const int4 d = __ldg(reinterpret_cast<const int4*>(np + 3));
      PCLB_STAT(ws, 1);
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
            PCLB_STAT(ws, 3);
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
    if (node != skip_a && node != skip_b) {
      const int leaf = ~node;
      v.leaf(pts + (size_t)leaf * kLeafSize, leaf * kLeafSize);
      PCLB_STAT(ws, 2);
    }
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

// plain walk from `root` with a private stack
template <typename Visitor>
__device__ __forceinline__ bool traverse(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts,
                                         int root, float qx, float qy, float qz, Visitor& v)
{
  int stack_node[kStackSize];
  float stack_dist[kStackSize];
  WalkStats ws;
  return walk(nodes, pts, stack_node, stack_dist, 0, root, qx, qy, qz, v, kDone, kDone, ws);
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

// ---- cell table: walks that start at the candidate ball instead of the root ------------------------------------
// The tree is a radix tree over Morton codes, so for every level b (cells of 2^-b of the index's frame per axis) and every
// OCCUPIED cell there is one deepest node — or leaf — that holds EVERY indexed point of that cell (lbvh.cu inserts exactly
// these (level, cell) -> reference pairs into one open-addressing hash table, levels 1..bmax).  A 1-NN walk that already
// holds a candidate at squared distance `best` only has to look at indexed points inside the closed ball B(q, r),
// r = sqrt(best): all of them lie in the box [q - r, q + r]^3, the build's quantiser is monotone, so their cell
// coordinates lie between those of the box's two corners; at the level where the box is at most one cell wide that is
// at most 2 x 2 x 2 cells.  The walk therefore looks those cells up, pushes their subtrees (with an arithmetic lower
// bound on the distance to the cell) and runs the ordinary exact traversal from there: the ~20 levels between the root
// and the neighbourhood of the query are never touched, and the cost follows the number of points near the ball, not
// the size of the cloud.  Exactness needs nothing beyond the monotone quantiser (an integer argument) — the float
// bounds only ever prune a cell when every point of it is provably farther than the current best.
// walk cells 2^bias times wider than the minimum (fewer cells to look up, deeper subtrees): 0 and 1 measured equal
// (profiles/r2d: 20.08 vs 20.11 ms per 10-iteration step)
constexpr int kCellLevelBias = 0;
constexpr int kCellMaxBits = 10;  // finest level: 3 * 10 bits of cell coordinates + a level marker fit one 32-bit key

struct CellTable {
  const uint2* slots = nullptr;  // {key, reference}; key 0 = empty.  nullptr: the index has no table (walks start at the root)
  unsigned shift = 0;            // 32 - log2(#slots)
  unsigned mask = 0;             // #slots - 1
  int bmax = 0;                  // finest level present
  float lo[3] = {0.f, 0.f, 0.f}; // frame of the index (same numbers the Morton keys were built with)
  float scale = 1.f;
  float inv_scale = 1.f;
  float margin = 0.f;            // absolute slack of a cell boundary computed in fp32 (see cell_gap2)
};

struct TreeView {
  const BvhNode* nodes;
  const float4* pts;
  int root;
  CellTable cells;
};

// device-side view of an index (host helper)
inline TreeView tree_view(const Index& idx)
{
  TreeView t;
  t.nodes = idx.nodes.p;
  t.pts = idx.pts.p;
  t.root = idx.root;
  if (idx.cell_slots.p && idx.cells.log2_slots) {
    t.cells.slots = idx.cell_slots.p;
    t.cells.shift = 32u - idx.cells.log2_slots;
    t.cells.mask = (1u << idx.cells.log2_slots) - 1u;
    t.cells.bmax = idx.cells.bmax;
    for (int d = 0; d < 3; ++d)
      t.cells.lo[d] = idx.lo[d];
    t.cells.scale = idx.morton_scale;
    t.cells.inv_scale = 1.f / idx.morton_scale;
    t.cells.margin = idx.cells.margin;
  }
  return t;
}

__host__ __device__ __forceinline__ unsigned cell_key(int b, unsigned cx, unsigned cy, unsigned cz)
{
  return (1u << (3 * b)) | (cz << (2 * b)) | (cy << b) | cx;
}

__device__ __forceinline__ int cell_lookup(const CellTable& C, unsigned key)
{
  unsigned h = (key * 0x9E3779B1u) >> C.shift;
  for (;;) {
    const uint2 s = __ldg(C.slots + h);
    if (s.x == key)
      return (int)s.y;
    if (s.x == 0u)
      return kDone;
    h = (h + 1u) & C.mask;
  }
}

// Lower bound (squared, rounded down) on the per-axis distance from coordinate q to every point whose level-(21-s) cell
// along this axis is `other`, given that q's own cell is `home` != other.  Points of a cell `other` > home have
// morton_cell(x) >= K = other << s, points of a cell < home have morton_cell(x) < K = home << s.  morton_cell is
// floor(fl(fl(x - lo) * scale)) clamped, so its switch-over point differs from lo + K / scale by at most a few ulps of
// the frame (|lo|, extent); `margin` = 2e-6 * max(|lo|, |hi|, extent) covers that with a wide berth, and the factor
// (1 - 2e-6) covers the rounding of the subtraction and of the squared distances compared against it (~3e-7 relative).
__device__ __forceinline__ float cell_gap2(const CellTable& C, int axis, float q, unsigned home, unsigned other, int s)
{
  const unsigned K = (other > home ? other : home) << s;
  const float xk = __fadd_rn(C.lo[axis], __fmul_rn((float)K, C.inv_scale));
  float g = other > home ? __fsub_rn(xk, q) : __fsub_rn(q, xk);
  g = __fsub_rd(__fmul_rd(g, 0.999998f), C.margin);
  g = fmaxf(g, 0.f);
  return __fmul_rd(g, g);
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


// ---- exact 1-NN with cell-table starts ----------------------------------------------------------------------------
//   1. seed: the previous iteration's match (its leaf is scanned first), if any;
//   2. home seed: when there is no candidate yet, or the candidate ball is much wider than the finest cells (the query
//      moved far since the seed was found), the leaf reached by a greedy descent from the finest occupied cell that
//      contains q is scanned as well — a candidate at about the local point spacing, whatever the motion was;
//   3. the <= 8 cells the candidate ball reaches are looked up and walked (see CellTable).  Without a table, or when the
//      ball is wider than half the frame, the walk starts at the root.
// INFLATE > 1 widens the ball (TRACK visitors: every point outside the visited cells is farther than the inflated
// radius, which is what bounds Nearest1T::lower_bound2 from above).
template <bool TRACK>
__device__ __forceinline__ bool nearest1(const TreeView& T, float qx, float qy, float qz, Nearest1T<TRACK>& v,
                                         int seed_pos, float inflate, WalkStats& ws)
{
  int stack_node[kStackSize];
  float stack_dist[kStackSize];
  int sp = 0;
  int skip_a = kDone, skip_b = kDone;
  const CellTable& C = T.cells;
  if (seed_pos >= 0) {
    const int leaf = seed_pos / kLeafSize;
    v.template scan<true>(T.pts + (size_t)leaf * kLeafSize, leaf * kLeafSize);
    PCLB_STAT(ws, 2);
    skip_a = ~leaf;
  }
  bool rooted = true;
  if (C.slots != nullptr) {
    const unsigned cqx = morton_cell(qx, C.lo[0], C.scale), cqy = morton_cell(qy, C.lo[1], C.scale),
                   cqz = morton_cell(qz, C.lo[2], C.scale);
    const int smin = 21 - C.bmax;
    unsigned ax, ay, az, bx, by, bz;
    float r;
    int s;
    // level at which the candidate ball's box is at most one cell wide on every axis: d <= 2^s
    auto ball_level = [&]() {
      r = __fmul_ru(__fsqrt_ru(v.best), inflate);
      ax = morton_cell(__fsub_rd(qx, r), C.lo[0], C.scale); bx = morton_cell(__fadd_ru(qx, r), C.lo[0], C.scale);
      ay = morton_cell(__fsub_rd(qy, r), C.lo[1], C.scale); by = morton_cell(__fadd_ru(qy, r), C.lo[1], C.scale);
      az = morton_cell(__fsub_rd(qz, r), C.lo[2], C.scale); bz = morton_cell(__fadd_ru(qz, r), C.lo[2], C.scale);
      const unsigned d = max(max(bx - ax, by - ay), bz - az);
      s = (d <= 1u ? 0 : 32 - __clz((int)(d - 1u))) + kCellLevelBias;
    };
    ball_level();
    if (v.best_pos < 0 || s > smin + 2) {
      // finest occupied cell that contains q: occupancy is monotone in the level (a coarser cell contains the finer
      // one), so a binary search over the levels needs ~log2(bmax) lookups
      int ref = kDone;
      int lo_b = 0, hi_b = C.bmax;  // invariant: level lo_b is occupied (level 0 = the root), levels > hi_b are not
      while (lo_b < hi_b) {
        const int b = (lo_b + hi_b + 1) >> 1;
        const int sh = 21 - b;
        const int rr = cell_lookup(C, cell_key(b, cqx >> sh, cqy >> sh, cqz >> sh));
        PCLB_STAT(ws, 0);
        if (rr != kDone) {
          lo_b = b;
          ref = rr;
        }
        else
          hi_b = b - 1;
      }
      if (ref != kDone) {
        while (ref >= 0) {  // greedy descent: nearer child, no stack
          const float4* np = reinterpret_cast<const float4*>(T.nodes + ref);
          const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
          const int4 d = __ldg(reinterpret_cast<const int4*>(np + 3));
          const float dl = box_dist2_rn(qx, qy, qz, a.x, a.y, a.z, a.w, b.x, b.y);
          const float dr = box_dist2_rn(qx, qy, qz, b.z, b.w, c.x, c.y, c.z, c.w);
          ref = dr < dl ? d.y : d.x;
          PCLB_STAT(ws, 1);
        }
        if (ref != skip_a) {
          const int leaf = ~ref;
          v.template scan<true>(T.pts + (size_t)leaf * kLeafSize, leaf * kLeafSize);
          PCLB_STAT(ws, 2);
          PCLB_STAT(ws, 4);
          skip_b = ref;
          ball_level();
        }
      }
    }
    if (s < smin)
      s = smin;
    if (s <= 20) {
      rooted = false;
      const int b = 21 - s;
      const unsigned hx = cqx >> s, hy = cqy >> s, hz = cqz >> s;
      const unsigned ox = (ax >> s) + (bx >> s) - hx, oy = (ay >> s) + (by >> s) - hy, oz = (az >> s) + (bz >> s) - hz;
      const unsigned E = (ox != hx ? 1u : 0u) | (oy != hy ? 2u : 0u) | (oz != hz ? 4u : 0u);
      const float gx2 = (E & 1u) ? cell_gap2(C, 0, qx, hx, ox, s) : 0.f;
      const float gy2 = (E & 2u) ? cell_gap2(C, 1, qy, hy, oy, s) : 0.f;
      const float gz2 = (E & 4u) ? cell_gap2(C, 2, qz, hz, oz, s) : 0.f;
      for (unsigned m = E;; m = (m - 1u) & E) {  // submasks of E, the home cell (m = 0) last = popped first
        const float bound = __fadd_rd(__fadd_rd((m & 1u) ? gx2 : 0.f, (m & 2u) ? gy2 : 0.f), (m & 4u) ? gz2 : 0.f);
        if (bound <= v.bound()) {
          const int ref = cell_lookup(C, cell_key(b, (m & 1u) ? ox : hx, (m & 2u) ? oy : hy, (m & 4u) ? oz : hz));
          PCLB_STAT(ws, 0);
          if (ref != kDone && ref != skip_a && ref != skip_b) {
            stack_node[sp] = ref;
            stack_dist[sp] = bound;
            ++sp;
            PCLB_STAT(ws, 6);
          }
        }
        else
          v.prune(bound);
        if (m == 0u)
          break;
      }
      v.prune(__fmul_rd(r, r));  // every indexed point outside the cells lies outside [q - r, q + r]^3
    }
  }
  if (rooted) {
    PCLB_STAT(ws, 5);
    stack_node[0] = T.root;
    stack_dist[0] = 0.f;
    sp = 1;
  }
  PCLB_STAT(ws, 7);
  return walk(T.nodes, T.pts, stack_node, stack_dist, sp, kDone, qx, qy, qz, v, skip_a, skip_b, ws);
}

// ---- temporal coherence between ICP iterations (used by icp.cu: k_search<.., TRACK>) ------------------------------------
constexpr float kRelMargin = 1e-5f;  // >> fp32 rounding of the distances involved (~2e-7): keeps the skip test exact
// TRACK walks look at a ball kTrackInflate times wider than the candidate distance: every point they do not see is
// then at least that much farther than the match, which is the head-room the next iteration's skip test lives on
constexpr float kTrackInflate = 2.5f;

// Temporal-coherence test.  Let m be the previous match at distance D1 from the old query position, L a lower
// bound on the distance from the old position to every other point, and delta the distance the query moved.
// Triangle inequality: the new distance to m is <= D1 + delta, the new distance to any other x is >= L - delta.
// If D1 + delta < L - delta (with a relative safety margin far above fp32 round-off) m is still the UNIQUE nearest
// neighbour, so the exact search result is (m, dist2(p_new, m)) and the tree walk can be skipped.
__device__ __forceinline__ bool still_nearest(float prev_d2, float prev_lb, float delta, float* new_lb)
{
  if (!(prev_lb > 0.f))
    return false;
  const float D1 = sqrtf(prev_d2);
  const float lhs = (D1 + 2.f * delta) * (1.f + kRelMargin);
  const float rhs = prev_lb * (1.f - kRelMargin);
  *new_lb = (prev_lb - delta * (1.f + kRelMargin)) * (1.f - kRelMargin);
  return lhs < rhs && *new_lb > 0.f;
}

}  // namespace pclb200
