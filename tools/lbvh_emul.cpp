// lbvh_emul.cpp — offline (CPU) emulator of the device LBVH + packet traversal, used to compare tree-construction
// variants by the number of node / leaf visits per 32-query packet before spending GPU time on them.
//   g++ -O2 -std=c++17 -fopenmp tools/lbvh_emul.cpp -o /tmp/lbvh_emul && /tmp/lbvh_emul [n] [variant] [leaf]
// variants: 0 = fixed blocks of `leaf` consecutive Morton points (what lbvh.cu builds)
//           1 = radix-tree cells: split by Morton bits until a range holds <= `leaf` points (cell-aligned leaves)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

struct P3 { float x, y, z; };
struct Box { float lo[3], hi[3]; };
struct Node { Box b[2]; int child[2]; };           // child >= 0 internal, < 0: ~leaf
struct Leaf { int first, count; Box box; };

static uint64_t expand21(uint64_t v)
{
  v &= 0x1fffffULL;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}
static uint64_t morton(const P3& p, const float* lo, float s)
{
  auto q = [&](float v, float l) { return (uint64_t)std::min(std::max((v - l) * s, 0.f), 2097151.f); };
  return (expand21(q(p.z, lo[2])) << 2) | (expand21(q(p.y, lo[1])) << 1) | expand21(q(p.x, lo[0]));
}
static uint64_t hilbert(const P3& p, const float* lo, float s)
{
  uint32_t X[3];
  auto q = [&](float v, float l) { return (uint32_t)std::min(std::max((v - l) * s, 0.f), 2097151.f); };
  X[0] = q(p.x, lo[0]); X[1] = q(p.y, lo[1]); X[2] = q(p.z, lo[2]);
  const uint32_t M = 1u << 20;
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    uint32_t Pm = Q - 1;
    for (int i = 0; i < 3; ++i) {
      if (X[i] & Q) X[0] ^= Pm;
      else { uint32_t t = (X[0] ^ X[i]) & Pm; X[0] ^= t; X[i] ^= t; }
    }
  }
  X[1] ^= X[0]; X[2] ^= X[1];
  uint32_t t = 0;
  for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  return (expand21(X[0]) << 2) | (expand21(X[1]) << 1) | expand21(X[2]);
}
static Box box_of(const std::vector<P3>& p, int a, int b)
{
  Box r;
  for (int d = 0; d < 3; ++d) { r.lo[d] = 1e30f; r.hi[d] = -1e30f; }
  for (int i = a; i < b; ++i) {
    const float v[3] = {p[i].x, p[i].y, p[i].z};
    for (int d = 0; d < 3; ++d) { r.lo[d] = std::min(r.lo[d], v[d]); r.hi[d] = std::max(r.hi[d], v[d]); }
  }
  return r;
}
static Box merge(const Box& a, const Box& b)
{
  Box r;
  for (int d = 0; d < 3; ++d) { r.lo[d] = std::min(a.lo[d], b.lo[d]); r.hi[d] = std::max(a.hi[d], b.hi[d]); }
  return r;
}
static float bdist(const P3& q, const Box& b)
{
  float dx = std::max(std::max(b.lo[0] - q.x, q.x - b.hi[0]), 0.f);
  float dy = std::max(std::max(b.lo[1] - q.y, q.y - b.hi[1]), 0.f);
  float dz = std::max(std::max(b.lo[2] - q.z, q.z - b.hi[2]), 0.f);
  return dx * dx + dy * dy + dz * dz;
}
static float pdist(const P3& q, const P3& p)
{
  float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
  return dx * dx + dy * dy + dz * dz;
}

struct Tree {
  std::vector<P3> pts;
  std::vector<uint64_t> keys;
  std::vector<Node> nodes;
  std::vector<Leaf> leaves;
  int root = 0;
};

// generic radix split of key range [a,b) of `keys` (one key per element); elements are leaf ids or points
static int split_pos(const std::vector<uint64_t>& k, int a, int b)
{
  uint64_t fa = k[a], fb = k[b - 1];
  if (fa == fb) return (a + b) / 2;
  int prefix = __builtin_clzll(fa ^ fb);
  int lo = a, hi = b - 1;  // find last index whose key shares more than `prefix` bits with fa
  while (lo + 1 < hi) {
    int mid = (lo + hi) / 2;
    if (__builtin_clzll(fa ^ k[mid]) > prefix || fa == k[mid]) lo = mid; else hi = mid;
  }
  return lo + 1;
}

static int build_blocks(Tree& t, const std::vector<uint64_t>& lk, int a, int b, Box* out)
{
  if (b - a == 1) { *out = t.leaves[a].box; return ~a; }
  int s = split_pos(lk, a, b);
  Node nd;
  int id = (int)t.nodes.size();
  t.nodes.push_back(nd);
  Box bl, br;
  int cl = build_blocks(t, lk, a, s, &bl), cr = build_blocks(t, lk, s, b, &br);
  t.nodes[id].b[0] = bl; t.nodes[id].b[1] = br; t.nodes[id].child[0] = cl; t.nodes[id].child[1] = cr;
  *out = merge(bl, br);
  return id;
}
static int build_cells(Tree& t, int a, int b, int L, Box* out)
{
  if (b - a <= L) {
    Leaf lf{a, b - a, box_of(t.pts, a, b)};
    t.leaves.push_back(lf);
    *out = lf.box;
    return ~(int)(t.leaves.size() - 1);
  }
  int s = split_pos(t.keys, a, b);
  int id = (int)t.nodes.size();
  t.nodes.push_back(Node());
  Box bl, br;
  int cl = build_cells(t, a, s, L, &bl), cr = build_cells(t, s, b, L, &br);
  t.nodes[id].b[0] = bl; t.nodes[id].b[1] = br; t.nodes[id].child[0] = cl; t.nodes[id].child[1] = cr;
  *out = merge(bl, br);
  return id;
}

int main(int argc, char** argv)
{
  int n = argc > 1 ? atoi(argv[1]) : 1000000;
  int variant = argc > 2 ? atoi(argv[2]) : 0;
  int L = argc > 3 ? atoi(argv[3]) : 8;
  int qorder = argc > 4 ? atoi(argv[4]) : 1;  // 1 hilbert, 0 morton
  const double side = std::sqrt(n / 1e5);      // same density as the 10 M / 100 unit^2 bench cloud
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(0.0, side);
  std::normal_distribution<double> N(0.0, 0.002);
  auto surf = [&](std::vector<P3>& v) {
    v.resize(n);
    for (auto& p : v) { double x = U(rng), y = U(rng); p = {(float)x, (float)y, (float)(0.5 * std::sin(x) * std::cos(0.7 * y) + N(rng))}; }
  };
  Tree t;
  std::vector<P3> src;
  surf(t.pts);
  surf(src);
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (auto& p : t.pts) { const float v[3] = {p.x, p.y, p.z}; for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], v[d]); hi[d] = std::max(hi[d], v[d]); } }
  float ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
  // the GPU normalises by the 10-unit extent of the real cloud: keep the same cell size
  float scale = 2097152.f / 10.0f;
  (void)ext;
  {
    std::vector<int> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::vector<uint64_t> k(n);
    for (int i = 0; i < n; ++i) k[i] = morton(t.pts[i], lo, scale);
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return k[a] < k[b]; });
    std::vector<P3> sp(n);
    t.keys.resize(n);
    for (int i = 0; i < n; ++i) { sp[i] = t.pts[ord[i]]; t.keys[i] = k[ord[i]]; }
    t.pts.swap(sp);
  }
  Box rb;
  if (variant == 0) {
    int nl = (n + L - 1) / L;
    std::vector<uint64_t> lk(nl);
    for (int j = 0; j < nl; ++j) {
      int a = j * L, b = std::min(n, a + L);
      t.leaves.push_back(Leaf{a, b - a, box_of(t.pts, a, b)});
      lk[j] = t.keys[a];
    }
    t.root = build_blocks(t, lk, 0, nl, &rb);
  }
  else
    t.root = build_cells(t, 0, n, L, &rb);
  // leaf statistics
  double fill = 0, diag = 0;
  for (auto& l : t.leaves) {
    fill += l.count;
    diag += std::sqrt((l.box.hi[0] - l.box.lo[0]) * (l.box.hi[0] - l.box.lo[0]) + (l.box.hi[1] - l.box.lo[1]) * (l.box.hi[1] - l.box.lo[1]) +
                      (l.box.hi[2] - l.box.lo[2]) * (l.box.hi[2] - l.box.lo[2]));
  }
  std::printf("variant %d L %d: %zu leaves (avg fill %.2f, avg diag %.5f), %zu nodes\n", variant, L, t.leaves.size(), fill / t.leaves.size(),
              diag / t.leaves.size(), t.nodes.size());
  // queries in Hilbert / Morton order
  {
    std::vector<int> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::vector<uint64_t> k(n);
    for (int i = 0; i < n; ++i) k[i] = qorder ? hilbert(src[i], lo, scale) : morton(src[i], lo, scale);
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return k[a] < k[b]; });
    std::vector<P3> sp(n);
    for (int i = 0; i < n; ++i) sp[i] = src[ord[i]];
    src.swap(sp);
  }
  // emulate seeded packet traversal on a sample of packets
  const int npk = std::min(n / 32, 20000);
  const int stride = (n / 32) / npk;
  double tot_nodes = 0, tot_leaves = 0, tot_pts = 0, tot_single_nodes = 0, tot_single_leaves = 0;
  double w_nodes = 0, w_boxes = 0, w_leaves = 0;
  double ls_max_nodes = 0, ls_max_leaves = 0, ls_depth = 0;  // lock-step view of the 32 single walks of a warp
#pragma omp parallel for reduction(+ : tot_nodes, tot_leaves, tot_pts, tot_single_nodes, tot_single_leaves, w_nodes, w_boxes, w_leaves, ls_max_nodes, ls_max_leaves, ls_depth) schedule(dynamic, 64)
  for (int pk = 0; pk < npk; ++pk) {
    const P3* q = &src[(size_t)pk * stride * 32];
    float best[32];
    int lane_nodes[32], lane_leaves[32];
    // seed = exact NN distance (what the previous iteration's match provides once ICP has nearly converged):
    // obtained here by a plain single-query traversal
    for (int l = 0; l < 32; ++l) {
      float b = 1e30f;
      std::vector<int> st{t.root};
      while (!st.empty()) {
        int nd = st.back(); st.pop_back();
        if (nd < 0) { const Leaf& lf = t.leaves[~nd]; for (int i = 0; i < lf.count; ++i) b = std::min(b, pdist(q[l], t.pts[lf.first + i])); continue; }
        float d0 = bdist(q[l], t.nodes[nd].b[0]), d1 = bdist(q[l], t.nodes[nd].b[1]);
        int c0 = t.nodes[nd].child[0], c1 = t.nodes[nd].child[1];
        if (d1 < d0) { std::swap(d0, d1); std::swap(c0, c1); }
        if (d1 <= b) st.push_back(c1);
        if (d0 <= b) st.push_back(c0);
      }
      best[l] = b;
      // single-query seeded walk cost
      std::vector<std::pair<int, float>> s2{{t.root, 0.f}};
      lane_nodes[l] = lane_leaves[l] = 0;
      while (!s2.empty()) {
        auto [nd, dd] = s2.back(); s2.pop_back();
        if (dd > b) continue;
        if (nd < 0) { tot_single_leaves += 1; lane_leaves[l] += 1; continue; }
        tot_single_nodes += 1;
        lane_nodes[l] += 1;
        float d0 = bdist(q[l], t.nodes[nd].b[0]), d1 = bdist(q[l], t.nodes[nd].b[1]);
        if (d1 <= b) s2.push_back({t.nodes[nd].child[1], d1});
        if (d0 <= b) s2.push_back({t.nodes[nd].child[0], d0});
      }
    }
    {
      int mn = 0, ml = 0;
      for (int l = 0; l < 32; ++l) { mn = std::max(mn, lane_nodes[l]); ml = std::max(ml, lane_leaves[l]); }
      ls_max_nodes += mn; ls_max_leaves += ml;
      // depth of the query's own leaf = nodes on the direct path (lane 0)
      int d = 0, nd = t.root;
      while (nd >= 0) { ++d; float d0 = bdist(q[0], t.nodes[nd].b[0]), d1 = bdist(q[0], t.nodes[nd].b[1]); nd = d0 <= d1 ? t.nodes[nd].child[0] : t.nodes[nd].child[1]; }
      ls_depth += d;
    }
    // 4-wide packet walk: a wide node = a binary node with its internal children expanded one level
    {
      float b4[32];
      for (int l = 0; l < 32; ++l) b4[l] = best[l];
      std::vector<std::pair<int, float>> s4{{t.root, 0.f}};
      while (!s4.empty()) {
        auto e = s4.back(); s4.pop_back();
        float wmax = 0;
        for (int l = 0; l < 32; ++l) wmax = std::max(wmax, b4[l]);
        if (e.second > wmax) continue;
        if (e.first < 0) {
          const Leaf& lf = t.leaves[~e.first];
          w_leaves += 1;
          for (int l = 0; l < 32; ++l)
            for (int i = 0; i < lf.count; ++i) b4[l] = std::min(b4[l], pdist(q[l], t.pts[lf.first + i]));
          continue;
        }
        w_nodes += 1;
        // gather up to 4 children
        int ch[4]; Box bx[4]; int nc = 0;
        const Node& nd = t.nodes[e.first];
        for (int k = 0; k < 2; ++k) {
          int c = nd.child[k];
          if (c >= 0) { for (int k2 = 0; k2 < 2; ++k2) { ch[nc] = t.nodes[c].child[k2]; bx[nc] = t.nodes[c].b[k2]; ++nc; } }
          else { ch[nc] = c; bx[nc] = nd.b[k]; ++nc; }
        }
        w_boxes += nc;
        std::pair<float, int> want[4]; int nw = 0;
        for (int k = 0; k < nc; ++k) {
          float m = 1e30f;
          for (int l = 0; l < 32; ++l) { float d = bdist(q[l], bx[k]); if (d <= b4[l]) m = std::min(m, d); }
          if (m != 1e30f) want[nw++] = {m, ch[k]};
        }
        std::sort(want, want + nw, [](auto& a, auto& b) { return a.first > b.first; });
        for (int k = 0; k < nw; ++k) s4.push_back({want[k].second, want[k].first});
      }
    }
    // packet walk
    std::vector<std::pair<int, float>> st;
    int node = t.root;
    const int DONE = 0x7fffffff;
    auto pop = [&]() {
      float wmax = 0;
      for (int l = 0; l < 32; ++l) wmax = std::max(wmax, best[l]);
      node = DONE;
      while (!st.empty()) {
        auto e = st.back(); st.pop_back();
        if (e.second <= wmax) { node = e.first; break; }
      }
    };
    while (node != DONE) {
      if (node >= 0) {
        tot_nodes += 1;
        const Node& nd = t.nodes[node];
        float ml = 1e30f, mr = 1e30f;
        for (int l = 0; l < 32; ++l) {
          float dl = bdist(q[l], nd.b[0]), dr = bdist(q[l], nd.b[1]);
          if (dl <= best[l]) ml = std::min(ml, dl);
          if (dr <= best[l]) mr = std::min(mr, dr);
        }
        if (ml == 1e30f && mr == 1e30f) pop();
        else if (ml != 1e30f && mr != 1e30f) {
          bool lf = ml <= mr;
          st.push_back({lf ? nd.child[1] : nd.child[0], lf ? mr : ml});
          node = lf ? nd.child[0] : nd.child[1];
        }
        else node = ml != 1e30f ? nd.child[0] : nd.child[1];
      }
      else {
        const Leaf& lf = t.leaves[~node];
        tot_leaves += 1;
        tot_pts += lf.count;
        for (int l = 0; l < 32; ++l)
          for (int i = 0; i < lf.count; ++i) best[l] = std::min(best[l], pdist(q[l], t.pts[lf.first + i]));
        pop();
      }
    }
  }
  std::printf("packet walk: %.1f nodes, %.1f leaves, %.1f points per packet | single walk: %.1f nodes, %.1f leaves per query\n", tot_nodes / npk,
              tot_leaves / npk, tot_pts / npk, tot_single_nodes / npk / 32, tot_single_leaves / npk / 32);
  std::printf("lock-step view of 32 seeded single walks: slowest lane %.1f nodes / %.1f leaves vs mean %.1f / %.1f (lane utilisation if\n"
              "  all lanes ran to the slowest lane's count: nodes %.0f %%, leaves %.0f %%); depth of the direct path %.1f\n",
              ls_max_nodes / npk, ls_max_leaves / npk, tot_single_nodes / npk / 32, tot_single_leaves / npk / 32,
              100.0 * (tot_single_nodes / npk / 32) / (ls_max_nodes / npk), 100.0 * (tot_single_leaves / npk / 32) / (ls_max_leaves / npk),
              ls_depth / npk);
  std::printf("4-wide packet walk: %.1f wide nodes, %.1f box tests, %.1f leaves per packet (binary: %.1f box tests)\n", w_nodes / npk, w_boxes / npk,
              w_leaves / npk, 2 * tot_nodes / npk);
  return 0;
}
