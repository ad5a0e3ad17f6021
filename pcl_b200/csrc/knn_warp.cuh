// knn_warp.cuh — device code of the normals epilogue (eigen33 in fp32, the single-pass moments) and of the
// warp-per-query k-NN kernel.  A header so that search.cu's kernels share it and so that tests/host/knn_warp_host_test.cpp
// can compile the SAME source for the host (a lock-step emulation of one warp) and check it against brute force.
#pragma once
#include "internal.cuh"
#include "traverse.cuh"

namespace pclb200 {

// ---- normals ---------------------------------------------------------------------------------------
// pcl::eigen33 smallest eigenpair in fp32 — common/include/pcl/common/impl/eigen.hpp:52-133 (roots),
// :273-288 (largest cross product), :293-326 (eigen33); m = row-major symmetric 3x3.
__device__ __forceinline__ void roots2_dev(float b, float c, float* r)
{
  r[0] = 0.f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.f)
    d = 0.f;
  float sd = sqrtf(d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}

__device__ void roots3_dev(const float* m, float* r)
{
  float c0 = m[0] * m[4] * m[8] + 2.f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -
             m[8] * m[1] * m[1];
  float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  float c2 = m[0] + m[4] + m[8];
  if (fabsf(c0) < FLT_EPSILON) {
    roots2_dev(c2, c1, r);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = sqrtf(3.f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f)
    a_over_3 = 0.f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f)
    q = 0.f;
  float rho = sqrtf(-a_over_3);
  float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
  float cos_theta = cosf(theta), sin_theta = sinf(theta);
  r[0] = c2_over_3 + 2.f * rho * cos_theta;
  r[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  r[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  float t;
  if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  if (r[1] >= r[2]) {
    t = r[1]; r[1] = r[2]; r[2] = t;
    if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  }
  if (r[0] <= 0.f)
    roots2_dev(c2, c1, r);
}

__device__ void largest_eigvec_dev(const float* s, float* v)
{
  float c[3][3] = {{s[1] * s[5] - s[2] * s[4], s[2] * s[3] - s[0] * s[5], s[0] * s[4] - s[1] * s[3]},
                   {s[1] * s[8] - s[2] * s[7], s[2] * s[6] - s[0] * s[8], s[0] * s[7] - s[1] * s[6]},
                   {s[4] * s[8] - s[5] * s[7], s[5] * s[6] - s[3] * s[8], s[3] * s[7] - s[4] * s[6]}};
  float len[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    len[i] = sqrtf(c[i][0] * c[i][0] + c[i][1] * c[i][1] + c[i][2] * c[i][2]);
  int idx = 0;
  if (len[1] > len[idx]) idx = 1;
  if (len[2] > len[idx]) idx = 2;
  float l = idx == 0 ? len[0] : (idx == 1 ? len[1] : len[2]);
#pragma unroll
  for (int d = 0; d < 3; ++d)
    v[d] = (idx == 0 ? c[0][d] : (idx == 1 ? c[1][d] : c[2][d])) / l;
}

__device__ void eigen33_smallest_dev(const float* mat, float& eigenvalue, float* ev)
{
  float scale = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i)
    scale = fmaxf(scale, fabsf(mat[i]));
  if (scale <= FLT_MIN)
    scale = 1.f;
  float s[9];
#pragma unroll
  for (int i = 0; i < 9; ++i)
    s[i] = __fdiv_rn(mat[i], scale);
  float r[3];
  roots3_dev(s, r);
  eigenvalue = r[0] * scale;
  if ((r[1] - r[0]) > FLT_EPSILON) {
    s[0] -= r[0]; s[4] -= r[0]; s[8] -= r[0];
    largest_eigvec_dev(s, ev);
  }
  else if ((r[2] - r[0]) > FLT_EPSILON) {
    s[0] -= r[2]; s[4] -= r[2]; s[8] -= r[2];
    float v[3];
    largest_eigvec_dev(s, v);
    // Eigen unitOrthogonal()
    bool a = fabsf(v[0]) <= fabsf(v[2]) * FLT_EPSILON, b = fabsf(v[1]) <= fabsf(v[2]) * FLT_EPSILON;
    if (!a || !b) {
      float inv = 1.f / sqrtf(v[0] * v[0] + v[1] * v[1]);
      ev[0] = -v[1] * inv; ev[1] = v[0] * inv; ev[2] = 0.f;
    }
    else {
      float inv = 1.f / sqrtf(v[1] * v[1] + v[2] * v[2]);
      ev[0] = 0.f; ev[1] = -v[2] * inv; ev[2] = v[1] * inv;
    }
  }
  else {
    ev[0] = 1.f; ev[1] = 0.f; ev[2] = 0.f;
  }
}

// shifted single-pass moments (centroid.hpp:605-640): one neighbour, K = the first neighbour of the list
__device__ __forceinline__ void moments_add(float (&accu)[9], float Kx, float Ky, float Kz, const float4 p)
{
  const float x = __fsub_rn(p.x, Kx), y = __fsub_rn(p.y, Ky), z = __fsub_rn(p.z, Kz);
  accu[0] = __fadd_rn(accu[0], __fmul_rn(x, x));
  accu[1] = __fadd_rn(accu[1], __fmul_rn(x, y));
  accu[2] = __fadd_rn(accu[2], __fmul_rn(x, z));
  accu[3] = __fadd_rn(accu[3], __fmul_rn(y, y));
  accu[4] = __fadd_rn(accu[4], __fmul_rn(y, z));
  accu[5] = __fadd_rn(accu[5], __fmul_rn(z, z));
  accu[6] = __fadd_rn(accu[6], x);
  accu[7] = __fadd_rn(accu[7], y);
  accu[8] = __fadd_rn(accu[8], z);
}

// moments -> covariance (centroid.hpp:641-651) -> solvePlaneParameters (feature.hpp:65-92) ->
// flipNormalTowardsViewpoint (normal_3d.h:169-188); returns {nx, ny, nz, curvature}
__device__ __forceinline__ float4 normal_from_moments(float (&accu)[9], int cnt, const float4 qq, float vpx, float vpy,
                                                      float vpz, int* __restrict__ not_dense)
{
  const float fc = (float)cnt;
#pragma unroll
  for (int t = 0; t < 9; ++t)
    accu[t] = __fdiv_rn(accu[t], fc);
  float cov[9];
  cov[0] = __fsub_rn(accu[0], __fmul_rn(accu[6], accu[6]));
  cov[1] = __fsub_rn(accu[1], __fmul_rn(accu[6], accu[7]));
  cov[2] = __fsub_rn(accu[2], __fmul_rn(accu[6], accu[8]));
  cov[4] = __fsub_rn(accu[3], __fmul_rn(accu[7], accu[7]));
  cov[5] = __fsub_rn(accu[4], __fmul_rn(accu[7], accu[8]));
  cov[8] = __fsub_rn(accu[5], __fmul_rn(accu[8], accu[8]));
  cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
  float ev, n[3];
  eigen33_smallest_dev(cov, ev, n);
  const float eig_sum = __fadd_rn(__fadd_rn(cov[0], cov[4]), cov[8]);
  const float curv = eig_sum != 0.f ? fabsf(__fdiv_rn(ev, eig_sum)) : 0.f;
  const float vx = vpx - qq.x, vy = vpy - qq.y, vz = vpz - qq.z;
  const float cos_theta = vx * n[0] + vy * n[1] + vz * n[2];
  if (cos_theta < 0.f) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  if (!(isfinite(n[0]) && isfinite(n[1]) && isfinite(n[2]) && isfinite(curv)))
    *not_dense = 1;
  return make_float4(n[0], n[1], n[2], curv);
}

// =============================================================================================================
// Warp-cooperative k-NN (k <= 32): one WARP per query, brute force over the cells that hold the answer
// =============================================================================================================
// The per-thread kernels above keep a k-deep sorted list in registers: at k = 16 every accepted candidate costs a
// ~100-instruction dependent bubble pass executed under divergence (ncu, profiles/r2i: 914 warp-instructions per query,
// 10 of 32 lanes active, 19 % issue utilisation, 42 ms for 10 M normals).  Here the list is ONE ENTRY PER LANE, sorted
// across the warp, and candidates arrive 32 at a time from CONTIGUOUS memory:
//   * the cell table (traverse.cuh) maps a cell of level b to the subtree holding exactly its points; a subtree's leaves
//     are consecutive in the Morton array, so "all points of a cell" is one coalesced range;
//   * a ball of radius R = half a level-b cell reaches at most 2 x 2 x 2 cells.  The warp gathers those cells, keeps the k
//     smallest (d2, index) — a bitonic sort / merge by shuffles when many candidates beat the current k-th, a ranked
//     insertion (ballot + shuffle-up) when few do — and the result is EXACT iff the k-th distance is below R: every
//     point outside the gathered cells lies outside [q - R, q + R]^3.  Otherwise the next coarser level (R doubles).
//   * the start level comes from the index's density (cells that hold ~2k points on average), so one attempt is the norm.
// Queries the scheme does not fit (far outside the cloud, > kWarpKnnMaxLeaves leaves in reach, no level certifies) are
// flagged and redone by the per-thread kernel — same results, the exact walk is the fallback, never an approximation.
constexpr int kWarpKnnMaxLeaves = 1024;

__device__ __forceinline__ bool lex_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

// ascending bitonic sort of one (d, i, p) triple per lane
__device__ __forceinline__ void warp_sort32(float& d, int& i, int& p, int lane)
{
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const float pd = __shfl_xor_sync(full, d, j);
      const int pi = __shfl_xor_sync(full, i, j);
      const int pp = __shfl_xor_sync(full, p, j);
      const bool asc = (lane & k2) == 0, lower = (lane & j) == 0;
      const bool mine_first = lex_less(d, i, pd, pi);
      const bool keep = (lower == asc) ? mine_first : !mine_first;
      if (!keep) {
        d = pd; i = pi; p = pp;
      }
    }
}

// list (sorted ascending across lanes) <- the 32 smallest of list U cand (cand sorted ascending across lanes)
__device__ __forceinline__ void warp_merge32(float& ld, int& li, int& lp, float cd, int ci, int cp, int lane)
{
  const unsigned full = 0xffffffffu;
  const float rd = __shfl_sync(full, cd, 31 - lane);
  const int ri = __shfl_sync(full, ci, 31 - lane);
  const int rp = __shfl_sync(full, cp, 31 - lane);
  if (lex_less(rd, ri, ld, li)) {  // elementwise min of an ascending and a descending sequence: bitonic, the 32 smallest
    ld = rd; li = ri; lp = rp;
  }
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const float pd = __shfl_xor_sync(full, ld, j);
    const int pi = __shfl_xor_sync(full, li, j);
    const int pp = __shfl_xor_sync(full, lp, j);
    const bool lower = (lane & j) == 0;
    const bool mine_first = lex_less(ld, li, pd, pi);
    if (lower != mine_first) {
      ld = pd; li = pi; lp = pp;
    }
  }
}

template <bool NORMALS>
__global__ void __launch_bounds__(256)
k_knn_warp(const TreeView T, const int2* __restrict__ node_leaves, int b_start, float r_first, const float4* __restrict__ q, size_t nq,
           int k, int32_t* __restrict__ out_idx, float* __restrict__ out_d2, float vpx, float vpy, float vpz,
           float4* __restrict__ out_n, int* __restrict__ not_dense, unsigned char* __restrict__ redo)
{
  __shared__ int s_pos[NORMALS ? 8 : 1][NORMALS ? 32 : 1][NORMALS ? 33 : 1];  // per warp: 32 queries x k neighbour positions (+1: no bank conflicts)
  __shared__ float s_cd[8][64];  // per warp: buffered candidates (d2, original index, Morton position)
  __shared__ int s_ci[8][64];
  __shared__ int s_cp[8][64];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const CellTable& C = T.cells;
  const float inf = __int_as_float(0x7f800000);
  const float qnan = __int_as_float(0x7fc00000);
  const size_t n_warps = (size_t)gridDim.x * (blockDim.x >> 5);
  const size_t n_batches = (nq + 31) / 32;
  for (size_t batch = (size_t)blockIdx.x * (blockDim.x >> 5) + warp; batch < n_batches; batch += n_warps) {
    int my_state = 0;  // epilogue (NORMALS): state of query batch*32 + lane: 0 = none, 1 = list ready, 2 = NaN row, 3 = redo
    for (int t = 0; t < 32; ++t) {
      const size_t qi = batch * 32 + t;
      if (qi >= nq)
        break;
      const float4 qq = __ldg(q + qi);
      const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
      if (!(isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z))) {
        if (NORMALS) {
          if (lane == t)
            my_state = 2;
        }
        else if (lane < k) {
          out_idx[slot * k + lane] = -1;
          out_d2[slot * k + lane] = inf;
        }
        continue;
      }
      const unsigned cqx = morton_cell(qq.x, C.lo[0], C.scale), cqy = morton_cell(qq.y, C.lo[1], C.scale),
                     cqz = morton_cell(qq.z, C.lo[2], C.scale);
      float ld = inf;
      int li = kSentinelIndex, lp = -1;
      bool done = false;
      // attempts: [a ball sized for ~1.4 k points of this cloud's density, when that is well inside half a cell of the
      // start level,] then half a cell of the start level, then of every coarser level
      for (int attempt = 0; !done; ++attempt) {
        const bool sized = r_first > 0.f && attempt == 0;
        const int b = b_start - (r_first > 0.f ? max(attempt - 1, 0) : attempt);
        if (b < 1)
          break;
        const int s = 21 - b;
        const float R = sized ? r_first : __fmul_rd(__fmul_rd(0.5f * (float)(1u << s), C.inv_scale), 0.999f);
        const unsigned ax = morton_cell(__fsub_rd(qq.x, R), C.lo[0], C.scale), bx = morton_cell(__fadd_ru(qq.x, R), C.lo[0], C.scale);
        const unsigned ay = morton_cell(__fsub_rd(qq.y, R), C.lo[1], C.scale), by = morton_cell(__fadd_ru(qq.y, R), C.lo[1], C.scale);
        const unsigned az = morton_cell(__fsub_rd(qq.z, R), C.lo[2], C.scale), bz = morton_cell(__fadd_ru(qq.z, R), C.lo[2], C.scale);
        if ((bx >> s) - (ax >> s) > 1u || (by >> s) - (ay >> s) > 1u || (bz >> s) - (az >> s) > 1u)
          continue;  // (rounding at a cell edge) the box needs the next coarser level
        const unsigned hx = cqx >> s, hy = cqy >> s, hz = cqz >> s;
        const unsigned ox = (ax >> s) + (bx >> s) - hx, oy = (ay >> s) + (by >> s) - hy, oz = (az >> s) + (bz >> s) - hz;
        const unsigned E = (ox != hx ? 1u : 0u) | (oy != hy ? 2u : 0u) | (oz != hz ? 4u : 0u);
        // lanes 0..7 look one cell up each; a leaf that spans several cells comes back several times: keep one
        int ref = kDone;
        if (lane < 8 && ((unsigned)lane & ~E) == 0u)
          ref = cell_lookup(C, cell_key(b, (lane & 1) ? ox : hx, (lane & 2) ? oy : hy, (lane & 4) ? oz : hz));
        const unsigned grp = __match_any_sync(full, ref != kDone ? ref : (int)(0x40000000 | lane));
        // a leaf shared by several cells is kept once, by the lowest lane, whose cell need not be the nearest of them
        // (cells 3 and 5 share a leaf, cell 1 is empty): such an entry is never pruned by a cell bound
        const bool shared = ref != kDone && (grp & (grp - 1u)) != 0u;
        if (ref != kDone && (grp & lt) != 0u)
          ref = kDone;
        int first = 0, cnt = 0;
        if (ref != kDone) {
          if (ref < 0) {
            first = ~ref;
            cnt = 1;
          }
          else {
            const int2 r = __ldg(node_leaves + ref);
            first = r.x;
            cnt = r.y;
          }
        }
        int tot = cnt;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
          tot += __shfl_xor_sync(full, tot, o);
        tot = __shfl_sync(full, tot, 0);
        if (tot > kWarpKnnMaxLeaves)
          break;  // a dense knot (duplicates): the exact walk prunes it, a brute-force gather would not
        ld = inf;
        li = kSentinelIndex;
        lp = -1;
        // Candidates that beat the current k-th are only APPENDED to a per-warp buffer; the list is updated (sort + merge,
        // or a few ranked insertions) when 32 have collected and at the end of the home cell, so the expensive network
        // runs once per ~32 survivors instead of once per round.  A stale threshold only admits extra candidates.
        // the running threshold starts at the certification radius: a candidate beyond it cannot be part of a certified
        // answer, so only the points inside the ball (~k..2k of the few hundred gathered) ever reach the sorting network
        const float R2cert = __fmul_rd(__fmul_rd(R, R), 0.999998f);
        float Tk = R2cert;
        int Ti = kSentinelIndex;
        int nbuf = 0;
        bool have_list = false;
        auto flush = [&](int take) {  // fold the first `take` (<= 32) buffered candidates into the list
          float cd = inf;
          int ci = kSentinelIndex, cp = -1;
          if (lane < take) {
            cd = s_cd[warp][lane];
            ci = s_ci[warp][lane];
            cp = s_cp[warp][lane];
          }
          if (!have_list || take > 12) {
            warp_sort32(cd, ci, cp, lane);
            if (have_list)
              warp_merge32(ld, li, lp, cd, ci, cp, lane);
            else {  // nothing to merge with yet: the sorted candidates are the list
              ld = cd; li = ci; lp = cp;
              have_list = true;
            }
          }
          else {
            for (int src = 0; src < take; ++src) {
              const float xd = __shfl_sync(full, cd, src);
              const int xi = __shfl_sync(full, ci, src), xp = __shfl_sync(full, cp, src);
              const int rank = __popc(__ballot_sync(full, lex_less(ld, li, xd, xi)));  // entries that stay in front
              const float ud = __shfl_up_sync(full, ld, 1);
              const int ui = __shfl_up_sync(full, li, 1), up = __shfl_up_sync(full, lp, 1);
              if (lane > rank) {
                ld = ud; li = ui; lp = up;
              }
              else if (lane == rank) {
                ld = xd; li = xi; lp = xp;
              }
            }
          }
          __syncwarp();
          // keep what is left of the buffer (at most 31 entries) at its front
          const int rest = nbuf - take;
          float md = 0.f;
          int mi = 0, mp = 0;
          if (lane < rest) {
            md = s_cd[warp][take + lane];
            mi = s_ci[warp][take + lane];
            mp = s_cp[warp][take + lane];
          }
          __syncwarp();
          if (lane < rest) {
            s_cd[warp][lane] = md;
            s_ci[warp][lane] = mi;
            s_cp[warp][lane] = mp;
          }
          nbuf = rest;
          {
            const float nk = __shfl_sync(full, ld, k - 1);
            const int ni = __shfl_sync(full, li, k - 1);
            if (lex_less(nk, ni, Tk, Ti)) {  // never looser than the certification radius
              Tk = nk;
              Ti = ni;
            }
          }
          __syncwarp();
        };
        const float gx2 = (E & 1u) ? cell_gap2(C, 0, qq.x, hx, ox, s) : 0.f;
        const float gy2 = (E & 2u) ? cell_gap2(C, 1, qq.y, hy, oy, s) : 0.f;
        const float gz2 = (E & 4u) ? cell_gap2(C, 2, qq.z, hz, oz, s) : 0.f;
        for (int c = 0; c < 8; ++c) {
          const int f = __shfl_sync(full, first, c), n = __shfl_sync(full, cnt, c);
          if (n == 0)
            continue;
          // every point of cell c is at least this far (traverse.cuh: cell_gap2): a cell the k-th already beats is skipped
          const float bound = __shfl_sync(full, shared ? 1 : 0, c)
                                  ? 0.f
                                  : __fadd_rd(__fadd_rd((c & 1) ? gx2 : 0.f, (c & 2) ? gy2 : 0.f), (c & 4) ? gz2 : 0.f);
          if (!(bound <= Tk))
            continue;
          const int end = (f + n) * kLeafSize;
          const float4 pad = make_float4(inf, inf, inf, __int_as_float(kSentinelIndex));
          float4 pnext = f * kLeafSize + lane < end ? ldg4(T.pts + f * kLeafSize + lane) : pad;
          for (int base = f * kLeafSize; base < end; base += 32) {
            const int sidx = base + lane;
            const float4 p = pnext;
            if (base + 32 < end)  // the next round's line is in flight while this one is folded
              pnext = sidx + 32 < end ? ldg4(T.pts + sidx + 32) : pad;
            const float d = dist2_rn(qq.x, qq.y, qq.z, p.x, p.y, p.z);  // +inf for padding slots
            const int oi = __float_as_int(p.w);
            const bool pass = d < inf && lex_less(d, oi, Tk, Ti);
            const unsigned pm = __ballot_sync(full, pass);
            if (!pm)
              continue;
            if (pass) {
              const int at = nbuf + __popc(pm & lt);
              s_cd[warp][at] = d;
              s_ci[warp][at] = oi;
              s_cp[warp][at] = sidx;
            }
            nbuf += __popc(pm);
            __syncwarp();
            if (nbuf >= 32)
              flush(32);
          }
        }
        if (nbuf > 0)
          flush(nbuf);
        // exact iff the k-th neighbour lies strictly inside the gathered box (margin >> fp32 rounding of d2)
        const float dk = __shfl_sync(full, ld, k - 1);
        done = dk < R2cert;
      }
      if (!done) {
        redo[qi] = 1;  // (all lanes store the same byte)
        if (NORMALS && lane == t)
          my_state = 3;
        continue;
      }
      if (NORMALS) {
        if (lane < k)
          s_pos[warp][t][lane] = lp;
        if (lane == t)
          my_state = 1;
      }
      else if (lane < k) {
        out_idx[slot * k + lane] = li;
        out_d2[slot * k + lane] = ld;
      }
    }
    if (NORMALS) {
      // epilogue: lane t folds query t's neighbours sequentially, in list order — the arithmetic of k_normals
      __syncwarp();
      const size_t qi = batch * 32 + lane;
      if (qi < nq && my_state != 0 && my_state != 3) {
        const float4 qq = __ldg(q + qi);
        const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
        if (my_state == 2) {
          out_n[slot] = make_float4(qnan, qnan, qnan, qnan);
          *not_dense = 1;
        }
        else {
          float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          float Kx = 0.f, Ky = 0.f, Kz = 0.f;
          for (int j = 0; j < k; ++j) {
            const float4 p = ldg4(T.pts + s_pos[warp][lane][j]);
            if (j == 0) { Kx = p.x; Ky = p.y; Kz = p.z; }
            moments_add(accu, Kx, Ky, Kz, p);
          }
          out_n[slot] = normal_from_moments(accu, k, qq, vpx, vpy, vpz, not_dense);
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace pclb200
