// search.cu — batch k-NN, radius search and k-NN normals over the LBVH.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch / radiusSearch (kdtree/include/pcl/kdtree/impl/
// kdtree_flann.hpp:234-274, 372-414), the OpenMP batch loops of pcl::search::Search
// (search/include/pcl/search/impl/search.hpp:111-194) and NormalEstimation::computeFeature
// (features/include/pcl/features/impl/normal_3d.hpp:47-96).
// One query per thread, queries in Morton order so a warp walks one neighbourhood of the tree
// (node lines are shared through L1/L2); candidate lists live in registers (k <= 32).
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>

#include "internal.cuh"
#include "traverse.cuh"
#include "knn_warp.cuh"

namespace pclb200 {

static inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

// ---- k nearest, k <= K (compile-time), ascending (d2, original index) ------------------------------
template <int K>
struct NearestK {
  float qx, qy, qz;
  const float4* pts;
  float d[K];
  int pos[K];
  __device__ __forceinline__ void init(float bound)
  {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      d[j] = bound;
      pos[j] = -1;
    }
  }
  __device__ __forceinline__ int orig(int p) const
  {
    return p < 0 ? kSentinelIndex : __float_as_int(__ldg(&pts[p].w));
  }
  __device__ __forceinline__ float bound() const { return d[K - 1]; }
  __device__ __forceinline__ void prune(float) {}
  __device__ __forceinline__ void leaf(const float4* lp, int first_pos)
  {
    float4 p[kLeafSize];
#pragma unroll
    for (int j = 0; j < kLeafSize; ++j)
      p[j] = ldg4(lp + j);
#pragma unroll
    for (int j = 0; j < kLeafSize; ++j) {
      float dd = dist2_rn(qx, qy, qz, p[j].x, p[j].y, p[j].z);
      int oi = __float_as_int(p[j].w);
      if (dd < d[K - 1] || (dd == d[K - 1] && oi < orig(pos[K - 1]))) {
        d[K - 1] = dd;
        pos[K - 1] = first_pos + j;
        // one bubble pass keeps the list sorted; exact-distance ties compare original indices
#pragma unroll
        for (int t = K - 1; t > 0; --t) {
          bool lt = d[t] < d[t - 1] || (d[t] == d[t - 1] && orig(pos[t]) < orig(pos[t - 1]));
          if (lt) {
            float td = d[t]; d[t] = d[t - 1]; d[t - 1] = td;
            int tp = pos[t]; pos[t] = pos[t - 1]; pos[t - 1] = tp;
          }
        }
      }
    }
  }
};

template <int K>
__global__ void __launch_bounds__(128)
k_knn(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts, int root,
      const float4* __restrict__ q, size_t nq, int k_out, float init_bound,
      int32_t* __restrict__ out_idx, float* __restrict__ out_d2, int* __restrict__ d_error,
      const unsigned char* __restrict__ only = nullptr)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq || (only && !only[i]))  // fix-up pass: only the queries the warp kernel handed back
    return;
  const float4 qq = __ldg(q + i);
  const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
  NearestK<K> v;
  v.qx = qq.x; v.qy = qq.y; v.qz = qq.z;
  v.pts = pts;
  v.init(init_bound);
  // a non-finite query has no neighbours (every comparison with NaN fails; rows stay (-1, +inf)) — without the guard
  // it would walk the whole tree, because a NaN box bound never prunes
  if (isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z) && !traverse(nodes, pts, root, qq.x, qq.y, qq.z, v))
    atomicExch(d_error, 1);
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (j < k_out) {
      const bool have = v.pos[j] >= 0;
      out_idx[slot * k_out + j] = have ? v.orig(v.pos[j]) : -1;
      out_d2[slot * k_out + j] = have ? v.d[j] : __int_as_float(0x7f800000);
    }
}

// ---- any k: the candidate list lives in the output rows themselves (global memory) -----------------
struct NearestAny {
  float qx, qy, qz;
  const float4* pts;
  float* d;   // k entries, ascending
  int* pos;   // k entries (positions; converted to original indices afterwards)
  int k;
  __device__ __forceinline__ int orig(int p) const
  {
    return p < 0 ? kSentinelIndex : __float_as_int(__ldg(&pts[p].w));
  }
  __device__ __forceinline__ float bound() const { return d[k - 1]; }
  __device__ __forceinline__ void prune(float) {}
  __device__ __forceinline__ void leaf(const float4* lp, int first_pos)
  {
    for (int j = 0; j < kLeafSize; ++j) {
      float4 p = ldg4(lp + j);
      float dd = dist2_rn(qx, qy, qz, p.x, p.y, p.z);
      int oi = __float_as_int(p.w);
      if (dd < d[k - 1] || (dd == d[k - 1] && oi < orig(pos[k - 1]))) {
        int t = k - 1;
        while (t > 0 && (dd < d[t - 1] || (dd == d[t - 1] && oi < orig(pos[t - 1])))) {
          d[t] = d[t - 1];
          pos[t] = pos[t - 1];
          --t;
        }
        d[t] = dd;
        pos[t] = first_pos + j;
      }
    }
  }
};

__global__ void __launch_bounds__(128)
k_knn_any(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts, int root,
          const float4* __restrict__ q, size_t nq, int k, float init_bound, int32_t* out_idx, float* out_d2,
          int* __restrict__ d_error)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq)
    return;
  const float4 qq = __ldg(q + i);
  const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
  NearestAny v;
  v.qx = qq.x; v.qy = qq.y; v.qz = qq.z;
  v.pts = pts;
  v.k = k;
  v.d = out_d2 + slot * k;
  v.pos = out_idx + slot * k;
  for (int j = 0; j < k; ++j) {
    v.d[j] = init_bound;
    v.pos[j] = -1;
  }
  if (isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z) && !traverse(nodes, pts, root, qq.x, qq.y, qq.z, v))
    atomicExch(d_error, 1);
  for (int j = 0; j < k; ++j) {
    int p = v.pos[j];
    v.pos[j] = p >= 0 ? v.orig(p) : -1;
    if (p < 0)
      v.d[j] = __int_as_float(0x7f800000);
  }
}

// warp-cooperative path (defined below): returns the per-query "redo" flags of the queries it handed back, or an empty
// buffer when it does not apply (no cell table, k outside 8..32)
static bool warp_knn_lists(Ctx& c, const Index& idx, const float4* d_q, size_t nq, int k, int32_t* d_out_idx,
                           float* d_out_d2, DevBuf<unsigned char>& redo);

void launch_knn(Ctx& c, const Index& idx, const float4* d_q, size_t nq, int k, float init_bound, int32_t* d_out_idx,
                float* d_out_d2)
{
  if (nq == 0 || k <= 0)
    return;
  cudaStream_t s = c.stream;
  const unsigned g = grid_for(nq, 128);
  DevBuf<unsigned char> redo;
  const unsigned char* only = nullptr;
  if (init_bound == std::numeric_limits<float>::infinity() && warp_knn_lists(c, idx, d_q, nq, k, d_out_idx, d_out_d2, redo))
    only = redo.p;  // the per-thread kernel below only redoes what the warp kernel handed back
#define PCLB_KNN_CASE(KK)                                                                                            \
  k_knn<KK><<<g, 128, 0, s>>>(idx.nodes.p, idx.pts.p, idx.root, d_q, nq, k, init_bound, d_out_idx, d_out_d2, c.d_error, only)
  if (k == 1) PCLB_KNN_CASE(1);
  else if (k == 2) PCLB_KNN_CASE(2);
  else if (k <= 4) PCLB_KNN_CASE(4);
  else if (k <= 8) PCLB_KNN_CASE(8);
  else if (k <= 10) PCLB_KNN_CASE(10);
  else if (k <= 16) PCLB_KNN_CASE(16);
  else if (k <= 20) PCLB_KNN_CASE(20);
  else if (k <= 32) PCLB_KNN_CASE(32);
  else
    k_knn_any<<<g, 128, 0, s>>>(idx.nodes.p, idx.pts.p, idx.root, d_q, nq, k, init_bound, d_out_idx, d_out_d2,
                                c.d_error);
#undef PCLB_KNN_CASE
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

// ---- per-query statistics of the k nearest neighbours (outlier filters) -------------------------------------------
// mean[slot] = float( sum_{j=1..k'-1} sqrt(double(d2_j)) / (k'-1) )   (statistical_outlier_removal.hpp:88-97; j = 0 is
//              the query itself when it belongs to the cloud), 0 for non-finite queries
// kth[slot]  = d2 of neighbour k-1, +inf when fewer than k points are indexed (radius_outlier_removal.hpp:86-118)
template <int K>
__global__ void __launch_bounds__(128)
k_knn_stats(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts, int root, const float4* __restrict__ q,
            size_t nq, int k, float* __restrict__ out_mean, float* __restrict__ out_kth, int* __restrict__ d_error)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq)
    return;
  const float4 qq = __ldg(q + i);
  const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
  float mean = 0.f, kth = __int_as_float(0x7f800000);
  if (isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z)) {
    NearestK<K> v;
    v.qx = qq.x; v.qy = qq.y; v.qz = qq.z;
    v.pts = pts;
    v.init(__int_as_float(0x7f800000));
    if (!traverse(nodes, pts, root, qq.x, qq.y, qq.z, v))
      atomicExch(d_error, 1);
    double sum = 0.0;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (j < k && v.pos[j] >= 0) {
        if (j >= 1)
          sum += sqrt((double)v.d[j]);
        ++cnt;
        if (j == k - 1)
          kth = v.d[j];
      }
    if (cnt > 1)
      mean = (float)(sum / (double)(cnt - 1));
  }
  if (out_mean)
    out_mean[slot] = mean;
  if (out_kth)
    out_kth[slot] = kth;
}

// same from materialised lists (k > 32): rows of pitch k by slot
__global__ void k_stats_from_lists(const float4* __restrict__ q, size_t nq, int k, const int32_t* __restrict__ li,
                                   const float* __restrict__ ld, float* __restrict__ out_mean, float* __restrict__ out_kth)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq)
    return;
  const float4 qq = __ldg(q + i);
  const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
  float mean = 0.f, kth = __int_as_float(0x7f800000);
  if (isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z)) {
    double sum = 0.0;
    int cnt = 0;
    for (int j = 0; j < k; ++j)
      if (li[slot * k + j] >= 0) {
        if (j >= 1)
          sum += sqrt((double)ld[slot * k + j]);
        ++cnt;
        if (j == k - 1)
          kth = ld[slot * k + j];
      }
    if (cnt > 1)
      mean = (float)(sum / (double)(cnt - 1));
  }
  if (out_mean)
    out_mean[slot] = mean;
  if (out_kth)
    out_kth[slot] = kth;
}

void launch_knn_stats(Ctx& c, const Index& idx, const float4* d_q, size_t nq, int k, float* d_mean, float* d_kth)
{
  if (!nq || k <= 0)
    return;
  cudaStream_t s = c.stream;
  const unsigned g = grid_for(nq, 128);
  if (k > 32) {
    // large neighbourhoods (StatisticalOutlierRemoval's usual mean_k = 50): exact lists first, folded afterwards.
    // NaN queries get empty rows from the walk (no box is ever within a NaN bound) and are zeroed by the fold.
    DevBuf<int32_t> li;
    DevBuf<float> ld;
    li.alloc(nq * (size_t)k, s);
    ld.alloc(nq * (size_t)k, s);
    launch_knn(c, idx, d_q, nq, k, __builtin_inff(), li.p, ld.p);
    k_stats_from_lists<<<grid_for(nq, 256), 256, 0, s>>>(d_q, nq, k, li.p, ld.p, d_mean, d_kth);
    ++c.launches;
    PCLB_CUDA(cudaGetLastError());
    return;
  }
#define PCLB_STAT_CASE(KK) \
  k_knn_stats<KK><<<g, 128, 0, s>>>(idx.nodes.p, idx.pts.p, idx.root, d_q, nq, k, d_mean, d_kth, c.d_error)
  if (k <= 2) PCLB_STAT_CASE(2);
  else if (k <= 4) PCLB_STAT_CASE(4);
  else if (k <= 8) PCLB_STAT_CASE(8);
  else if (k <= 10) PCLB_STAT_CASE(10);
  else if (k <= 16) PCLB_STAT_CASE(16);
  else if (k <= 20) PCLB_STAT_CASE(20);
  else PCLB_STAT_CASE(32);
#undef PCLB_STAT_CASE
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

// ---- radius search: count, then fill keys (d2 bits << 32 | original index) -------------------------
struct RadiusCount {
  float qx, qy, qz, r2;
  float r2_below;  // largest float < r2: subtrees with bound > r2_below hold no d2 < r2
  unsigned long long n;
  __device__ __forceinline__ float bound() const { return r2_below; }
  __device__ __forceinline__ void prune(float) {}
  __device__ __forceinline__ void leaf(const float4* lp, int)
  {
#pragma unroll
    for (int j = 0; j < kLeafSize; ++j) {
      float4 p = ldg4(lp + j);
      if (dist2_rn(qx, qy, qz, p.x, p.y, p.z) < r2)
        ++n;
    }
  }
};

struct RadiusFill {
  float qx, qy, qz, r2;
  float r2_below;
  unsigned long long* out;
  __device__ __forceinline__ float bound() const { return r2_below; }
  __device__ __forceinline__ void prune(float) {}
  __device__ __forceinline__ void leaf(const float4* lp, int)
  {
#pragma unroll
    for (int j = 0; j < kLeafSize; ++j) {
      float4 p = ldg4(lp + j);
      float dd = dist2_rn(qx, qy, qz, p.x, p.y, p.z);
      if (dd < r2)
        *out++ = ((unsigned long long)__float_as_uint(dd) << 32) | (unsigned)__float_as_int(p.w);
    }
  }
};

__global__ void __launch_bounds__(128)
k_radius_count(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts, int root,
               const float4* __restrict__ q, size_t nq, float r2, float r2_below,
               unsigned long long* __restrict__ counts, int* __restrict__ d_error)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq)
    return;
  const float4 qq = __ldg(q + i);
  RadiusCount v{qq.x, qq.y, qq.z, r2, r2_below, 0ULL};
  // a non-finite query has no neighbour (every comparison with NaN fails) — and would otherwise walk the whole tree
  if (isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z) && !traverse(nodes, pts, root, qq.x, qq.y, qq.z, v))
    atomicExch(d_error, 1);
  counts[(size_t)(unsigned)__float_as_int(qq.w)] = v.n;
}

__global__ void __launch_bounds__(128)
k_radius_fill(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts, int root,
              const float4* __restrict__ q, size_t nq, float r2, float r2_below,
              const unsigned long long* __restrict__ offsets, unsigned long long* __restrict__ keys,
              int* __restrict__ d_error)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq)
    return;
  const float4 qq = __ldg(q + i);
  RadiusFill v{qq.x, qq.y, qq.z, r2, r2_below, keys + offsets[(size_t)(unsigned)__float_as_int(qq.w)]};
  if (isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z) && !traverse(nodes, pts, root, qq.x, qq.y, qq.z, v))
    atomicExch(d_error, 1);
}

static float float_below(float x) { return std::nextafter(x, -INFINITY); }

void launch_radius_count(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2,
                         unsigned long long* d_counts)
{
  if (!nq)
    return;
  k_radius_count<<<grid_for(nq, 128), 128, 0, c.stream>>>(idx.nodes.p, idx.pts.p, idx.root, d_q, nq, r2,
                                                         float_below(r2), d_counts, c.d_error);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

void launch_radius_fill(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2,
                        const unsigned long long* d_offsets, unsigned long long* d_keys)
{
  if (!nq)
    return;
  k_radius_fill<<<grid_for(nq, 128), 128, 0, c.stream>>>(idx.nodes.p, idx.pts.p, idx.root, d_q, nq, r2,
                                                        float_below(r2), d_offsets, d_keys, c.d_error);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

// Radius neighbourhoods of a query batch as CSR rows addressed by the query's slot, in two steps so the caller can look
// at the total before materialising anything:
//   radius_count       counts / offsets (nq + 1 entries, offsets[nq] = total); synchronises once to read the total
//   radius_fill_sorted keys = (d2 bits << 32) | original index, ascending inside every row (= ascending (d2, index))
void radius_count(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2, DevBuf<unsigned long long>& counts,
                  DevBuf<unsigned long long>& offsets, unsigned long long& total)
{
  cudaStream_t st = c.stream;
  counts.alloc(nq + 1, st);
  offsets.alloc(nq + 1, st);
  PCLB_CUDA(cudaMemsetAsync(counts.p, 0, (nq + 1) * sizeof(unsigned long long), st));
  launch_radius_count(c, idx, d_q, nq, r2, counts.p);
  size_t tb = 0;
  PCLB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, offsets.p, (int)(nq + 1), st));
  DevBuf<unsigned char> tmp;
  tmp.alloc(tb, st);
  PCLB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, offsets.p, (int)(nq + 1), st));
  ++c.launches;
  total = 0;
  PCLB_CUDA(cudaMemcpyAsync(&total, offsets.p + nq, sizeof(total), cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
}

void radius_fill_sorted(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2,
                        const DevBuf<unsigned long long>& offsets, unsigned long long total,
                        DevBuf<unsigned long long>& keys_sorted)
{
  if (total == 0)
    return;
  cudaStream_t st = c.stream;
  PCLB_REQUIRE(total < (unsigned long long)std::numeric_limits<int>::max(), PCLB200_ERR_INVALID,
               "radius search result exceeds 2^31 neighbours; lower the radius or set max_nn (<= 32 is searched without "
               "materialising the whole ball)");
  DevBuf<unsigned long long> keys;
  keys.alloc(total, st);
  keys_sorted.alloc(total, st);
  launch_radius_fill(c, idx, d_q, nq, r2, offsets.p, keys.p);
  size_t tb2 = 0;
  PCLB_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, tb2, keys.p, keys_sorted.p, (int)total, (int)nq, offsets.p,
                                               offsets.p + 1, st));
  DevBuf<unsigned char> tmp2;
  tmp2.alloc(tb2, st);
  PCLB_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp2.p, tb2, keys.p, keys_sorted.p, (int)total, (int)nq, offsets.p,
                                               offsets.p + 1, st));
  c.launches += 3;
}

void radius_csr(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2, DevBuf<unsigned long long>& counts,
                DevBuf<unsigned long long>& offsets, DevBuf<unsigned long long>& keys_sorted, unsigned long long& total)
{
  radius_count(c, idx, d_q, nq, r2, counts, offsets, total);
  radius_fill_sorted(c, idx, d_q, nq, r2, offsets, total, keys_sorted);
}

// ---- normals: eigen33 / moments (knn_warp.cuh) ------------------------------------------------------------------
// k-NN -> shifted single-pass covariance in the neighbour order the search returns
// (common/include/pcl/common/impl/centroid.hpp:578-652, Scalar = float, same operation order, no fma)
// -> solvePlaneParameters (features/impl/feature.hpp:65-92) -> flipNormalTowardsViewpoint
// (features/normal_3d.h:169-188).  The neighbour list is never materialised in HBM.
template <int K>
__global__ void __launch_bounds__(128)
k_normals(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts, int root,
          const float4* __restrict__ q, size_t nq, int k_req, float vpx, float vpy, float vpz,
          float4* __restrict__ out, int* __restrict__ not_dense, int* __restrict__ d_error,
          const unsigned char* __restrict__ only = nullptr)
{
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq || (only && !only[i]))
    return;
  const float4 qq = __ldg(q + i);
  const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
  const float qnan = __int_as_float(0x7fc00000);
  if (!(isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z))) {
    out[slot] = make_float4(qnan, qnan, qnan, qnan);
    *not_dense = 1;
    return;
  }
  NearestK<K> v;
  v.qx = qq.x; v.qy = qq.y; v.qz = qq.z;
  v.pts = pts;
  v.init(__int_as_float(0x7f800000));
  if (!traverse(nodes, pts, root, qq.x, qq.y, qq.z, v))
    atomicExch(d_error, 1);
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (j < k_req && v.pos[j] >= 0)
      ++cnt;
  if (cnt < 3) {
    out[slot] = make_float4(qnan, qnan, qnan, qnan);
    *not_dense = 1;
    return;
  }
  float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float Kx = 0.f, Ky = 0.f, Kz = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (j < cnt) {
      const float4 p = ldg4(pts + v.pos[j]);
      if (j == 0) { Kx = p.x; Ky = p.y; Kz = p.z; }
      moments_add(accu, Kx, Ky, Kz, p);
    }
  out[slot] = normal_from_moments(accu, cnt, qq, vpx, vpy, vpz, not_dense);
}

// normals from materialised neighbour lists (k > 32): same arithmetic as k_normals, neighbours fetched through
// pos_of_orig.  lists are rows of pitch k indexed by the query's slot.
__global__ void __launch_bounds__(128)
k_normals_from_lists(const float4* __restrict__ pts, const int32_t* __restrict__ pos_of_orig,
                     const float4* __restrict__ q, size_t nq, int k, const int32_t* __restrict__ lists, size_t slot0,
                     float vpx, float vpy, float vpz, float4* __restrict__ out, int* __restrict__ not_dense)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq)
    return;
  const float4 qq = __ldg(q + i);
  const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
  const int32_t* nn = lists + (slot - slot0) * (size_t)k;
  const float qnan = __int_as_float(0x7fc00000);
  int cnt = 0;
  for (int j = 0; j < k; ++j)
    if (nn[j] >= 0)
      ++cnt;
  if (!(isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z)) || cnt < 3) {
    out[slot] = make_float4(qnan, qnan, qnan, qnan);
    *not_dense = 1;
    return;
  }
  float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float Kx = 0.f, Ky = 0.f, Kz = 0.f;
  for (int j = 0; j < cnt; ++j) {
    const float4 p = ldg4(pts + pos_of_orig[nn[j]]);
    if (j == 0) { Kx = p.x; Ky = p.y; Kz = p.z; }
    moments_add(accu, Kx, Ky, Kz, p);
  }
  out[slot] = normal_from_moments(accu, cnt, qq, vpx, vpy, vpz, not_dense);
}

// normals from radius neighbourhoods (setRadiusSearch): rows of a CSR of packed keys (radius_csr), addressed by the
// query's slot; same arithmetic as k_normals over a variable-length, (d2, index)-ascending neighbour list.
__global__ void __launch_bounds__(128)
k_normals_from_csr(const float4* __restrict__ pts, const int32_t* __restrict__ pos_of_orig,
                   const float4* __restrict__ q, size_t nq, const unsigned long long* __restrict__ offsets,
                   const unsigned long long* __restrict__ keys, float vpx, float vpy, float vpz,
                   float4* __restrict__ out, int* __restrict__ not_dense)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq)
    return;
  const float4 qq = __ldg(q + i);
  const size_t slot = (size_t)(unsigned)__float_as_int(qq.w);
  const unsigned long long b = offsets[slot], e = offsets[slot + 1];
  const float qnan = __int_as_float(0x7fc00000);
  if (!(isfinite(qq.x) && isfinite(qq.y) && isfinite(qq.z)) || e - b < 3ULL) {  // normal_3d.h:308-312
    out[slot] = make_float4(qnan, qnan, qnan, qnan);
    *not_dense = 1;
    return;
  }
  float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float Kx = 0.f, Ky = 0.f, Kz = 0.f;
  for (unsigned long long j = b; j < e; ++j) {
    const int oi = (int)(unsigned)(keys[j] & 0xffffffffULL);
    const float4 p = ldg4(pts + pos_of_orig[oi]);
    if (j == b) { Kx = p.x; Ky = p.y; Kz = p.z; }
    moments_add(accu, Kx, Ky, Kz, p);
  }
  out[slot] = normal_from_moments(accu, (int)(e - b), qq, vpx, vpy, vpz, not_dense);
}


#ifndef PCLB_KNN_TARGET_X10
#define PCLB_KNN_TARGET_X10 12
#endif
#ifndef PCLB_KNN_LEVEL_X10
#define PCLB_KNN_LEVEL_X10 10
#endif

// Where the warp kernel starts.  The ball it certifies should hold a little more than k points: from the occupancy of
// the cell levels, n / occupied(l) points per cell and the ratio of two levels give the cloud's local density and
// intrinsic dimension (4 per level = a surface, 8 = a volume), hence the radius of a ball of ~1.2 k + 4 points.  The
// start level is the finest whose half cell holds ~k points by that estimate (the ball's box must span <= 2 cells per axis); the
// first attempt uses the sized ball when it is well inside half a cell, later attempts half a cell of ever coarser
// levels.  Only speed depends on this estimate: an attempt counts only if the k-th distance certifies it.
struct WarpKnnStart {
  int level = 0;      // 0: no level fits (tiny cloud) -> per-thread kernels
  float r_first = 0;  // 0: start with half a cell
};
static WarpKnnStart warp_knn_start(const Index& idx, int k)
{
  WarpKnnStart w;
  const CellTableHost& c = idx.cells;
  auto avg = [&](int l) { return (double)idx.n_valid / (double)std::max<unsigned long long>(1, c.occupied[l]); };
  int ref = 0;
  for (int l = 2; l <= c.bmax; ++l)
    if (c.occupied[l] > 0 && c.occupied[l - 1] > 0 && avg(l) >= 8.0)
      ref = l;
  if (ref < 2)
    return w;
  const double dim = std::min(3.0, std::max(1.0, std::log2(avg(ref - 1) / avg(ref))));
  const double unit_ball = std::pow(M_PI, 0.5 * dim) / std::tgamma(0.5 * dim + 1.0);
  const double width = (double)(1u << (21 - ref)) / (double)idx.morton_scale;
  auto ball_of = [&](double points) { return width * std::pow(points / (avg(ref) * unit_ball), 1.0 / dim); };
  const double radius = ball_of(PCLB_KNN_TARGET_X10 * 0.1 * k + 4.0);
  // the level: the finest whose half cell still holds ~k points by the estimate, which runs ~15 % large on tilted
  // surfaces (a level finer gathers 2-3 times fewer candidates,
  // measured 16 vs 20 ms at k = 16, and its half-cell ball certifies most queries)
  const double need = ball_of(PCLB_KNN_LEVEL_X10 * 0.1 * k);
  for (int l = 1; l <= c.bmax; ++l) {
    const double half = 0.5 * 0.999 * (double)(1u << (21 - l)) / (double)idx.morton_scale;
    if (c.occupied[l] > 0 && half >= need)
      w.level = l;
  }
  if (w.level >= 1) {
    const double half = 0.5 * 0.999 * (double)(1u << (21 - w.level)) / (double)idx.morton_scale;
    if (radius < 0.85 * half)
      w.r_first = (float)radius;
  }
  return w;
}

static bool warp_knn_lists(Ctx& c, const Index& idx, const float4* d_q, size_t nq, int k, int32_t* d_out_idx,
                           float* d_out_d2, DevBuf<unsigned char>& redo)
{
  if (!(idx.cell_slots.p && idx.node_leaves.p) || k < 12 || k > 32 || (size_t)k > idx.n_valid)
    return false;
  const WarpKnnStart w0 = warp_knn_start(idx, k);
  if (w0.level < 1)
    return false;
  cudaStream_t s = c.stream;
  redo.alloc(nq, s);
  PCLB_CUDA(cudaMemsetAsync(redo.p, 0, nq, s));
  const unsigned wg = (unsigned)std::min<size_t>((nq + 255) / 256, (size_t)c.sm_count * 8);
  k_knn_warp<false><<<wg, 256, 0, s>>>(tree_view(idx), idx.node_leaves.p, w0.level, w0.r_first, d_q, nq, k, d_out_idx, d_out_d2, 0.f, 0.f, 0.f,
                                      nullptr, nullptr, redo.p);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
  return true;
}

void launch_normals_radius(Ctx& c, Index& idx, const float4* d_q, size_t nq, float r2, const float vp[3], float4* d_out,
                           int* d_not_dense)
{
  if (!nq)
    return;
  cudaStream_t s = c.stream;
  DevBuf<unsigned long long> counts, offsets, keys;
  unsigned long long total = 0;
  radius_csr(c, idx, d_q, nq, r2, counts, offsets, keys, total);
  ensure_pos_of_orig(c, idx);
  k_normals_from_csr<<<grid_for(nq, 128), 128, 0, s>>>(idx.pts.p, idx.pos_of_orig.p, d_q, nq, offsets.p, keys.p, vp[0],
                                                      vp[1], vp[2], d_out, d_not_dense);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

void launch_normals(Ctx& c, Index& idx, const float4* d_q, size_t nq, int k, const float vp[3], float4* d_out,
                    int* d_not_dense)
{
  if (!nq)
    return;
  cudaStream_t s = c.stream;
  const unsigned g = grid_for(nq, 128);
  if (k > 32) {
    // large neighbourhoods: materialise the exact k-NN lists (any-k kernel), then fold them
    ensure_pos_of_orig(c, idx);
    DevBuf<int32_t> li;
    DevBuf<float> ld;
    li.alloc(nq * (size_t)k, s);
    ld.alloc(nq * (size_t)k, s);
    launch_knn(c, idx, d_q, nq, k, __builtin_inff(), li.p, ld.p);
    k_normals_from_lists<<<g, 128, 0, s>>>(idx.pts.p, idx.pos_of_orig.p, d_q, nq, k, li.p, 0, vp[0], vp[1], vp[2], d_out,
                                           d_not_dense);
    ++c.launches;
    PCLB_CUDA(cudaGetLastError());
    return;
  }
  DevBuf<unsigned char> redo;
  const unsigned char* only = nullptr;
  {
    const WarpKnnStart w0 = (idx.cell_slots.p && idx.node_leaves.p && k >= 12) ? warp_knn_start(idx, k) : WarpKnnStart{};
    if (w0.level >= 1 && (size_t)k <= idx.n_valid) {
      redo.alloc(nq, s);
      PCLB_CUDA(cudaMemsetAsync(redo.p, 0, nq, s));
      const unsigned wg = (unsigned)std::min<size_t>((nq + 255) / 256, (size_t)c.sm_count * 8);
      k_knn_warp<true><<<wg, 256, 0, s>>>(tree_view(idx), idx.node_leaves.p, w0.level, w0.r_first, d_q, nq, k, nullptr, nullptr, vp[0], vp[1],
                                         vp[2], d_out, d_not_dense, redo.p);
      ++c.launches;
      only = redo.p;
    }
  }
#define PCLB_NRM_CASE(KK)                                                                                              \
  k_normals<KK><<<g, 128, 0, s>>>(idx.nodes.p, idx.pts.p, idx.root, d_q, nq, k, vp[0], vp[1], vp[2], d_out, d_not_dense, \
                                  c.d_error, only)
  if (k <= 4) PCLB_NRM_CASE(4);
  else if (k <= 8) PCLB_NRM_CASE(8);
  else if (k <= 10) PCLB_NRM_CASE(10);
  else if (k <= 16) PCLB_NRM_CASE(16);
  else if (k <= 20) PCLB_NRM_CASE(20);
  else PCLB_NRM_CASE(32);
#undef PCLB_NRM_CASE
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

}  // namespace pclb200
