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
#ifndef PCLB_KNN_TARGET_X10
#define PCLB_KNN_TARGET_X10 12
#endif
#ifndef PCLB_KNN_LEVEL_X10
#define PCLB_KNN_LEVEL_X10 10
#endif

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
