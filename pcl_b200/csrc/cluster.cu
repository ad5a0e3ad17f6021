// cluster.cu — Euclidean clustering = connected components of the "closer than the tolerance" graph.
//
// Replaces the breadth-first flood fill of pcl::extractEuclideanClusters
// (segmentation/include/pcl/segmentation/impl/extract_clusters.hpp:45-119, 124-223), which issues one radiusSearch per
// point from a single thread.  The squared distance is symmetric in fp32 (dist2_rn squares the same differences), so the
// sets the flood fill produces are exactly the connected components of the undirected graph with an edge wherever
// d2 < float(tolerance^2) (the strict FLANN radius test, kdtree_flann.hpp:398); they do not depend on the seed order.
// One thread per point walks the LBVH once and merges itself with every neighbour at a smaller position through a
// lock-free union-find (atomicCAS hooks the larger root under the smaller one, so parents only ever decrease and no cycle
// can form; path halving writes are benign races).  The label of a component is the smallest ORIGINAL index in it.
#include <limits>

#include "internal.cuh"
#include "traverse.cuh"

namespace pclb200 {

static inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

__device__ __forceinline__ int cc_find(int* parent, int x)
{
  volatile int* vp = parent;
  int cur = x;
  while (true) {
    const int p = vp[cur];
    if (p == cur)
      return cur;
    const int gp = vp[p];
    if (gp != p)
      vp[cur] = gp;  // path halving: gp was an ancestor of cur and stays one (roots never become roots again)
    cur = p;
  }
}

__device__ __forceinline__ void cc_unite(int* parent, int a, int b)
{
  while (true) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b)
      return;
    const int hi = a > b ? a : b, lo = a > b ? b : a;
    if (atomicCAS(parent + hi, hi, lo) == hi)
      return;
    a = hi;  // somebody hooked `hi` first: start again from the new roots
    b = lo;
  }
}

struct UnionVisitor {
  float qx, qy, qz, r2, r2_below;
  int self_pos;
  int* parent;
  __device__ __forceinline__ float bound() const { return r2_below; }
  __device__ __forceinline__ void prune(float) {}
  __device__ __forceinline__ void leaf(const float4* lp, int first_pos)
  {
#pragma unroll
    for (int j = 0; j < kLeafSize; ++j) {
      const float4 p = ldg4(lp + j);
      // padding slots hold +inf and never pass; every edge is handled once, from its larger end
      if (first_pos + j < self_pos && dist2_rn(qx, qy, qz, p.x, p.y, p.z) < r2)
        cc_unite(parent, self_pos, first_pos + j);
    }
  }
};

__global__ void k_cc_init(int* __restrict__ parent, int* __restrict__ min_orig, size_t n_padded)
{
  const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j < n_padded) {
    parent[j] = (int)j;
    min_orig[j] = kSentinelIndex;
  }
}

__global__ void __launch_bounds__(128)
k_cc_union(const BvhNode* __restrict__ nodes, const float4* __restrict__ pts, int root, size_t n_padded, float r2,
           float r2_below, int* parent, int* __restrict__ d_error)
{
  const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n_padded)
    return;
  const float4 q = ldg4(pts + j);
  if (__float_as_int(q.w) == kSentinelIndex)
    return;  // padding slot
  UnionVisitor v{q.x, q.y, q.z, r2, r2_below, (int)j, parent};
  if (!traverse(nodes, pts, root, q.x, q.y, q.z, v))
    atomicExch(d_error, 1);
}

// smallest original index of every component, kept at the component's root
__global__ void k_cc_min_orig(const float4* __restrict__ pts, size_t n_padded, int* parent, int* __restrict__ min_orig)
{
  const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n_padded)
    return;
  const int oi = __float_as_int(pts[j].w);
  if (oi == kSentinelIndex)
    return;
  atomicMin(min_orig + cc_find(parent, (int)j), oi);
}

__global__ void k_cc_labels(const float4* __restrict__ pts, size_t n_padded, int* parent, const int* __restrict__ min_orig,
                            int32_t* __restrict__ labels)
{
  const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n_padded)
    return;
  const int oi = __float_as_int(pts[j].w);
  if (oi == kSentinelIndex)
    return;
  labels[oi] = min_orig[cc_find(parent, (int)j)];
}

// labels: idx.n_cloud entries on the device; -1 for points the index does not hold (non-finite, outside the subset)
void cluster_labels(Ctx& c, const Index& idx, double tolerance, int32_t* d_labels)
{
  cudaStream_t st = c.stream;
  PCLB_CUDA(cudaMemsetAsync(d_labels, 0xff, idx.n_cloud * sizeof(int32_t), st));
  if (idx.n_valid == 0)
    return;
  // extract_clusters.hpp:246: the tolerance is narrowed to float before it reaches radiusSearch(point, double radius),
  // which squares it in double and narrows again (kdtree_flann.hpp:398)
  const double tol = (double)(float)tolerance;
  const float r2 = (float)(tol * tol);
  const float r2_below = std::nextafter(r2, -std::numeric_limits<float>::infinity());
  const size_t np = idx.pts.n;
  DevBuf<int> parent, min_orig;
  parent.alloc(np, st);
  min_orig.alloc(np, st);
  k_cc_init<<<grid_for(np, 256), 256, 0, st>>>(parent.p, min_orig.p, np);
  {
    ProfScope ps(c, "cluster_union");
    k_cc_union<<<grid_for(np, 128), 128, 0, st>>>(idx.nodes.p, idx.pts.p, idx.root, np, r2, r2_below, parent.p, c.d_error);
  }
  k_cc_min_orig<<<grid_for(np, 256), 256, 0, st>>>(idx.pts.p, np, parent.p, min_orig.p);
  k_cc_labels<<<grid_for(np, 256), 256, 0, st>>>(idx.pts.p, np, parent.p, min_orig.p, d_labels);
  c.launches += 4;
  PCLB_CUDA(cudaGetLastError());
}

}  // namespace pclb200
