// lbvh.cu — target-side data layout: Morton-sorted linear BVH in HBM.
//
// Replaces what pcl::KdTreeFLANN::setInputCloud builds on the CPU
// (kdtree/include/pcl/kdtree/impl/kdtree_flann.hpp:100-136, 429-498: drop non-finite points, keep
// index_mapping_, build a FLANN KDTreeSingleIndex with <=15-point leaves).  Here:
//   1. bbox + finite count (one streaming pass, block reduce + atomics on ordered-int floats)
//   2. 63-bit Morton key per point (21 bits/axis, ONE isotropic scale so cells are cubes);
//      non-finite points get key ~0 and sort to the tail
//   3. cub::DeviceRadixSort (key64, value = original index) — stable, so equal keys keep index order
//   4. Karras 2012 binary radix tree over the POINTS (ties by index), cut where a cell holds <= 8 points: every leaf is a
//      whole radix-tree cell, stored as one 128-byte line of float4 {x,y,z, original-index bits} padded with +inf
//   5. bottom-up refit (atomic arrival flags), then both children's boxes are packed into the parent's 64-byte node so
//      one node visit = four independent 128-bit loads
//   6. cell table: every occupied cell of the levels 1..bmax of the Morton grid -> the deepest node / leaf that holds all
//      its points (open-addressing hash, 8 B per slot), so that walks start at the candidate ball instead of the root
//      (traverse.cuh: CellTable, nearest1)
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>

#include "internal.cuh"
#include "traverse.cuh"
#include "lbvh_kernels.cuh"

namespace pclb200 {

bool is_device_ptr(const void* p)
{
  if (!p)
    return false;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ---- strided record -> dense float4 ------------------------------------------------------------
__global__ void k_strided_to_float4(const unsigned char* __restrict__ src, size_t stride,
                                    const int32_t* __restrict__ subset, size_t n, float w_fill,
                                    float4* __restrict__ out)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  size_t r = subset ? (size_t)subset[i] : i;
  const float* p = reinterpret_cast<const float*>(src + r * stride);
  out[i] = make_float4(p[0], p[1], p[2], w_fill);
}

static void strided_to_float4(Ctx& c, const void* src, size_t n_records, size_t stride, const int32_t* subset,
                              size_t n_subset, float w_fill, float4* d_out, cudaStream_t s)
{
  const size_t n = subset ? n_subset : n_records;
  if (n == 0)
    return;
  PCLB_REQUIRE(src != nullptr, PCLB200_ERR_INVALID, "null point array");
  PCLB_REQUIRE(stride >= 12 && stride % 4 == 0, PCLB200_ERR_INVALID, "stride must be a multiple of 4 and >= 12");
  const bool on_device = is_device_ptr(src);
  if (!subset && stride == 16) {
    // pcl::PointXYZ records already are float4: one straight copy (w is never read by the kernels)
    PCLB_CUDA(cudaMemcpyAsync(d_out, src, n * 16, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    return;
  }
  DevBuf<unsigned char> staged;
  DevBuf<int32_t> staged_sub;
  const unsigned char* d_src = static_cast<const unsigned char*>(src);
  if (!on_device) {
    // only up to the last byte a record's xyz can occupy: `src` may point INSIDE the caller's records (normals at
    // offset 16 of a PointNormal), so n_records * stride bytes would run past the end of the caller's array
    const size_t bytes = (n_records - 1) * stride + 12;
    staged.alloc(bytes, s);
    PCLB_CUDA(cudaMemcpyAsync(staged.p, src, bytes, cudaMemcpyHostToDevice, s));
    d_src = staged.p;
  }
  const int32_t* d_sub = subset;
  if (subset && !is_device_ptr(subset)) {
    staged_sub.alloc(n_subset, s);
    PCLB_CUDA(cudaMemcpyAsync(staged_sub.p, subset, n_subset * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    d_sub = staged_sub.p;
  }
  k_strided_to_float4<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_src, stride, d_sub, n, w_fill, d_out);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

void load_xyz_as_float4(Ctx& c, const void* src, size_t n_records, size_t stride, const int32_t* subset,
                        size_t n_subset, float4* d_out, cudaStream_t s)
{
  strided_to_float4(c, src, n_records, stride, subset, n_subset, 1.f, d_out, s);
}

void load_vec3_as_float4(Ctx& c, const void* src, size_t n_records, size_t stride, float4* d_out, cudaStream_t s)
{
  strided_to_float4(c, src, n_records, stride, nullptr, 0, 0.f, d_out, s);
}

static inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

struct SortedCloud {
  DevBuf<unsigned long long> keys;  // sorted
  DevBuf<int32_t> vals;             // sorted slots
  size_t n_valid = 0;
  float lo[3], hi[3];
  float scale = 1.f;
};

// bbox (when frame == nullptr) + keys + radix sort of n dense float4 points
static void morton_sort(Ctx& c, const float4* d_pts, size_t n, const Index* frame, SortedCloud& out)
{
  cudaStream_t s = c.stream;
  if (frame) {
    for (int d = 0; d < 3; ++d) {
      out.lo[d] = frame->lo[d];
      out.hi[d] = frame->hi[d];
    }
    out.scale = frame->morton_scale;
    out.n_valid = n;  // callers of the query path never pass non-finite points they care about
  }
  else {
    DevBuf<BBoxAcc> acc;
    acc.alloc(1, s);
    BBoxAcc init;
    for (int d = 0; d < 3; ++d) {
      init.lo[d] = 0x7fffffff;
      init.hi[d] = (int)0x80000000;
    }
    init.count = 0;
    PCLB_CUDA(cudaMemcpyAsync(acc.p, &init, sizeof(init), cudaMemcpyHostToDevice, s));
    unsigned g = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)c.sm_count * 8);
    k_bbox<<<g ? g : 1, 256, 0, s>>>(d_pts, n, acc.p);
    ++c.launches;
    BBoxAcc h;
    PCLB_CUDA(cudaMemcpyAsync(&h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    PCLB_CUDA(cudaStreamSynchronize(s));
    out.n_valid = (size_t)h.count;
    if (out.n_valid == 0)
      return;
    float ext = 0.f;
    for (int d = 0; d < 3; ++d) {
      out.lo[d] = ord2f(h.lo[d]);
      out.hi[d] = ord2f(h.hi[d]);
      ext = std::max(ext, out.hi[d] - out.lo[d]);
    }
    out.scale = ext > 0.f ? 2097152.f / ext : 1.f;
    if (!std::isfinite(out.scale))
      out.scale = 1.f;
  }
  DevBuf<unsigned long long> keys_in;
  DevBuf<int32_t> vals_in;
  keys_in.alloc(n, s);
  vals_in.alloc(n, s);
  out.keys.alloc(n, s);
  out.vals.alloc(n, s);
  const bool hilbert = frame != nullptr;  // queries: Hilbert order (compact runs of 32); the tree itself: Morton
  if (hilbert)
    k_morton<true><<<grid_for(n, 256), 256, 0, s>>>(d_pts, n, out.lo[0], out.lo[1], out.lo[2], out.scale, keys_in.p,
                                                   vals_in.p);
  else
    k_morton<false><<<grid_for(n, 256), 256, 0, s>>>(d_pts, n, out.lo[0], out.lo[1], out.lo[2], out.scale, keys_in.p,
                                                    vals_in.p);
  ++c.launches;
  size_t tmp_bytes = 0;
  // The index needs the full 63-bit order (leaves are radix-tree cells).  A query batch only needs spatial coherence: the
  // top 32 bits of the Hilbert index (10.7 bits per axis, cells about a leaf wide) order it just as well and halve the
  // radix passes (stable sort: points of one such cell keep their input order).
  const int begin_bit = frame ? 31 : 0, end_bit = frame ? 63 : 64;
  PCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.p, out.keys.p, vals_in.p, out.vals.p,
                                            (int)n, begin_bit, end_bit, s));
  DevBuf<unsigned char> tmp;
  tmp.alloc(tmp_bytes, s);
  PCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.p, out.keys.p, vals_in.p, out.vals.p, (int)n,
                                            begin_bit, end_bit, s));
  c.launches += frame ? 5 : 9;  // onesweep: histogram + one pass per 8-bit digit
  PCLB_CUDA(cudaGetLastError());
}

static Index* build_from_dense(Ctx& c, const float4* d_pts, size_t n, const int32_t* d_orig_of_slot, size_t n_cloud,
                                bool build_cell_table)
{
  cudaStream_t s = c.stream;
  PCLB_REQUIRE(n > 0, PCLB200_ERR_EMPTY, "cannot index an empty cloud (kdtree_flann.hpp:118-129)");
  PCLB_REQUIRE(n < (size_t)0x7fffffff, PCLB200_ERR_INVALID, "cloud too large for int32 indices");
  SortedCloud sc;
  morton_sort(c, d_pts, n, nullptr, sc);
  PCLB_REQUIRE(sc.n_valid > 0, PCLB200_ERR_EMPTY, "no finite point in the cloud (kdtree_flann.hpp:124-129)");
  std::unique_ptr<Index> idx(new Index());
  idx->ctx = &c;
  idx->device = c.device;
  idx->n_cloud = n_cloud;
  idx->n_valid = sc.n_valid;
  for (int d = 0; d < 3; ++d) {
    idx->lo[d] = sc.lo[d];
    idx->hi[d] = sc.hi[d];
  }
  idx->morton_scale = sc.scale;
  const int nv = (int)sc.n_valid;
  if (nv <= kLeafSize) {
    // the whole cloud is one leaf
    idx->n_leaves = 1;
    idx->root = ~0;
    idx->pts.alloc(kLeafSize, s);
    k_gather_sorted<<<1, 256, 0, s>>>(d_pts, sc.vals.p, d_orig_of_slot, sc.n_valid, kLeafSize, idx->pts.p);
    ++c.launches;
  }
  else {
    // 1. radix tree over the points, 2. cut it where a cell holds <= kLeafSize points, 3. renumber, 4. refit + pack
    const int ni = nv - 1;
    DevBuf<int2> kchildren, krange;
    DevBuf<int> knode_parent, kpoint_parent, keep, new_id, leaf_flag, leaf_incl;
    kchildren.alloc(ni, s);
    krange.alloc(ni, s);
    knode_parent.alloc(ni, s);
    kpoint_parent.alloc(nv, s);
    keep.alloc(ni, s);
    new_id.alloc(ni, s);
    leaf_flag.alloc(nv, s);
    leaf_incl.alloc(nv, s);
    PCLB_CUDA(cudaMemsetAsync(leaf_flag.p, 0, leaf_flag.bytes(), s));
    k_karras_points<<<grid_for(ni, 256), 256, 0, s>>>(sc.keys.p, nv, kchildren.p, knode_parent.p, kpoint_parent.p,
                                                      krange.p);
    k_mark_cells<<<grid_for(nv, 256), 256, 0, s>>>(nv, krange.p, knode_parent.p, kpoint_parent.p, keep.p, leaf_flag.p);
    size_t tb = 0, tb2 = 0;
    PCLB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, leaf_flag.p, leaf_incl.p, nv, s));
    PCLB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, keep.p, new_id.p, ni, s));
    DevBuf<unsigned char> tmp;
    tmp.alloc(std::max(tb, tb2), s);
    PCLB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, leaf_flag.p, leaf_incl.p, nv, s));
    PCLB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, keep.p, new_id.p, ni, s));
    c.launches += 4;
    DevBuf<unsigned> hist;
    unsigned h_hist[64] = {0};
    const bool want_cells = build_cell_table && nv >= 256;
    if (want_cells) {
      hist.alloc(64, s);
      PCLB_CUDA(cudaMemsetAsync(hist.p, 0, 64 * sizeof(unsigned), s));
      k_prefix_hist<<<std::min<unsigned>(grid_for(nv, 256), (unsigned)c.sm_count * 8), 256, 0, s>>>(sc.keys.p, nv, hist.p);
      ++c.launches;
      PCLB_CUDA(cudaMemcpyAsync(h_hist, hist.p, sizeof(h_hist), cudaMemcpyDeviceToHost, s));
    }
    int h_leaves = 0, h_last_id = 0, h_last_keep = 0;
    PCLB_CUDA(cudaMemcpyAsync(&h_leaves, leaf_incl.p + (nv - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    PCLB_CUDA(cudaMemcpyAsync(&h_last_id, new_id.p + (ni - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    PCLB_CUDA(cudaMemcpyAsync(&h_last_keep, keep.p + (ni - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    PCLB_CUDA(cudaStreamSynchronize(s));
    const int n_leaves = h_leaves;
    const int n_int = h_last_id + h_last_keep;
    PCLB_REQUIRE(n_leaves >= 2 && n_int == n_leaves - 1, PCLB200_ERR_INTERNAL, "LBVH build: inconsistent cell cut");
    idx->n_leaves = n_leaves;
    idx->root = 0;
    const size_t n_padded = (size_t)n_leaves * kLeafSize;
    idx->pts.alloc(n_padded, s);
    idx->nodes.alloc(n_int, s);
    DevBuf<int> leaf_start;
    DevBuf<int> node_parent, leaf_parent;  // build-time only (bottom-up refit)
    DevBuf<int2> children;
    DevBuf<float4> leaf_lo, leaf_hi, node_lo, node_hi;
    DevBuf<unsigned> flags;
    leaf_start.alloc(n_leaves, s);
    node_parent.alloc(n_int, s);
    leaf_parent.alloc(n_leaves, s);
    children.alloc(n_int, s);
    idx->node_leaves.alloc(n_int, s);
    leaf_lo.alloc(n_leaves, s);
    leaf_hi.alloc(n_leaves, s);
    node_lo.alloc(n_int, s);
    node_hi.alloc(n_int, s);
    flags.alloc(n_int, s);
    PCLB_CUDA(cudaMemsetAsync(flags.p, 0, flags.bytes(), s));
    k_leaf_starts<<<grid_for(nv, 256), 256, 0, s>>>(nv, leaf_flag.p, leaf_incl.p, leaf_start.p);
    k_fill_sentinels<<<grid_for(n_padded, 256), 256, 0, s>>>(idx->pts.p, n_padded);
    k_scatter_cells<<<grid_for(nv, 256), 256, 0, s>>>(d_pts, sc.vals.p, d_orig_of_slot, nv, leaf_incl.p, leaf_start.p,
                                                      idx->pts.p);
    // cell table (traverse.cuh: CellTable): levels 1..bmax, bmax = the finest level (<= kCellMaxBits) whose occupied
    // cells still hold >= 4 indexed points on average — finer levels would mostly map single points.  The number of
    // entries is known exactly from the prefix-length histogram; the table gets >= 2 slots per entry.
    CellTableW cellw{nullptr, 0, 0, 0};
    if (want_cells) {
      uint64_t cum = 1, entries = 0;
      int l = 0, bmax = 0;
      for (int b = 1; b <= kCellMaxBits; ++b) {
        for (; l < 3 * b; ++l)
          cum += h_hist[l];
        if (cum > (uint64_t)nv / 4)
          break;
        bmax = b;
        entries += cum;
        idx->cells.occupied[b] = cum;
      }
      if (bmax >= 1) {
        unsigned lg = 6;
        while (((uint64_t)1 << lg) < 2 * entries)
          ++lg;
        idx->cell_slots.alloc((size_t)1 << lg, s);
        PCLB_CUDA(cudaMemsetAsync(idx->cell_slots.p, 0, idx->cell_slots.bytes(), s));
        idx->cells.log2_slots = lg;
        idx->cells.bmax = bmax;
        idx->cells.entries = entries;
        float m = 0.f;
        for (int d = 0; d < 3; ++d)
          m = std::max(m, std::max(std::fabs(idx->lo[d]), std::fabs(idx->hi[d])));
        m = std::max(m, 2097152.f / idx->morton_scale);
        idx->cells.margin = 2e-6f * m;
        cellw.slots = reinterpret_cast<unsigned long long*>(idx->cell_slots.p);
        cellw.shift = 32 - lg;
        cellw.mask = (1u << lg) - 1u;
        cellw.bmax = bmax;
      }
    }
    k_link_cells<<<grid_for(ni, 256), 256, 0, s>>>(nv, kchildren.p, krange.p, keep.p, new_id.p, leaf_incl.p, sc.keys.p,
                                                   children.p, node_parent.p, leaf_parent.p, idx->node_leaves.p, cellw);
    k_refit<<<grid_for(n_leaves, 256), 256, 0, s>>>(idx->pts.p, n_leaves, children.p, node_parent.p, leaf_parent.p,
                                                    leaf_lo.p, leaf_hi.p, node_lo.p, node_hi.p, flags.p);
    k_pack_nodes<<<grid_for(n_int, 256), 256, 0, s>>>(n_int, children.p, leaf_lo.p, leaf_hi.p, node_lo.p, node_hi.p,
                                                     idx->nodes.p);
    c.launches += 8;
    PCLB_CUDA(cudaGetLastError());
  }
  PCLB_CUDA(cudaStreamSynchronize(s));  // temporaries are released stream-ordered; surface build errors here
  return idx.release();
}

Index* build_index(Ctx& c, const void* pts, size_t n, size_t stride, const int32_t* subset, size_t n_subset)
{
  cudaStream_t s = c.stream;
  const size_t cnt = subset ? n_subset : n;
  PCLB_REQUIRE(pts != nullptr && cnt > 0, PCLB200_ERR_EMPTY, "cannot index an empty cloud (kdtree_flann.hpp:118-129)");
  DevBuf<float4> dense;
  dense.alloc(cnt, s);
  load_xyz_as_float4(c, pts, n, stride, subset, n_subset, dense.p, s);
  DevBuf<int32_t> d_sub;
  const int32_t* d_orig = nullptr;
  if (subset) {
    if (is_device_ptr(subset))
      d_orig = subset;
    else {
      d_sub.alloc(n_subset, s);
      PCLB_CUDA(cudaMemcpyAsync(d_sub.p, subset, n_subset * sizeof(int32_t), cudaMemcpyHostToDevice, s));
      d_orig = d_sub.p;
    }
  }
  return build_from_dense(c, dense.p, cnt, d_orig, n, true);
}

Index* build_index_from_device(Ctx& c, const float4* d_pts, size_t n, const int32_t* d_orig)
{
  return build_from_dense(c, d_pts, n, d_orig, n, false);  // rebuilt every reciprocal iteration: no cell table
}

// ---- position of each original index in the Morton array (for gathers by index_match) -------------
__global__ void k_pos_of_orig(const float4* __restrict__ pts, size_t n_padded, int32_t* __restrict__ pos)
{
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n_padded)
    return;
  int oi = __float_as_int(pts[j].w);
  if (oi != kSentinelIndex)
    pos[oi] = (int32_t)j;
}

void ensure_pos_of_orig(Ctx& c, Index& idx)
{
  if (idx.pos_of_orig.p)
    return;
  cudaStream_t s = c.stream;
  idx.pos_of_orig.alloc(idx.n_cloud, s);
  PCLB_CUDA(cudaMemsetAsync(idx.pos_of_orig.p, 0xff, idx.pos_of_orig.bytes(), s));
  k_pos_of_orig<<<grid_for(idx.pts.n, 256), 256, 0, s>>>(idx.pts.p, idx.pts.n, idx.pos_of_orig.p);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

// ---- query batches ------------------------------------------------------------------------------
__global__ void k_gather_queries(const float4* __restrict__ q, const int32_t* __restrict__ vals, size_t n,
                                 float4* __restrict__ out)
{
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n)
    return;
  int32_t slot = vals[j];
  float4 v = __ldg(q + slot);
  out[j] = make_float4(v.x, v.y, v.z, __int_as_float(slot));
}

void make_query_batch(Ctx& c, const Index& frame, const float4* d_q, size_t n, QueryBatch& out)
{
  cudaStream_t s = c.stream;
  out.n = n;
  if (n == 0)
    return;
  SortedCloud sc;
  morton_sort(c, d_q, n, &frame, sc);
  out.q.alloc(n, s);
  k_gather_queries<<<grid_for(n, 256), 256, 0, s>>>(d_q, sc.vals.p, n, out.q.p);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

}  // namespace pclb200
