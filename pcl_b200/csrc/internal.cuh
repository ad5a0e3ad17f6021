// internal.cuh — shared declarations of libpclb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pclb200.h"

namespace pclb200 {

// ---------------------------------------------------------------------------------------------
// errors: C++ exceptions inside the library, converted to status codes at the C boundary
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PCLB_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw ::pclb200::Error(PCLB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + \
                                                   " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

#define PCLB_REQUIRE(cond, code, msg)                 \
  do {                                                \
    if (!(cond))                                      \
      throw ::pclb200::Error((code), (msg));          \
  } while (0)

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct Comm;  // NCCL wrapper (comm.cu)

struct ProfEvent {
  const char* name;
  cudaEvent_t a, b;
};

struct Ctx {
  std::recursive_mutex mu;           // every C-ABI entry point holds it: calls on one ctx serialise (PCL's query
                                     // methods are const and may be called from several OpenMP threads)
  int device = 0;
  bool profiling = false;            // pclb200_profile_enable: CUDA-event pairs around the named kernels
  std::vector<ProfEvent> prof;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;  // kernels launched by this library (incl. CUB passes, counted per call)
  void* pinned = nullptr; // small pinned staging block for per-iteration read-backs
  size_t pinned_bytes = 0;
  int* d_error = nullptr; // device-side invariant flag (traversal stack overflow)
  Comm* comm = nullptr;
};

// records a CUDA-event pair around a region of the ctx stream when profiling is on
struct ProfScope {
  Ctx& c;
  cudaEvent_t b = nullptr;
  ProfScope(Ctx& ctx, const char* name) : c(ctx)
  {
    if (!c.profiling)
      return;
    ProfEvent e{name, nullptr, nullptr};
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess)
      return;
    cudaEventRecord(e.a, c.stream);
    b = e.b;
    c.prof.push_back(e);
  }
  ~ProfScope()
  {
    if (b)
      cudaEventRecord(b, c.stream);
  }
};

// Device memory comes from a small per-stream block cache (alloc.cu): after the first align() every request is
// served from blocks the previous call returned, so the steady-state loop never enters the driver's allocator
// (cudaMallocAsync pool growth showed up as multi-millisecond host stalls between launches on shared boxes).
// Re-use is safe without events because a ctx issues all its work on ONE stream: a block handed out again is
// only touched by work enqueued after the work that used it before.
void* cached_alloc(cudaStream_t s, size_t bytes);
void cached_free(cudaStream_t s, void* p);
void cached_release_all(cudaStream_t s);

// stream-ordered device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept
  {
    if (this != &o) {
      release();
      p = o.p; n = o.n; s = o.s;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  void alloc(size_t count, cudaStream_t stream)
  {
    release();
    s = stream;
    n = count;
    if (count)
      p = static_cast<T*>(cached_alloc(stream, count * sizeof(T)));
  }
  void release()
  {
    if (p)
      cached_free(s, p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  size_t bytes() const { return n * sizeof(T); }
};

bool is_device_ptr(const void* p);

// Copies `count` records of `rec_bytes` (<= stride) from a strided host-or-device array into a
// dense device array of float4 {x,y,z,w_fill}: the layout every kernel reads with 128-bit loads.
// If `subset` != NULL, record i is src[subset[i]].
void load_xyz_as_float4(Ctx& c, const void* src, size_t n_records, size_t stride, const int32_t* subset,
                        size_t n_subset, float4* d_out, cudaStream_t s);
// same for normals (3 floats at `src`, any stride) -> float4 {nx,ny,nz,0}
void load_vec3_as_float4(Ctx& c, const void* src, size_t n_records, size_t stride, float4* d_out,
                         cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// LBVH index
// ---------------------------------------------------------------------------------------------
// One internal node = 64 B = four 128-bit loads: both children's boxes + their references.
//   a = (l.min.x, l.min.y, l.min.z, l.max.x)  b = (l.max.y, l.max.z, r.min.x, r.min.y)
//   c = (r.min.z, r.max.x, r.max.y, r.max.z)  d = (left, right, 0, 0)
// child reference >= 0: internal node id; < 0: ~leaf_id.
struct __align__(16) BvhNode {
  float4 a, b, c;
  int4 d;
};

#ifndef PCLB_LEAF
#define PCLB_LEAF 8
#endif
constexpr int kLeafSize = PCLB_LEAF;  // points per leaf (multiple of 8): 8 points = one 128-byte line of float4
static_assert(kLeafSize % 8 == 0 && kLeafSize >= 8 && kLeafSize <= 64, "leaf size must be a multiple of 8");
// A walk pushes at most one entry per level.  The radix tree is at most 63 (key bits) + 31 (index tie-break bits)
// = 94 levels deep, so 96 entries can never overflow (the overflow check stays as a guard, not as an expected path).
constexpr int kStackSize = 96;      // traversal stack entries per query
constexpr int kSentinelIndex = 0x7fffffff;

// host-side description of the index's cell table (lbvh.cu builds it, traverse.cuh: CellTable reads it)
struct CellTableHost {
  unsigned log2_slots = 0;  // 0 = no table
  int bmax = 0;             // finest level present (cells of 2^-bmax of the frame per axis)
  uint64_t entries = 0;
  float margin = 0.f;
  uint64_t occupied[16] = {0};  // occupied cells per level (index = level), from the prefix-length histogram
};

struct Index {
  Ctx* ctx = nullptr;
  int device = 0;
  size_t n_cloud = 0;   // records in the caller's cloud (index space of results)
  size_t n_valid = 0;   // finite points indexed
  int n_leaves = 0;
  int root = 0;         // child-reference encoding
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};  // bounding box of the valid points
  float morton_scale = 1.f;                     // 2^21 / max extent
  DevBuf<float4> pts;   // n_leaves*kLeafSize, Morton order, w = original index bits; padded with +inf
  DevBuf<BvhNode> nodes;  // n_leaves-1
  DevBuf<int32_t> pos_of_orig;  // n_cloud: position in `pts` of original index i, or -1 (lazy)
  DevBuf<int2> node_leaves;     // per internal node: {first leaf, number of leaves} — a subtree's leaves are consecutive
  DevBuf<uint2> cell_slots;     // hash table (level, cell) -> deepest node / leaf holding every point of the cell
  CellTableHost cells;          // its description (traverse.cuh: CellTable is the device-side view)
  size_t bytes() const { return pts.bytes() + nodes.bytes() + pos_of_orig.bytes() + cell_slots.bytes() + node_leaves.bytes(); }
};

Index* build_index(Ctx& c, const void* pts, size_t n, size_t stride, const int32_t* subset, size_t n_subset);
// device-resident source already dense float4 (w ignored): used for the reciprocal source tree
Index* build_index_from_device(Ctx& c, const float4* d_pts, size_t n, const int32_t* d_orig /*nullable*/);
void ensure_pos_of_orig(Ctx& c, Index& idx);

// Morton-ordered batch of query points (dense float4 xyz + original slot), so neighbouring threads
// walk neighbouring subtrees.
struct QueryBatch {
  size_t n = 0;
  DevBuf<float4> q;       // n, xyz + w = original slot bits
};
void make_query_batch(Ctx& c, const Index& ref_frame, const float4* d_q /*n, w ignored*/, size_t n,
                      QueryBatch& out);

// ---------------------------------------------------------------------------------------------
// search / registration kernels (search.cu, icp.cu, voxel.cu)
// ---------------------------------------------------------------------------------------------
// exact k-NN for a Morton-ordered batch; results scattered to row `slot` (w of the query)
void launch_knn(Ctx& c, const Index& idx, const float4* d_q, size_t nq, int k, float init_bound,
                int32_t* d_out_idx, float* d_out_d2 /* nq*k, row = slot */);
// counts / fills for radius search: neighbours with d2 < r2
void launch_radius_count(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2,
                         unsigned long long* d_counts /* by slot */);
void launch_radius_fill(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2,
                        const unsigned long long* d_offsets /* by slot, exclusive */,
                        unsigned long long* d_keys /* (d2 bits << 32) | index */);

// sorted radius neighbourhoods as CSR rows of packed keys ((d2 bits << 32) | original index), addressed by query slot
void radius_count(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2, DevBuf<unsigned long long>& counts,
                  DevBuf<unsigned long long>& offsets, unsigned long long& total);
void radius_fill_sorted(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2,
                        const DevBuf<unsigned long long>& offsets, unsigned long long total,
                        DevBuf<unsigned long long>& keys_sorted);
void radius_csr(Ctx& c, const Index& idx, const float4* d_q, size_t nq, float r2, DevBuf<unsigned long long>& counts,
                DevBuf<unsigned long long>& offsets, DevBuf<unsigned long long>& keys_sorted, unsigned long long& total);
void launch_normals_radius(Ctx& c, Index& idx, const float4* d_q, size_t nq, float r2, const float vp[3], float4* d_out,
                           int* d_not_dense);
// correspondences chosen among the k nearest with the help of normals; surface-normal rejector (normals_corr.cu)
size_t correspondences_by_normals(Ctx& c, Index& tgt, int kind, const void* src, size_t n, size_t stride,
                                  const void* src_normals, size_t stride_sn, const void* tgt_normals, size_t stride_tn,
                                  const int32_t* indices, size_t n_idx, int k, double max_dist, pclb200_corr* out);
size_t reject_surface_normal(Ctx& c, const pclb200_corr* in, size_t n, const void* src_normals, size_t n_src,
                             size_t stride_sn, const void* tgt_normals, size_t n_tgt, size_t stride_tn,
                             double threshold, pclb200_corr* out);

// connected components of the d2 < tolerance^2 graph over the indexed points (cluster.cu)
void cluster_labels(Ctx& c, const Index& idx, double tolerance, int32_t* d_labels /* idx.n_cloud */);

// flat per-correspondence arrays the rejectors work on (reject.cu)
struct RejectArrays {
  size_t n = 0;
  const float* d2 = nullptr;      // squared distance
  const int* match = nullptr;     // index_match (any non-negative id that identifies the target point)
  const unsigned* tie = nullptr;  // position in the input list (tie-break)
  int* acc = nullptr;             // in/out: 1 = still a correspondence
};
void apply_rejector(Ctx& c, const pclb200_rejector& r, const RejectArrays& a, int* perm, int* keep_sorted, double* d_info,
                    int* trimmed_flag);

// accumulators of one ICP iteration (all fp64), see icp.cu
constexpr int kAccum = 40;

// Cross-GPU exchange buffers of the fused reduce (comm.cu): every rank owns one block
//   flags[kMaxRanks] (u64, latest sequence number published by each peer) | slots[2][kMaxRanks][kAccum] (fp64)
// mapped into every peer through CUDA IPC, so a kernel on rank r can store its partial sums straight into
// rank p's slots[.][r][.] over NVLink.
constexpr int kMaxRanks = 8;
struct PeerView {
  int rank = 0;
  int nranks = 0;                       // 0 / 1 = no exchange
  unsigned long long* flags[kMaxRanks]; // flags[p] = base of rank p's flag array (peer-mapped, [rank] = self)
  double* slots[kMaxRanks];             // slots[p] = base of rank p's slot array
};

}  // namespace pclb200
