// icp.cu — the per-iteration hot loop of pcl::IterativeClosestPoint, device resident.
//
// Reference loop: registration/include/pcl/registration/impl/icp.hpp:164-241
//   determineCorrespondences (impl/correspondence_estimation.hpp:145-218)  N_src x 1-NN + distance gate
//   estimateRigidTransformation (impl/transformation_estimation_svd.hpp:137-155 -> Umeyama,
//       common/impl/eigen.hpp:675-734; or impl/transformation_estimation_point_to_plane_lls.hpp:166-268)
//   transformCloud (impl/icp.hpp:49-111) ; final = T_k * final (:223) ; convergence (:238)
//
// One iteration here = ONE streaming kernel over the Morton-ordered source:
//   [apply the previous iteration's T_k to the point, in the reference's fp32 operation order, and
//    write it back]  ->  exact 1-NN in the LBVH with the gate as the initial bound  ->  fp64 accumulation of
//   the 17 (Umeyama) or 29 (point-to-plane) sums  ->  warp/block/grid reduction with a FIXED order
//   (deterministic), last block folds the per-block partials
// followed (after the cross-GPU all-reduce, if any) by a one-warp solve kernel that leaves T_k in
// device memory for the next iteration.  The host only reads back ~200 bytes per iteration to run
// DefaultConvergenceCriteria (default_convergence_criteria.hpp:49-140) in the caller's Scalar.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>

#include "internal.cuh"
#include "corr_select.cuh"
#include "traverse.cuh"
#include "icp_kernels.cuh"

namespace pclb200 {

void comm_allreduce_sum(Ctx& c, double* d_buf, int count);  // comm.cu (no-op without a communicator)
bool comm_active(const Ctx& c);
bool comm_peer_view(Ctx& c, PeerView* view, unsigned long long* seq);
bool comm_peer_fused(const Ctx& c);  // the per-iteration exchange happens inside the accumulate kernel

static inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

// ---- misc streaming kernels ------------------------------------------------------------------------------
__global__ void k_apply_pending(float4* __restrict__ pts, size_t n, const Pending* __restrict__ pending,
                                float4* __restrict__ normals /* nullable */)
{
  const Pending P = *pending;
  if (!P.apply)
    return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 p = pts[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z)))
      continue;
    apply_pending(P, p.x, p.y, p.z);
    pts[i] = p;
    if (normals) {
      float4 m = normals[i];
      if (P.mode == 0 && !(isfinite(m.x) && isfinite(m.y) && isfinite(m.z)))
        continue;  // icp.hpp:97-98
      apply_pending_normal(P, m.x, m.y, m.z);
      normals[i] = m;
    }
  }
}

__global__ void k_clear_apply(Pending* p) { p->apply = 0; }

// strided write-back of a dense float4 array: xyz into out + i*stride (+ optional w = 1 when stride >= 16)
__global__ void k_scatter_xyz(const float4* __restrict__ src, size_t n, unsigned char* __restrict__ out, size_t stride,
                              int write_w, float w)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float4 v = src[i];
  float* o = reinterpret_cast<float*>(out + i * stride);
  o[0] = v.x; o[1] = v.y; o[2] = v.z;
  if (write_w)
    o[3] = w;
}

// =============================================================================================================
// host side
// =============================================================================================================
static void check_device_error(Ctx& c);

// ---- GICP covariances -----------------------------------------------------------------------------------------------
// GeneralizedIterativeClosestPoint::computeCovariances (registration/impl/gicp.hpp:69-147): per point, the covariance
// of its k nearest neighbours relative to the point (float differences widened to double), mean removed, SVD, singular
// values replaced by (1, 1, gicp_epsilon), reassembled from the columns of U.  One thread per point over the exact k-NN
// rows of launch_knn; the 3x3 Jacobi SVD is k_solve's.
void gicp_covariances(Ctx& c, Index& idx, const void* pts, size_t n, size_t stride, int k, double gicp_epsilon,
                      double* out)
{
  cudaStream_t st = c.stream;
  if (n == 0)
    return;
  PCLB_REQUIRE(k >= 1, PCLB200_ERR_INVALID, "k must be positive");
  const int keff = (int)std::min<size_t>((size_t)k, idx.n_valid);
  DevBuf<float4> dense;
  dense.alloc(n, st);
  load_xyz_as_float4(c, pts, n, stride, nullptr, 0, dense.p, st);
  QueryBatch qb;
  make_query_batch(c, idx, dense.p, n, qb);
  DevBuf<int32_t> rows;
  DevBuf<float> rows_d2;
  rows.alloc(n * (size_t)keff, st);
  rows_d2.alloc(n * (size_t)keff, st);
  launch_knn(c, idx, qb.q.p, n, keff, std::numeric_limits<float>::infinity(), rows.p, rows_d2.p);
  ensure_pos_of_orig(c, idx);
  DevBuf<double> d_out;
  const bool on_dev = is_device_ptr(out);
  double* d_o = out;
  if (!on_dev) {
    d_out.alloc(n * 9, st);
    d_o = d_out.p;
  }
  k_gicp_cov<<<grid_for(n, 128), 128, 0, st>>>(dense.p, n, rows.p, keff, k, idx.pts.p, idx.pos_of_orig.p, gicp_epsilon, d_o);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
  if (!on_dev)
    PCLB_CUDA(cudaMemcpyAsync(out, d_out.p, n * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
  check_device_error(c);
}

#ifdef PCLB_STATS
extern "C" __attribute__((visibility("default"))) int pclb200_debug_walk_stats(unsigned long long out[8], int reset)
{
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_walk_stats, 8 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_walk_stats, z, sizeof(z));
  }
  return 0;
}
extern "C" __attribute__((visibility("default"))) int pclb200_debug_walk_hist(unsigned long long out[66], int reset)
{
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_walk_imbalance, 2 * sizeof(unsigned long long));
  cudaMemcpyFromSymbol(out + 2, g_walk_hist, 64 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[64] = {0};
    cudaMemcpyToSymbol(g_walk_imbalance, z, 2 * sizeof(unsigned long long));
    cudaMemcpyToSymbol(g_walk_hist, z, sizeof(z));
  }
  return 0;
}
#endif

float gate_from_max_dist(double max_dist)
{
  // d2 (float) is kept iff (double)d2 <= max_dist^2  <=>  d2 <= largest float <= max_dist^2
  const double m2 = max_dist * max_dist;
  if (!(m2 < (double)FLT_MAX))
    return FLT_MAX;
  float g = (float)m2;
  if ((double)g > m2)
    g = std::nextafter(g, -INFINITY);
  return g;
}

// C = A*B row-major, Eigen coefficient order, in Scalar S (final = T_k * final, icp.hpp:223)
template <typename S>
static void mat4_mul(const S* A, const S* B, S* C)
{
  S R[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c)
      R[4 * r + c] = ((A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c]) + A[4 * r + 2] * B[8 + c]) + A[4 * r + 3] * B[12 + c];
  for (int i = 0; i < 16; ++i)
    C[i] = R[i];
}

struct Reducer {  // scratch for block_reduce_and_publish
  DevBuf<double> partials;
  DevBuf<double> partials2;  // max_blocks * 128 (k_accum_dmma)
  DevBuf<unsigned> counter;
  DevBuf<double> accum;
  unsigned max_blocks = 0;
  void init(Ctx& c, unsigned blocks)
  {
    max_blocks = blocks;
    partials.alloc((size_t)blocks * kAccum, c.stream);
    partials2.alloc((size_t)blocks * 128, c.stream);
    counter.alloc(1, c.stream);
    accum.alloc(kAccum, c.stream);
    PCLB_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned), c.stream));
    PCLB_CUDA(cudaMemsetAsync(accum.p, 0, kAccum * sizeof(double), c.stream));
  }
};

static unsigned persistent_grid(const Ctx& c, size_t n, int block, int per_sm)
{
  size_t want = (n + block - 1) / block;
  size_t cap = (size_t)c.sm_count * per_sm;
  return (unsigned)std::max<size_t>(1, std::min(want, cap));
}

struct Icp {
  Ctx* ctx = nullptr;
  pclb200_icp_params P;
  const Index* tgt = nullptr;
  DevBuf<float4> tgt_normals;   // Morton order of tgt
  // source
  size_t n_all = 0;             // records in the caller's source cloud
  size_t n_q = 0;               // indexed source points (queries)
  DevBuf<float4> src_all;       // original order, all records (for the output cloud)
  DevBuf<float4> src_normals;   // original order (optional)
  DevBuf<int32_t> src_orig;     // slot -> original source index (when indices were given)
  DevBuf<float4> cur;           // Morton order, w = slot
  DevBuf<float4> cur_normals;   // source normals in the order of `cur` (symmetric objective only)
  DevBuf<Match> match;          // Morton order: this iteration's matches = next iteration's seeds
  DevBuf<LoopCtrl> ctrl;        // device-side loop state of the enqueue-ahead path (icp_iterate)
  DevBuf<float> lb;             // per query: lower bound on the distance to every OTHER target point (TRACK searches)
  bool lb_valid = false;        // lb was written by the previous search
  DevBuf<int32_t> cur_label;    // Morton order: original source index of cur[i] (labels of the reciprocal tree)
  bool have_src_normals = false;
  DevBuf<unsigned char> src_raw; // the caller's records verbatim (kept when stride != 16) for output = *input_
  size_t src_stride = 0;
  ptrdiff_t src_normal_off = -1; // byte offset of nx inside a record, when the normals live in the same records
  int64_t total_corr = 0;
  // device state
  DevBuf<Pending> pending;
  DevBuf<SolveOut> solve_out;
  DevBuf<unsigned long long> skip_count;
  int64_t total_skipped = 0;
  int searches = 0;             // search launches since set_source (the first one has no seeds)
  std::vector<pclb200_rejector> rejectors;  // applied in order after every search (icp.hpp:187-201)
  bool track_next = false;      // run the next search with lower-bound tracking / skip test (set per iteration)
  Reducer red;
  // host state (Scalar-typed values are kept in double; float mode rounds after every operation)
  double final_T[16], last_T[16];
  int iterations = 0;
  int state = PCLB200_CONV_NOT_CONVERGED;
  bool converged = false;
  int64_t n_corr = 0;
  double mse = 0.0;
  // DefaultConvergenceCriteria state
  double prev_mse = std::numeric_limits<double>::max();
  int iterations_similar = 0;
};

static void set_identity(double* T)
{
  for (int i = 0; i < 16; ++i)
    T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

static int transform_mode(const pclb200_icp_params& P)
{
  if (!P.with_normals_transform)
    return 0;
  return P.scalar_is_double ? 2 : 1;
}

static void upload_pending(Icp& s, const double* T, int apply)
{
  Pending h;
  for (int i = 0; i < 12; ++i) {
    h.f[i] = (float)T[i];
    h.d[i] = T[i];
  }
  h.apply = apply;
  h.mode = transform_mode(s.P);
  PCLB_CUDA(cudaMemcpyAsync(s.pending.p, &h, sizeof(h), cudaMemcpyHostToDevice, s.ctx->stream));
  PCLB_CUDA(cudaStreamSynchronize(s.ctx->stream));  // h is a stack object
}

Icp* icp_create(Ctx& c, const pclb200_icp_params& P)
{
  std::unique_ptr<Icp> s(new Icp());
  s->ctx = &c;
  s->P = P;
  s->pending.alloc(1, c.stream);
  s->solve_out.alloc(1, c.stream);
  s->ctrl.alloc(1, c.stream);
  s->skip_count.alloc(1, c.stream);
  PCLB_CUDA(cudaMemsetAsync(s->skip_count.p, 0, sizeof(unsigned long long), c.stream));
  s->red.init(c, (unsigned)c.sm_count * 16);
  set_identity(s->final_T);
  set_identity(s->last_T);
  upload_pending(*s, s->final_T, 0);
  return s.release();
}

void icp_destroy(Icp* s) { delete s; }

void icp_set_params(Icp& s, const pclb200_icp_params& P) { s.P = P; }

void icp_set_rejectors(Icp& s, const pclb200_rejector* list, int n)
{
  s.rejectors.assign(list, list + (n > 0 ? n : 0));
  for (const auto& r : s.rejectors)
    PCLB_REQUIRE(r.kind >= PCLB200_REJ_DISTANCE && r.kind <= PCLB200_REJ_SURFACE_NORMAL, PCLB200_ERR_INVALID,
                 "unknown rejector kind");
}

static bool has_surface_normal_rejector(const Icp& s)
{
  for (const auto& r : s.rejectors)
    if (r.kind == PCLB200_REJ_SURFACE_NORMAL)
      return true;
  return false;
}

// the session keeps a rotated copy of the source normals only for the components that read them every iteration
static bool needs_cur_normals(const Icp& s)
{
  return s.P.estimator == PCLB200_EST_SYMMETRIC_POINT_TO_PLANE_LLS || s.P.correspondence_kind != PCLB200_CORR_NEAREST ||
         has_surface_normal_rejector(s);
}

static void run_rejectors(Icp& s)
{
  Ctx& c = *s.ctx;
  cudaStream_t st = c.stream;
  const size_t n = s.n_q;
  DevBuf<float> d2;
  DevBuf<int> mt, acc;
  DevBuf<unsigned> tie;
  d2.alloc(n, st);
  mt.alloc(n, st);
  acc.alloc(n, st);
  tie.alloc(n, st);
  k_match_to_arrays<<<grid_for(n, 256), 256, 0, st>>>(s.cur.p, s.match.p, n, d2.p, mt.p, tie.p, acc.p);
  ++c.launches;
  RejectArrays a;
  a.n = n;
  a.d2 = d2.p;
  a.match = mt.p;
  a.tie = tie.p;
  a.acc = acc.p;
  for (const auto& r : s.rejectors) {
    if (r.kind == PCLB200_REJ_SURFACE_NORMAL) {
      PCLB_REQUIRE(s.cur_normals.p && s.tgt_normals.p, PCLB200_ERR_INVALID,
                   "icp: the surface-normal rejector needs source and target normals (set the rejectors before the source)");
      k_reject_surface_normal<<<grid_for(n, 256), 256, 0, st>>>(s.cur_normals.p, s.tgt_normals.p, mt.p, n, r.p, acc.p);
      ++c.launches;
      continue;
    }
    apply_rejector(c, r, a, nullptr, nullptr, nullptr, nullptr);
  }
  k_arrays_to_match<<<grid_for(n, 256), 256, 0, st>>>(acc.p, n, s.match.p);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
}

__global__ void k_permute_normals(const float4* __restrict__ nrm_orig, const float4* __restrict__ pts, size_t n_padded,
                                  float4* __restrict__ out)
{
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n_padded)
    return;
  const int oi = __float_as_int(pts[j].w);
  const float qn = __int_as_float(0x7fc00000);
  out[j] = oi == kSentinelIndex ? make_float4(qn, qn, qn, 0.f) : nrm_orig[oi];
}

void icp_set_target(Icp& s, const Index* idx, const void* tgt_normals, size_t stride_n)
{
  Ctx& c = *s.ctx;
  s.tgt = idx;
  s.tgt_normals.release();
  if (tgt_normals) {
    DevBuf<float4> dense;
    dense.alloc(idx->n_cloud, c.stream);
    load_vec3_as_float4(c, tgt_normals, idx->n_cloud, stride_n, dense.p, c.stream);
    s.tgt_normals.alloc(idx->pts.n, c.stream);
    k_permute_normals<<<grid_for(idx->pts.n, 256), 256, 0, c.stream>>>(dense.p, idx->pts.p, idx->pts.n,
                                                                      s.tgt_normals.p);
    ++c.launches;
    PCLB_CUDA(cudaGetLastError());
  }
}

__global__ void k_cur_labels(const float4* __restrict__ cur, const int32_t* __restrict__ src_orig, size_t n,
                             int32_t* __restrict__ label)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int slot = __float_as_int(cur[i].w);
  label[i] = src_orig ? src_orig[slot] : slot;
}

__global__ void k_gather_cur_normals(const float4* __restrict__ cur, const float4* __restrict__ nrm_all,
                                     const int32_t* __restrict__ src_orig, size_t n, float4* __restrict__ out)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int slot = __float_as_int(cur[i].w);
  out[i] = nrm_all[src_orig ? src_orig[slot] : slot];
}

__global__ void k_gather_subset(const float4* __restrict__ all, const int32_t* __restrict__ sub, size_t n,
                                float4* __restrict__ out)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = all[sub[i]];
}

void icp_reset_state(Icp& s, const double* guess)
{
  set_identity(s.last_T);
  if (guess) {
    for (int i = 0; i < 16; ++i)
      s.final_T[i] = s.P.scalar_is_double ? guess[i] : (double)(float)guess[i];
  }
  else
    set_identity(s.final_T);
  s.iterations = 0;
  s.state = PCLB200_CONV_NOT_CONVERGED;
  s.converged = false;
  s.n_corr = 0;
  s.track_next = false;
  s.lb_valid = false;
  s.total_skipped = 0;
  s.searches = 0;
  PCLB_CUDA(cudaMemsetAsync(s.skip_count.p, 0, sizeof(unsigned long long), s.ctx->stream));
  s.total_corr = 0;
  s.mse = 0.0;
  s.prev_mse = std::numeric_limits<double>::max();
  s.iterations_similar = 0;
}

void icp_set_source(Icp& s, const void* src, size_t n, size_t stride, const void* src_normals, size_t stride_n,
                    const int32_t* indices, size_t n_idx, const double* guess)
{
  Ctx& c = *s.ctx;
  cudaStream_t st = c.stream;
  PCLB_REQUIRE(s.tgt != nullptr, PCLB200_ERR_INVALID, "icp: no target set (registration.hpp:77-81)");
  PCLB_REQUIRE(src != nullptr && n > 0, PCLB200_ERR_INVALID, "icp: empty source");
  s.n_all = n;
  s.src_all.alloc(n, st);
  s.src_stride = stride;
  s.src_raw.release();
  s.src_normal_off = -1;
  if (src_normals) {
    const ptrdiff_t off = static_cast<const unsigned char*>(src_normals) - static_cast<const unsigned char*>(src);
    if (stride_n == stride && off > 0 && (size_t)off + 12 <= stride)
      s.src_normal_off = off;
  }
  const void* src_dev = src;
  if (stride != 16) {
    // one contiguous H2D of the records; every field other than xyz/normals is carried through to the output
    PCLB_REQUIRE(stride >= 12 && stride % 4 == 0, PCLB200_ERR_INVALID, "stride must be a multiple of 4 and >= 12");
    s.src_raw.alloc(n * stride, st);
    PCLB_CUDA(cudaMemcpyAsync(s.src_raw.p, src, n * stride,
                              is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    src_dev = s.src_raw.p;
  }
  load_xyz_as_float4(c, src_dev, n, stride, nullptr, 0, s.src_all.p, st);
  s.have_src_normals = src_normals != nullptr;
  if (src_normals) {
    s.src_normals.alloc(n, st);
    if (s.src_normal_off >= 0 && s.src_raw.p)
      load_vec3_as_float4(c, s.src_raw.p + s.src_normal_off, n, stride, s.src_normals.p, st);
    else
      load_vec3_as_float4(c, src_normals, n, stride_n, s.src_normals.p, st);
  }
  else
    s.src_normals.release();
  const float4* d_q = s.src_all.p;
  DevBuf<float4> sub;
  s.n_q = n;
  s.src_orig.release();
  if (indices) {
    s.n_q = n_idx;
    s.src_orig.alloc(n_idx, st);
    PCLB_CUDA(cudaMemcpyAsync(s.src_orig.p, indices, n_idx * sizeof(int32_t),
                              is_device_ptr(indices) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    sub.alloc(n_idx, st);
    if (n_idx) {
      k_gather_subset<<<grid_for(n_idx, 256), 256, 0, st>>>(s.src_all.p, s.src_orig.p, n_idx, sub.p);
      ++c.launches;
    }
    d_q = sub.p;
  }
  QueryBatch qb;
  {
    ProfScope ps(c, "query_sort");
    make_query_batch(c, *s.tgt, d_q, s.n_q, qb);
  }
  s.cur = std::move(qb.q);
  s.cur_normals.release();
  if (needs_cur_normals(s)) {
    PCLB_REQUIRE(s.have_src_normals, PCLB200_ERR_INVALID,
                 "icp: the symmetric objective, the normal-based correspondence estimators and the surface-normal "
                 "rejector need source normals");
    s.cur_normals.alloc(s.n_q, st);
    if (s.n_q) {
      k_gather_cur_normals<<<grid_for(s.n_q, 256), 256, 0, st>>>(s.cur.p, s.src_normals.p, s.src_orig.p, s.n_q,
                                                                 s.cur_normals.p);
      ++c.launches;
    }
  }
  s.match.alloc(s.n_q, st);
  PCLB_CUDA(cudaMemsetAsync(s.match.p, 0xff, s.n_q * sizeof(Match), st));  // pos = -1: no seed yet
  s.lb.alloc(s.n_q, st);
  s.lb_valid = false;
  s.cur_label.alloc(s.n_q, st);
  if (s.n_q) {
    k_cur_labels<<<grid_for(s.n_q, 256), 256, 0, st>>>(s.cur.p, s.src_orig.p, s.n_q, s.cur_label.p);
    ++c.launches;
  }
  icp_reset_state(s, guess);
  // icp.hpp:125-134: a non-identity guess is applied once, before the first search
  bool guess_identity = true;
  if (guess)
    for (int i = 0; i < 16; ++i)
      if (s.final_T[i] != ((i % 5 == 0) ? 1.0 : 0.0))
        guess_identity = false;
  if (!guess_identity) {
    upload_pending(s, s.final_T, 1);
    k_apply_pending<<<persistent_grid(c, s.n_q, 256, 8), 256, 0, st>>>(s.cur.p, s.n_q, s.pending.p, s.cur_normals.p);
    ++c.launches;
  }
  upload_pending(s, s.last_T, 0);
  PCLB_CUDA(cudaGetLastError());
}

static void check_device_error(Ctx& c)
{
  int h = 0;
  PCLB_CUDA(cudaMemcpyAsync(&h, c.d_error, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  PCLB_CUDA(cudaStreamSynchronize(c.stream));
  if (h) {
    PCLB_CUDA(cudaMemsetAsync(c.d_error, 0, sizeof(int), c.stream));
    throw Error(PCLB200_ERR_INTERNAL, "LBVH traversal stack overflow (tree deeper than the per-query stack)");
  }
}

// DefaultConvergenceCriteria::hasConverged — impl/default_convergence_criteria.hpp:49-140.
// T is the current T_k in Scalar (stored as doubles; float mode: exactly representable floats).
template <typename S>
static bool has_converged(Icp& s, const double* Td)
{
  const pclb200_icp_params& P = s.P;
  if (s.state != PCLB200_CONV_NOT_CONVERGED) {
    s.iterations_similar = 0;
    s.state = PCLB200_CONV_NOT_CONVERGED;
  }
  bool is_similar = false;
  if (s.iterations >= P.max_iterations) {
    if (!P.failure_after_max_iter) {
      s.state = PCLB200_CONV_ITERATIONS;
      return true;
    }
    s.state = PCLB200_CONV_FAILURE_AFTER_MAX_ITERATIONS;
  }
  S T[16];
  for (int i = 0; i < 16; ++i)
    T[i] = static_cast<S>(Td[i]);
  const double rotation_threshold = P.transformation_rotation_epsilon > 0 ? P.transformation_rotation_epsilon : 0.99999;
  const double translation_threshold = P.transformation_epsilon;
  const double mse_threshold_relative = P.euclidean_fitness_epsilon;
  const double mse_threshold_absolute = P.mse_threshold_absolute;
  double cos_angle = 0.5 * (T[0] + T[5] + T[10] - 1);
  double translation_sqr = T[3] * T[3] + T[7] * T[7] + T[11] * T[11];
  if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) {
    if (s.iterations_similar >= P.max_iterations_similar_transforms) {
      s.state = PCLB200_CONV_TRANSFORM;
      return true;
    }
    is_similar = true;
  }
  const double cur_mse = s.mse;
  if (std::abs(cur_mse - s.prev_mse) < mse_threshold_absolute) {
    if (s.iterations_similar >= P.max_iterations_similar_transforms) {
      s.state = PCLB200_CONV_ABS_MSE;
      return true;
    }
    is_similar = true;
  }
  if (std::abs(cur_mse - s.prev_mse) / s.prev_mse < mse_threshold_relative) {
    if (s.iterations_similar >= P.max_iterations_similar_transforms) {
      s.state = PCLB200_CONV_REL_MSE;
      return true;
    }
    is_similar = true;
  }
  if (is_similar)
    ++s.iterations_similar;
  else
    s.iterations_similar = 0;
  s.prev_mse = cur_mse;
  return false;
}

static void fill_stats(const Icp& s, pclb200_icp_stats* st)
{
  if (!st)
    return;
  st->converged = s.converged ? 1 : 0;
  st->state = s.state;
  st->iterations = s.iterations;
  st->reserved = 0;
  st->n_correspondences = s.n_corr;
  st->total_correspondences = s.total_corr;
  st->total_skipped_walks = s.total_skipped;
  st->mse = s.mse;
  for (int i = 0; i < 16; ++i) {
    st->final_transformation[i] = s.final_T[i];
    st->last_transformation[i] = s.last_T[i];
  }
}

void icp_iterate(Icp& s, int max_steps, pclb200_icp_stats* stats)
{
  Ctx& c = *s.ctx;
  cudaStream_t st = c.stream;
  PCLB_REQUIRE(s.tgt && s.cur.p, PCLB200_ERR_INVALID, "icp: set_target and set_source must precede iterate");
  PCLB_REQUIRE(s.P.estimator == PCLB200_EST_SVD || s.tgt_normals.p, PCLB200_ERR_INVALID,
               "icp: point-to-plane needs target normals");
  PCLB_REQUIRE(s.P.estimator != PCLB200_EST_SYMMETRIC_POINT_TO_PLANE_LLS || s.cur_normals.p, PCLB200_ERR_INVALID,
               "icp: the symmetric objective needs source normals (choose the estimator before set_source)");
  // a sharded source sees only its own shard in the reciprocal tree: a different answer, not a smaller one
  PCLB_REQUIRE(!(comm_active(c) && s.P.use_reciprocal), PCLB200_ERR_INVALID,
               "icp: reciprocal correspondences are not available with a sharded source (multi-GPU communicator)");
  const Index& T = *s.tgt;
  SolveOut* h_out = reinterpret_cast<SolveOut*>(c.pinned);
  int steps = 0;
  auto base_args = [&]() {
    IterArgs a;
    memset(&a, 0, sizeof(a));
    a.nodes = T.nodes.p;
    a.pts = T.pts.p;
    a.root = T.root;
    a.tgt_normals = s.tgt_normals.p;
    a.cur = s.cur.p;
    a.n = s.n_q;
    a.pending = s.pending.p;
    a.gate = gate_from_max_dist(s.P.max_correspondence_distance);
    a.ox = 0.5f * (T.lo[0] + T.hi[0]);
    a.oy = 0.5f * (T.lo[1] + T.hi[1]);
    a.oz = 0.5f * (T.lo[2] + T.hi[2]);
    a.partials = s.red.partials.p;
    a.partials2 = s.red.partials2.p;
    a.counter = s.red.counter.p;
    a.accum = s.red.accum.p;
    a.d_error = c.d_error;
    a.src_orig = s.src_orig.p;
    a.skip_count = s.skip_count.p;
    a.cur_normals = s.cur_normals.p;
    a.enforce_same_dir = s.P.enforce_same_direction_normals;
    a.cells = tree_view(T).cells;
    a.peer.nranks = 0;
    a.seq = 0;
    return a;
  };
  auto raise_device_error = [&](int h_err) {
    if (!h_err)
      return;
    PCLB_CUDA(cudaMemsetAsync(c.d_error, 0, sizeof(int), st));
    if (h_err == 2)
      throw Error(PCLB200_ERR_NCCL, "fused cross-GPU reduce timed out waiting for a peer rank");
    throw Error(PCLB200_ERR_INTERNAL, "LBVH traversal stack overflow (tree deeper than the per-query stack)");
  };
  // ---- enqueue-ahead path ------------------------------------------------------------------------------------------
  // The plain loop (nearest-neighbour correspondences, no rejector chain, no reciprocal tree) needs the host for nothing
  // between iterations: k_solve leaves T_k in device memory for the next search, composes the final transform, counts
  // the iteration and evaluates the convergence criteria (LoopCtrl).  So every iteration this call may run is enqueued
  // at once and the stream is synchronised ONCE; when a criterion fires early the kernels still queued find `done` set
  // and return.  With AUTO tracking both flavours of the search kernel are queued and the device picks one.
  const bool needs_host_between = s.P.correspondence_kind != PCLB200_CORR_NEAREST || !s.rejectors.empty() ||
                                  s.P.use_reciprocal || s.P.estimator == PCLB200_EST_SYMMETRIC_POINT_TO_PLANE_LLS ||
                                  (comm_active(c) && !comm_peer_fused(c));
  if (!needs_host_between && s.state == PCLB200_CONV_NOT_CONVERGED && max_steps > 0) {
    const int left = std::max(1, s.P.max_iterations - s.iterations);  // the do-while always runs once
    const int n_enq = std::min(max_steps, left);
    LoopCtrl* h_ctrl = reinterpret_cast<LoopCtrl*>(static_cast<unsigned char*>(c.pinned) + 1024);
    int* h_err = reinterpret_cast<int*>(static_cast<unsigned char*>(c.pinned) + 1024 + sizeof(LoopCtrl));
    memset(h_ctrl, 0, sizeof(LoopCtrl));
    h_ctrl->iterations = s.iterations;
    h_ctrl->state = s.state;
    h_ctrl->iterations_similar = s.iterations_similar;
    h_ctrl->track_next = s.P.track_mode == PCLB200_TRACK_ON ? 1 : (s.P.track_mode == PCLB200_TRACK_OFF ? 0 : (s.track_next ? 1 : 0));
    h_ctrl->lb_valid = s.lb_valid ? 1 : 0;
    h_ctrl->prev_mse = s.prev_mse;
    h_ctrl->total_corr = s.total_corr;
    for (int i = 0; i < 16; ++i) {
      h_ctrl->final_T[i] = s.final_T[i];
      h_ctrl->last_T[i] = s.last_T[i];
    }
    PCLB_CUDA(cudaMemcpyAsync(s.ctrl.p, h_ctrl, sizeof(LoopCtrl), cudaMemcpyHostToDevice, st));
    CritParams crit;
    crit.max_iterations = s.P.max_iterations;
    crit.failure_after_max_iter = s.P.failure_after_max_iter;
    crit.max_iterations_similar = s.P.max_iterations_similar_transforms;
    crit.scalar_is_double = s.P.scalar_is_double;
    crit.track_mode = s.P.track_mode;
    crit.rot_eps = s.P.transformation_rotation_epsilon;
    crit.trans_eps = s.P.transformation_epsilon;
    crit.rel_mse = s.P.euclidean_fitness_epsilon;
    crit.abs_mse = s.P.mse_threshold_absolute;
    {
      double rmax = 0.0;
      for (int r = 0; r < 3; ++r) {
        const double m = std::max(std::fabs((double)T.lo[r]), std::fabs((double)T.hi[r]));
        rmax += m * m;
      }
      crit.rmax = std::sqrt(rmax);
    }
    const unsigned wgrid = persistent_grid(c, s.n_q, 256, 18);
    const unsigned agrid = std::min(persistent_grid(c, s.n_q, 256, kAccumBlocksPerSM), s.red.max_blocks);
    for (int k = 0; k < n_enq; ++k) {
      IterArgs a = base_args();
      a.ctrl = s.ctrl.p;
      comm_peer_view(c, &a.peer, &a.seq);
      {
        ProfScope ps(c, "icp_search");
        if (s.P.track_mode == PCLB200_TRACK_AUTO) {
          a.track_sel = 1;
          k_search<false, false><<<wgrid, 256, 0, st>>>(a, s.match.p, s.lb.p);
          k_search<false, true><<<wgrid, 256, 0, st>>>(a, s.match.p, s.lb.p);
          c.launches += 2;
        }
        else {
          if (s.P.track_mode == PCLB200_TRACK_ON)
            k_search<false, true><<<wgrid, 256, 0, st>>>(a, s.match.p, s.lb.p);
          else
            k_search<false, false><<<wgrid, 256, 0, st>>>(a, s.match.p, s.lb.p);
          ++c.launches;
        }
      }
      {
        ProfScope ps(c, "icp_accum");
        if (s.P.estimator == PCLB200_EST_SVD)
          k_accum_dmma<PCLB200_EST_SVD><<<agrid, 256, 0, st>>>(a, s.match.p);
        else
          k_accum_dmma<PCLB200_EST_POINT_TO_PLANE_LLS><<<agrid, 256, 0, st>>>(a, s.match.p);
        ++c.launches;
      }
      {
        ProfScope ps(c, "solve");
        k_solve<<<1, 32, 0, st>>>(s.red.accum.p, s.P.estimator, s.P.scalar_is_double, transform_mode(s.P), (double)a.ox,
                                  (double)a.oy, (double)a.oz, 3, s.pending.p, s.solve_out.p, s.P.svd_no_umeyama, s.ctrl.p,
                                  crit);
        ++c.launches;
      }
    }
    PCLB_CUDA(cudaGetLastError());
    PCLB_CUDA(cudaMemcpyAsync(h_ctrl, s.ctrl.p, sizeof(LoopCtrl), cudaMemcpyDeviceToHost, st));
    PCLB_CUDA(cudaMemcpyAsync(h_err, c.d_error, sizeof(int), cudaMemcpyDeviceToHost, st));
    PCLB_CUDA(cudaStreamSynchronize(st));
    raise_device_error(*h_err);
    steps = h_ctrl->n_run;
    s.iterations = h_ctrl->iterations;
    s.state = h_ctrl->state;
    s.converged = h_ctrl->converged != 0;
    s.iterations_similar = h_ctrl->iterations_similar;
    s.prev_mse = h_ctrl->prev_mse;
    s.track_next = h_ctrl->track_next != 0;
    s.lb_valid = h_ctrl->lb_valid != 0;
    if (steps > 0) {
      s.n_corr = (int64_t)h_ctrl->n_corr;
      s.mse = h_ctrl->mse;
    }
    s.total_corr = h_ctrl->total_corr;
    for (int i = 0; i < 16; ++i) {
      s.final_T[i] = h_ctrl->final_T[i];
      s.last_T[i] = h_ctrl->last_T[i];
    }
    s.searches += steps;
  }
  // ---- stage-by-stage path: something between search and solve needs the host (rejector chain with its sorts, the
  // reciprocal tree rebuilt every iteration, the normal-based estimators' k-NN rows, a collective launched by NCCL) ----
  // the reference's do-while runs at least once per align(); a caller stepping one iteration at a time
  // resumes a NOT_CONVERGED session, and a fresh session (iterations == 0) always enters.
  while (needs_host_between && steps < max_steps && (s.state == PCLB200_CONV_NOT_CONVERGED)) {
    IterArgs a = base_args();
    const bool fused_reduce = comm_peer_view(c, &a.peer, &a.seq);
    const unsigned grid = persistent_grid(c, s.n_q, 256, 8);
    std::unique_ptr<Index> src_index;
    if (s.P.use_reciprocal) {
      // tree_reciprocal_ is rebuilt over the transformed source every iteration
      // (correspondence_estimation.hpp:117-135 via setInputSource at icp.hpp:175)
      k_apply_pending<<<grid, 256, 0, st>>>(s.cur.p, s.n_q, s.pending.p, s.cur_normals.p);
      k_clear_apply<<<1, 1, 0, st>>>(s.pending.p);
      c.launches += 2;
      src_index.reset(build_index_from_device(c, s.cur.p, s.n_q, s.cur_label.p));
      a.s_nodes = src_index->nodes.p;
      a.s_pts = src_index->pts.p;
      a.s_root = src_index->root;
    }
    if (s.P.correspondence_kind != PCLB200_CORR_NEAREST) {
      // setCorrespondenceEstimation(NormalShooting / BackProjection): k-NN candidates, then the normal-based choice
      ProfScope ps(c, "icp_search");
      const int kind = s.P.correspondence_kind;
      PCLB_REQUIRE(kind == PCLB200_CORR_NORMAL_SHOOTING || kind == PCLB200_CORR_BACK_PROJECTION, PCLB200_ERR_INVALID,
                   "icp: unknown correspondence_kind");
      PCLB_REQUIRE(!s.P.use_reciprocal, PCLB200_ERR_INVALID,
                   "icp: reciprocal correspondences are only built for the nearest-neighbour estimator");
      PCLB_REQUIRE(s.cur_normals.p, PCLB200_ERR_INVALID, "icp: the normal-based estimators need source normals");
      PCLB_REQUIRE(kind != PCLB200_CORR_BACK_PROJECTION || s.tgt_normals.p, PCLB200_ERR_INVALID,
                   "icp: back projection needs target normals");
      const int keff = (int)std::min<size_t>((size_t)std::max(s.P.correspondence_k, 0), T.n_valid);
      PCLB_REQUIRE(keff > 0, PCLB200_ERR_INVALID, "icp: correspondence_k must be positive");
      k_apply_pending<<<grid, 256, 0, st>>>(s.cur.p, s.n_q, s.pending.p, s.cur_normals.p);
      k_clear_apply<<<1, 1, 0, st>>>(s.pending.p);
      c.launches += 2;
      DevBuf<int32_t> nn_idx;
      DevBuf<float> nn_d2;
      nn_idx.alloc(s.n_q * (size_t)keff, st);
      nn_d2.alloc(s.n_q * (size_t)keff, st);
      launch_knn(c, T, s.cur.p, s.n_q, keff, std::numeric_limits<float>::infinity(), nn_idx.p, nn_d2.p);
      ensure_pos_of_orig(c, const_cast<Index&>(T));
      k_select_match<<<grid_for(s.n_q, 128), 128, 0, st>>>(s.cur.p, s.cur_normals.p, s.n_q, kind, keff, nn_idx.p, nn_d2.p,
                                                          T.pts.p, T.pos_of_orig.p, s.tgt_normals.p,
                                                          s.P.max_correspondence_distance, s.match.p);
      ++c.launches;
      PCLB_CUDA(cudaGetLastError());
    }
    else {
      ProfScope ps(c, "icp_search");
      ++s.searches;
      // Temporal coherence (still_nearest) pays once the cloud has almost stopped moving: tracking the lower bounds makes
      // a walk look at a wider ball, so it is switched on when the last increment displaced no point by more than half
      // the RMS correspondence distance (or always / never: pclb200_icp_params::track_mode), and the skip test then
      // fires from the following iteration on.
      bool track = s.P.track_mode == PCLB200_TRACK_ON || (s.P.track_mode == PCLB200_TRACK_AUTO && s.track_next);
      if (s.P.use_reciprocal)
        track = false;
      if (track && !s.lb_valid)  // bounds left by an earlier TRACK phase say nothing about the current matches
        PCLB_CUDA(cudaMemsetAsync(s.lb.p, 0, s.n_q * sizeof(float), st));
      s.lb_valid = track;
      const unsigned wgrid = persistent_grid(c, s.n_q, 256, 18);
      if (s.P.use_reciprocal)
        k_search<true, false><<<wgrid, 256, 0, st>>>(a, s.match.p, s.lb.p);
      else if (track)
        k_search<false, true><<<wgrid, 256, 0, st>>>(a, s.match.p, s.lb.p);
      else
        k_search<false, false><<<wgrid, 256, 0, st>>>(a, s.match.p, s.lb.p);
      ++c.launches;
    }
    if (!s.rejectors.empty()) {
      PCLB_REQUIRE(!comm_active(c), PCLB200_ERR_INVALID,
                   "correspondence rejectors need the whole correspondence set: not available with a sharded source");
      ProfScope ps(c, "icp_reject");
      run_rejectors(s);
    }
    {
      ProfScope ps(c, "icp_accum");
      const unsigned agrid = std::min(persistent_grid(c, s.n_q, 256, kAccumBlocksPerSM), s.red.max_blocks);
      if (s.P.estimator == PCLB200_EST_SVD)
        k_accum_dmma<PCLB200_EST_SVD><<<agrid, 256, 0, st>>>(a, s.match.p);
      else if (s.P.estimator == PCLB200_EST_POINT_TO_PLANE_LLS)
        k_accum_dmma<PCLB200_EST_POINT_TO_PLANE_LLS><<<agrid, 256, 0, st>>>(a, s.match.p);
      else
        k_accum<PCLB200_EST_SYMMETRIC_POINT_TO_PLANE_LLS><<<grid, 256, 0, st>>>(a, s.match.p);
      ++c.launches;
    }
    PCLB_CUDA(cudaGetLastError());
    if (!fused_reduce) {
      ProfScope ps(c, "allreduce");
      comm_allreduce_sum(c, s.red.accum.p, kAccum);
    }
    {
      ProfScope ps(c, "solve");
      k_solve<<<1, 32, 0, st>>>(s.red.accum.p, s.P.estimator, s.P.scalar_is_double, transform_mode(s.P), (double)a.ox,
                                (double)a.oy, (double)a.oz, 3, s.pending.p, s.solve_out.p, s.P.svd_no_umeyama, nullptr,
                                CritParams{});
    }
    ++c.launches;
    PCLB_CUDA(cudaMemcpyAsync(h_out, s.solve_out.p, sizeof(SolveOut), cudaMemcpyDeviceToHost, st));
    int h_err = 0;
    PCLB_CUDA(cudaMemcpyAsync(&h_err, c.d_error, sizeof(int), cudaMemcpyDeviceToHost, st));
    PCLB_CUDA(cudaStreamSynchronize(st));
    raise_device_error(h_err);
    ++steps;
    s.n_corr = (int64_t)h_out->n;
    s.total_corr += s.n_corr;
    s.mse = h_out->n > 0 ? h_out->sum_d / h_out->n : 0.0;
    if (!h_out->ok) {  // icp.hpp:204-213
      s.state = PCLB200_CONV_NO_CORRESPONDENCES;
      s.converged = false;
      break;
    }
    for (int i = 0; i < 16; ++i)
      s.last_T[i] = h_out->T[i];
    if (s.P.scalar_is_double)
      mat4_mul<double>(s.last_T, s.final_T, s.final_T);
    else {
      float A[16], B[16], C[16];
      for (int i = 0; i < 16; ++i) {
        A[i] = (float)s.last_T[i];
        B[i] = (float)s.final_T[i];
      }
      mat4_mul<float>(A, B, C);
      for (int i = 0; i < 16; ++i)
        s.final_T[i] = C[i];
    }
    ++s.iterations;
    s.converged = s.P.scalar_is_double ? has_converged<double>(s, s.last_T) : has_converged<float>(s, s.last_T);
    {
      // upper bound of the displacement T_k causes anywhere near the target: |R - I|_F * r_max + |t|
      double rf = 0.0, tn = 0.0, rmax = 0.0;
      for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < 3; ++cc) {
          const double d = s.last_T[4 * r + cc] - (r == cc ? 1.0 : 0.0);
          rf += d * d;
        }
        tn += s.last_T[4 * r + 3] * s.last_T[4 * r + 3];
        const double m = std::max(std::fabs((double)T.lo[r]), std::fabs((double)T.hi[r]));
        rmax += m * m;
      }
      const double disp = std::sqrt(rf) * std::sqrt(rmax) + std::sqrt(tn);
      s.track_next = disp < 0.5 * std::sqrt(std::max(s.mse, 0.0));
    }
  }
  if (stats) {
    unsigned long long h = 0;
    PCLB_CUDA(cudaMemcpyAsync(&h, s.skip_count.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    PCLB_CUDA(cudaStreamSynchronize(st));
    s.total_skipped = (int64_t)h;
  }
  fill_stats(s, stats);
}

struct CorrValid {
  __host__ __device__ bool operator()(const pclb200_corr& c) const { return c.index_match >= 0; }
};

// correspondences of the last evaluated iteration (Registration::correspondences_, what PCL hands to the
// visualisation callback at icp.hpp:228-236), ordered by position in the source index list
__global__ void k_matches_to_corr(const float4* __restrict__ cur, const Match* __restrict__ match, size_t n,
                                  const float4* __restrict__ tgt_pts, const int32_t* __restrict__ src_orig,
                                  pclb200_corr* __restrict__ by_slot)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int slot = __float_as_int(cur[i].w);
  const Match m = match[i];
  pclb200_corr r;
  r.index_query = src_orig ? src_orig[slot] : slot;
  const bool acc = match_accepted(m);
  r.index_match = acc ? __float_as_int(tgt_pts[m.pos].w) : -1;
  r.distance = acc ? m.d2 : 0.f;
  by_slot[slot] = r;
}

size_t icp_get_correspondences(Icp& s, pclb200_corr* out)
{
  Ctx& c = *s.ctx;
  cudaStream_t st = c.stream;
  PCLB_REQUIRE(s.tgt && s.cur.p && s.match.p, PCLB200_ERR_INVALID, "icp: no iteration has run");
  if (!s.n_q || s.iterations == 0)
    return 0;
  DevBuf<pclb200_corr> by_slot, compact;
  DevBuf<size_t> d_count;
  by_slot.alloc(s.n_q, st);
  compact.alloc(s.n_q, st);
  d_count.alloc(1, st);
  k_matches_to_corr<<<grid_for(s.n_q, 256), 256, 0, st>>>(s.cur.p, s.match.p, s.n_q, s.tgt->pts.p, s.src_orig.p, by_slot.p);
  ++c.launches;
  size_t tmp_bytes = 0;
  PCLB_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, by_slot.p, compact.p, d_count.p, (int)s.n_q, CorrValid(), st));
  DevBuf<unsigned char> tmp;
  tmp.alloc(tmp_bytes, st);
  PCLB_CUDA(cub::DeviceSelect::If(tmp.p, tmp_bytes, by_slot.p, compact.p, d_count.p, (int)s.n_q, CorrValid(), st));
  c.launches += 2;
  size_t m = 0;
  PCLB_CUDA(cudaMemcpyAsync(&m, d_count.p, sizeof(size_t), cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
  if (m)
    PCLB_CUDA(cudaMemcpyAsync(out, compact.p, m * sizeof(pclb200_corr),
                              is_device_ptr(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
  return m;
}

// output = *input_ ; transformCloud(*input_, output, final_transformation_) — icp.hpp:265-267
void icp_get_cloud(Icp& s, void* out_pts, size_t stride_out, void* out_normals, size_t stride_n)
{
  Ctx& c = *s.ctx;
  cudaStream_t st = c.stream;
  PCLB_REQUIRE(s.src_all.p, PCLB200_ERR_INVALID, "icp: no source");
  DevBuf<float4> tmp, tmpn;
  tmp.alloc(s.n_all, st);
  PCLB_CUDA(cudaMemcpyAsync(tmp.p, s.src_all.p, s.n_all * sizeof(float4), cudaMemcpyDeviceToDevice, st));
  const bool do_n = out_normals && s.have_src_normals;
  if (do_n) {
    tmpn.alloc(s.n_all, st);
    PCLB_CUDA(cudaMemcpyAsync(tmpn.p, s.src_normals.p, s.n_all * sizeof(float4), cudaMemcpyDeviceToDevice, st));
  }
  DevBuf<Pending> pend;
  pend.alloc(1, st);
  Pending h;
  for (int i = 0; i < 12; ++i) {
    h.f[i] = (float)s.final_T[i];
    h.d[i] = s.final_T[i];
  }
  h.apply = 1;
  h.mode = transform_mode(s.P);
  PCLB_CUDA(cudaMemcpyAsync(pend.p, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  {
    ProfScope ps(c, "transform_out");
    k_apply_pending<<<persistent_grid(c, s.n_all, 256, 8), 256, 0, st>>>(tmp.p, s.n_all, pend.p, do_n ? tmpn.p : nullptr);
  }
  ++c.launches;
  const bool normals_inside =
      do_n && s.src_normal_off >= 0 &&
      static_cast<unsigned char*>(out_normals) - static_cast<unsigned char*>(out_pts) == s.src_normal_off && stride_n == stride_out;
  if (s.src_raw.p && stride_out == s.src_stride && (!do_n || normals_inside)) {
    // output = *input_ (all fields), then xyz / normals overwritten: assembled on the device, one contiguous copy out
    DevBuf<unsigned char> rec;
    unsigned char* d_rec = nullptr;
    const bool out_dev = is_device_ptr(out_pts);
    if (out_dev)
      d_rec = static_cast<unsigned char*>(out_pts);
    else {
      rec.alloc(s.n_all * stride_out, st);
      d_rec = rec.p;
    }
    PCLB_CUDA(cudaMemcpyAsync(d_rec, s.src_raw.p, s.n_all * stride_out, cudaMemcpyDeviceToDevice, st));
    k_scatter_xyz<<<grid_for(s.n_all, 256), 256, 0, st>>>(tmp.p, s.n_all, d_rec, stride_out, 0, 0.f);
    ++c.launches;
    if (do_n) {
      k_scatter_xyz<<<grid_for(s.n_all, 256), 256, 0, st>>>(tmpn.p, s.n_all, d_rec + s.src_normal_off, stride_out, 0, 0.f);
      ++c.launches;
    }
    if (!out_dev)
      PCLB_CUDA(cudaMemcpyAsync(out_pts, d_rec, s.n_all * stride_out, cudaMemcpyDeviceToHost, st));
    PCLB_CUDA(cudaStreamSynchronize(st));
    return;
  }
  auto write_back = [&](const DevBuf<float4>& d, void* out, size_t stride, int write_w, float w) {
    if (stride == 16 && write_w) {
      PCLB_CUDA(cudaMemcpyAsync(out, d.p, s.n_all * 16, is_device_ptr(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
      return;
    }
    if (is_device_ptr(out)) {
      k_scatter_xyz<<<grid_for(s.n_all, 256), 256, 0, st>>>(d.p, s.n_all, (unsigned char*)out, stride, write_w, w);
      ++c.launches;
    }
    else {
      // host strided output: only x,y,z (and w) of each record may be touched
      std::vector<float4> hbuf(s.n_all);
      PCLB_CUDA(cudaMemcpyAsync(hbuf.data(), d.p, s.n_all * 16, cudaMemcpyDeviceToHost, st));
      PCLB_CUDA(cudaStreamSynchronize(st));
      unsigned char* o = static_cast<unsigned char*>(out);
      for (size_t i = 0; i < s.n_all; ++i) {
        float* f = reinterpret_cast<float*>(o + i * stride);
        f[0] = hbuf[i].x; f[1] = hbuf[i].y; f[2] = hbuf[i].z;
        if (write_w)
          f[3] = w;
      }
    }
  };
  // src_all keeps the caller's w (PointXYZ: 1.0 after align(), registration.hpp:213-216)
  write_back(tmp, out_pts, stride_out, stride_out >= 16 ? 1 : 0, 1.0f);
  if (do_n)
    write_back(tmpn, out_normals, stride_n, 0, 0.f);
  PCLB_CUDA(cudaStreamSynchronize(st));
}

// ---- stand-alone estimators ----------------------------------------------------------------------------------
void estimate_pairs(Ctx& c, int est, const void* src, size_t stride_s, const void* tgt, const void* tgt_normals,
                    size_t stride_t, const pclb200_corr* corr, size_t n, int scalar_is_double, double* T_out,
                    const void* src_normals, int enforce_same_dir, int svd_correlation)
{
  cudaStream_t st = c.stream;
  PCLB_REQUIRE(src && tgt && n > 0, PCLB200_ERR_INVALID, "estimate: empty input");
  // the caller's arrays are indexed by corr[].index_*; without their sizes we need the max index
  size_t n_src = n, n_tgt = n;
  DevBuf<pclb200_corr> d_corr;
  if (corr) {
    std::vector<pclb200_corr> hc;
    const pclb200_corr* hcorr = corr;
    if (is_device_ptr(corr)) {
      hc.resize(n);
      PCLB_CUDA(cudaMemcpy(hc.data(), corr, n * sizeof(pclb200_corr), cudaMemcpyDeviceToHost));
      hcorr = hc.data();
    }
    int mq = 0, mm = 0;
    for (size_t i = 0; i < n; ++i) {
      PCLB_REQUIRE(hcorr[i].index_query >= 0 && hcorr[i].index_match >= 0, PCLB200_ERR_INVALID, "negative index");
      mq = std::max(mq, hcorr[i].index_query);
      mm = std::max(mm, hcorr[i].index_match);
    }
    n_src = (size_t)mq + 1;
    n_tgt = (size_t)mm + 1;
    d_corr.alloc(n, st);
    PCLB_CUDA(cudaMemcpyAsync(d_corr.p, hcorr, n * sizeof(pclb200_corr), cudaMemcpyHostToDevice, st));
    PCLB_CUDA(cudaStreamSynchronize(st));
  }
  DevBuf<float4> ds, dt, dn;
  ds.alloc(n_src, st);
  dt.alloc(n_tgt, st);
  load_xyz_as_float4(c, src, n_src, stride_s, nullptr, 0, ds.p, st);
  load_xyz_as_float4(c, tgt, n_tgt, stride_t, nullptr, 0, dt.p, st);
  DevBuf<float4> dsn;
  if (est != PCLB200_EST_SVD) {
    PCLB_REQUIRE(tgt_normals, PCLB200_ERR_INVALID, "point-to-plane needs target normals");
    dn.alloc(n_tgt, st);
    load_vec3_as_float4(c, tgt_normals, n_tgt, stride_t, dn.p, st);
  }
  if (est == PCLB200_EST_SYMMETRIC_POINT_TO_PLANE_LLS) {
    PCLB_REQUIRE(src_normals, PCLB200_ERR_INVALID, "the symmetric objective needs source normals");
    dsn.alloc(n_src, st);
    load_vec3_as_float4(c, src_normals, n_src, stride_s, dsn.p, st);
  }
  Reducer red;
  const unsigned grid = persistent_grid(c, n, 256, 4);
  red.init(c, grid);
  DevBuf<SolveOut> so;
  so.alloc(1, st);
  PairArgs a;
  memset(&a, 0, sizeof(a));
  a.src = ds.p;
  a.tgt = dt.p;
  a.tgt_normals = dn.p;
  a.src_normals = dsn.p;
  a.enforce_same_dir = enforce_same_dir;
  a.corr = d_corr.p;
  a.n = n;
  a.ox = a.oy = a.oz = 0.f;
  a.pub.partials = red.partials.p;
  a.pub.counter = red.counter.p;
  a.pub.accum = red.accum.p;
  if (est == PCLB200_EST_SVD) {
    // shift by the first target point to keep the fp64 sums small
    float4 first;
    PCLB_CUDA(cudaMemcpyAsync(&first, dt.p, sizeof(float4), cudaMemcpyDeviceToHost, st));
    PCLB_CUDA(cudaStreamSynchronize(st));
    if (std::isfinite(first.x) && std::isfinite(first.y) && std::isfinite(first.z)) {
      a.ox = first.x; a.oy = first.y; a.oz = first.z;
    }
    k_accum_pairs<PCLB200_EST_SVD><<<grid, 256, 0, st>>>(a);
  }
  else if (est == PCLB200_EST_POINT_TO_PLANE_LLS)
    k_accum_pairs<PCLB200_EST_POINT_TO_PLANE_LLS><<<grid, 256, 0, st>>>(a);
  else
    k_accum_pairs<PCLB200_EST_SYMMETRIC_POINT_TO_PLANE_LLS><<<grid, 256, 0, st>>>(a);
  k_solve<<<1, 32, 0, st>>>(red.accum.p, est, scalar_is_double, 0, (double)a.ox, (double)a.oy, (double)a.oz, 1, nullptr,
                            so.p, svd_correlation, nullptr, CritParams{});
  c.launches += 2;
  SolveOut h;
  PCLB_CUDA(cudaMemcpyAsync(&h, so.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 16; ++i)
    T_out[i] = h.T[i];
}

// ---- correspondences (materialised) -----------------------------------------------------------------------------
size_t correspondences(Ctx& c, const Index& tgt, const Index* src_index, const void* src, size_t n, size_t stride,
                       const int32_t* indices, size_t n_idx, int is_dense, double max_dist, pclb200_corr* out,
                       const double* pre_T, int pre_mode, const float* gate_override)
{
  (void)is_dense;  // non-finite source points never produce a correspondence on either setting
  cudaStream_t st = c.stream;
  const size_t nq = indices ? n_idx : n;
  if (nq == 0)
    return 0;
  DevBuf<float4> dense;
  dense.alloc(nq, st);
  load_xyz_as_float4(c, src, n, stride, indices, n_idx, dense.p, st);
  DevBuf<Pending> pend;
  if (pre_T) {  // the caller's cloud is searched after a rigid transform (consumers that validate a pose)
    pend.alloc(1, st);
    Pending h;
    for (int i = 0; i < 12; ++i) {
      h.f[i] = (float)pre_T[i];
      h.d[i] = pre_T[i];
    }
    h.apply = 1;
    h.mode = pre_mode;
    PCLB_CUDA(cudaMemcpyAsync(pend.p, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    PCLB_CUDA(cudaStreamSynchronize(st));  // h is a stack object
    k_apply_pending<<<persistent_grid(c, nq, 256, 8), 256, 0, st>>>(dense.p, nq, pend.p, nullptr);
    ++c.launches;
  }
  DevBuf<int32_t> d_ind;
  if (indices) {
    d_ind.alloc(n_idx, st);
    PCLB_CUDA(cudaMemcpyAsync(d_ind.p, indices, n_idx * sizeof(int32_t),
                              is_device_ptr(indices) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  }
  QueryBatch qb;
  make_query_batch(c, tgt, dense.p, nq, qb);
  DevBuf<pclb200_corr> by_slot, compact;
  DevBuf<size_t> d_count;
  by_slot.alloc(nq, st);
  compact.alloc(nq, st);
  d_count.alloc(1, st);
  const float gate = gate_override ? *gate_override : gate_from_max_dist(max_dist);
  if (src_index)
    k_corr<true><<<grid_for(nq, 128), 128, 0, st>>>(tree_view(tgt), qb.q.p, nq, gate,
                                                   src_index->nodes.p, src_index->pts.p, src_index->root, d_ind.p,
                                                   by_slot.p, c.d_error);
  else
    k_corr<false><<<grid_for(nq, 128), 128, 0, st>>>(tree_view(tgt), qb.q.p, nq, gate, nullptr,
                                                    nullptr, 0, d_ind.p, by_slot.p, c.d_error);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  PCLB_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, by_slot.p, compact.p, d_count.p, (int)nq, CorrValid(), st));
  DevBuf<unsigned char> tmp;
  tmp.alloc(tmp_bytes, st);
  PCLB_CUDA(cub::DeviceSelect::If(tmp.p, tmp_bytes, by_slot.p, compact.p, d_count.p, (int)nq, CorrValid(), st));
  c.launches += 2;
  size_t m = 0;
  PCLB_CUDA(cudaMemcpyAsync(&m, d_count.p, sizeof(size_t), cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
  if (m)
    PCLB_CUDA(cudaMemcpyAsync(out, compact.p, m * sizeof(pclb200_corr),
                              is_device_ptr(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  check_device_error(c);
  return m;
}

// ---- fitness score ------------------------------------------------------------------------------------------------
double fitness_score(Ctx& c, const Index& tgt, const void* src, size_t n, size_t stride, const int32_t* indices,
                     size_t n_idx, const double* T, int scalar_is_double, double max_range, int mode_override)
{
  cudaStream_t st = c.stream;
  // registration.hpp:141-144: the index subset is used only when it is a strict subset
  const bool use_sub = indices && n_idx != n;
  const size_t nq = use_sub ? n_idx : n;
  if (!nq)
    return std::numeric_limits<double>::max();
  DevBuf<float4> dense;
  dense.alloc(nq, st);
  load_xyz_as_float4(c, src, n, stride, use_sub ? indices : nullptr, use_sub ? n_idx : 0, dense.p, st);
  DevBuf<Pending> pend;
  pend.alloc(1, st);
  Pending h;
  for (int i = 0; i < 12; ++i) {
    h.f[i] = (float)T[i];
    h.d[i] = scalar_is_double ? T[i] : (double)(float)T[i];
  }
  h.apply = 1;
  h.mode = mode_override >= 0 ? mode_override : (scalar_is_double ? 2 : 1);  // default: pcl::transformPointCloud
  PCLB_CUDA(cudaMemcpyAsync(pend.p, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  k_apply_pending<<<persistent_grid(c, nq, 256, 8), 256, 0, st>>>(dense.p, nq, pend.p, nullptr);
  ++c.launches;
  QueryBatch qb;
  make_query_batch(c, tgt, dense.p, nq, qb);
  Reducer red;
  const unsigned grid = persistent_grid(c, nq, 256, 8);
  red.init(c, grid);
  IterArgs pub;
  memset(&pub, 0, sizeof(pub));
  pub.partials = red.partials.p;
  pub.counter = red.counter.p;
  pub.accum = red.accum.p;
  pub.d_error = c.d_error;
  k_fitness<<<grid, 256, 0, st>>>(tree_view(tgt), qb.q.p, nq, max_range, pub);
  ++c.launches;
  double acc[2];
  PCLB_CUDA(cudaMemcpyAsync(acc, red.accum.p, sizeof(acc), cudaMemcpyDeviceToHost, st));
  check_device_error(c);
  return acc[0] > 0 ? acc[1] / acc[0] : std::numeric_limits<double>::max();
}

}  // namespace pclb200
