// normals_corr.cu — correspondence estimation and rejection that use surface normals.
//
// Replaces CorrespondenceEstimationNormalShooting / CorrespondenceEstimationBackProjection
// (registration/include/pcl/registration/impl/correspondence_estimation_normal_shooting.hpp:66-131,
//  impl/correspondence_estimation_backprojection.hpp:66-118) and CorrespondenceRejectorSurfaceNormal
// (registration/src/correspondence_rejection_surface_normal.cpp:43-66).
// The k candidates of every source point come from the exact k-NN kernels of search.cu (one batch launch), the
// choice among them is one thread per source point over its candidate row (corr_select.cuh).
#include <cub/cub.cuh>

#include <algorithm>
#include <limits>

#include "corr_select.cuh"
#include "internal.cuh"

namespace pclb200 {

static inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

struct CorrHasMatch {
  __host__ __device__ bool operator()(const pclb200_corr& c) const { return c.index_match >= 0; }
};

// one thread per source point (slot order = order of the source index list = order of the output)
__global__ void __launch_bounds__(128)
k_corr_by_normals(const float4* __restrict__ dense, size_t nq, const int32_t* __restrict__ src_orig,
                  const float4* __restrict__ src_nrm, int kind, int k, const int32_t* __restrict__ nn_idx,
                  const float* __restrict__ nn_d2, const float4* __restrict__ tgt_pts,
                  const int32_t* __restrict__ pos_of_orig, const float4* __restrict__ tgt_nrm, double max_dist,
                  pclb200_corr* __restrict__ by_slot)
{
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nq)
    return;
  const float4 p = dense[i];
  const int orig = src_orig ? src_orig[i] : (int)i;
  pclb200_corr r;
  r.index_query = orig;
  r.index_match = -1;
  r.distance = 0.f;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    const float4 n = src_nrm[orig];
    const int32_t* row_idx = nn_idx + i * (size_t)k;
    const float* row_d2 = nn_d2 + i * (size_t)k;
    const int j = select_by_normals<false>(kind, k, row_idx, row_d2, p.x, p.y, p.z, n.x, n.y, n.z, tgt_pts,
                                           pos_of_orig, tgt_nrm, max_dist);
    if (j >= 0) {
      r.index_match = row_idx[j];
      r.distance = row_d2[j];
    }
  }
  by_slot[i] = r;
}

size_t correspondences_by_normals(Ctx& c, Index& tgt, int kind, const void* src, size_t n, size_t stride,
                                  const void* src_normals, size_t stride_sn, const void* tgt_normals, size_t stride_tn,
                                  const int32_t* indices, size_t n_idx, int k, double max_dist, pclb200_corr* out)
{
  cudaStream_t st = c.stream;
  const size_t nq = indices ? n_idx : n;
  const int keff = (int)std::min<size_t>((size_t)std::max(k, 0), tgt.n_valid);  // kdtree_flann.hpp:241-242
  if (nq == 0 || keff <= 0)
    return 0;
  DevBuf<float4> dense, sn, tn;
  dense.alloc(nq, st);
  load_xyz_as_float4(c, src, n, stride, indices, n_idx, dense.p, st);
  sn.alloc(n, st);
  load_vec3_as_float4(c, src_normals, n, stride_sn, sn.p, st);
  if (kind == PCLB200_CORR_BACK_PROJECTION) {
    tn.alloc(tgt.n_cloud, st);
    load_vec3_as_float4(c, tgt_normals, tgt.n_cloud, stride_tn, tn.p, st);
  }
  DevBuf<int32_t> d_ind;
  if (indices) {
    d_ind.alloc(n_idx, st);
    PCLB_CUDA(cudaMemcpyAsync(d_ind.p, indices, n_idx * sizeof(int32_t),
                              is_device_ptr(indices) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  }
  QueryBatch qb;
  make_query_batch(c, tgt, dense.p, nq, qb);
  DevBuf<int32_t> nn_idx;
  DevBuf<float> nn_d2;
  nn_idx.alloc(nq * (size_t)keff, st);
  nn_d2.alloc(nq * (size_t)keff, st);
  {
    ProfScope ps(c, "knn");
    launch_knn(c, tgt, qb.q.p, nq, keff, std::numeric_limits<float>::infinity(), nn_idx.p, nn_d2.p);
  }
  ensure_pos_of_orig(c, tgt);
  DevBuf<pclb200_corr> by_slot, compact;
  DevBuf<size_t> d_count;
  by_slot.alloc(nq, st);
  compact.alloc(nq, st);
  d_count.alloc(1, st);
  k_corr_by_normals<<<grid_for(nq, 128), 128, 0, st>>>(dense.p, nq, d_ind.p, sn.p, kind, keff, nn_idx.p, nn_d2.p,
                                                      tgt.pts.p, tgt.pos_of_orig.p, tn.p, max_dist, by_slot.p);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  PCLB_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, by_slot.p, compact.p, d_count.p, (int)nq, CorrHasMatch(), st));
  DevBuf<unsigned char> tmp;
  tmp.alloc(tmp_bytes, st);
  PCLB_CUDA(cub::DeviceSelect::If(tmp.p, tmp_bytes, by_slot.p, compact.p, d_count.p, (int)nq, CorrHasMatch(), st));
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

// ---- CorrespondenceRejectorSurfaceNormal, stand-alone ---------------------------------------------------------------
__global__ void k_mark_surface_normal(const pclb200_corr* __restrict__ in, size_t n, const float4* __restrict__ sn,
                                      size_t n_src, const float4* __restrict__ tn, size_t n_tgt, double threshold,
                                      pclb200_corr* __restrict__ marked)
{
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  pclb200_corr r = in[i];
  const bool in_range = r.index_query >= 0 && (size_t)r.index_query < n_src && r.index_match >= 0 &&
                        (size_t)r.index_match < n_tgt;
  if (!in_range || !surface_normal_keeps(sn[r.index_query], tn[r.index_match], threshold))
    r.index_match = -1;
  marked[i] = r;
}

size_t reject_surface_normal(Ctx& c, const pclb200_corr* in, size_t n, const void* src_normals, size_t n_src,
                             size_t stride_sn, const void* tgt_normals, size_t n_tgt, size_t stride_tn,
                             double threshold, pclb200_corr* out)
{
  cudaStream_t st = c.stream;
  if (n == 0)
    return 0;
  DevBuf<float4> sn, tn;
  sn.alloc(n_src, st);
  tn.alloc(n_tgt, st);
  load_vec3_as_float4(c, src_normals, n_src, stride_sn, sn.p, st);
  load_vec3_as_float4(c, tgt_normals, n_tgt, stride_tn, tn.p, st);
  DevBuf<pclb200_corr> d_in, marked, compact;
  const pclb200_corr* pin = in;
  if (!is_device_ptr(in)) {
    d_in.alloc(n, st);
    PCLB_CUDA(cudaMemcpyAsync(d_in.p, in, n * sizeof(pclb200_corr), cudaMemcpyHostToDevice, st));
    pin = d_in.p;
  }
  marked.alloc(n, st);
  compact.alloc(n, st);
  DevBuf<size_t> d_count;
  d_count.alloc(1, st);
  k_mark_surface_normal<<<grid_for(n, 256), 256, 0, st>>>(pin, n, sn.p, n_src, tn.p, n_tgt, threshold, marked.p);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  PCLB_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, marked.p, compact.p, d_count.p, (int)n, CorrHasMatch(), st));
  DevBuf<unsigned char> tmp;
  tmp.alloc(tmp_bytes, st);
  PCLB_CUDA(cub::DeviceSelect::If(tmp.p, tmp_bytes, marked.p, compact.p, d_count.p, (int)n, CorrHasMatch(), st));
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

}  // namespace pclb200
