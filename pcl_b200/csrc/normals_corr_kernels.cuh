// normals_corr_kernels.cuh — the kernels of normals_corr.cu (normal shooting / back projection over k-NN rows, the stand-alone
// surface-normal rejector), in a header of their own so that tests/host/consumers_host_test.cpp can compile them for the host.
// Included by normals_corr.cu only (the choices themselves, corr_select.cuh, are shared with icp.cu).
#pragma once
#include "corr_select.cuh"

namespace pclb200 {

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

// CorrespondenceRejectorSurfaceNormal, stand-alone: a rejected record is marked index_match = -1 and compacted away
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

}  // namespace pclb200
