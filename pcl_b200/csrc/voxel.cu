// voxel.cu — pcl::VoxelGrid<PointT>::applyFilter on the device (downsample front-end of configs 2 and 5).
//
// Reference: filters/include/pcl/filters/impl/voxel_grid.hpp:596-814
//   getMinMax3D (:617) -> int32 overflow guard (:620-629) -> min_b/max_b/div_b/divb_mul (:632-644)
//   -> per point ijk = floor(p*inv_leaf) - min_b, idx = ijk . divb_mul (:705-719)
//   -> sort by idx (:724-725) -> runs of equal idx, drop runs < min_points_per_voxel (:737-748)
//   -> per run centroid = float sum / count (:779-812, accumulators.hpp:68-85)
// The reference sorts with an UNSTABLE spreadsort, so its within-voxel summation order is unspecified;
// here (and in the oracle) the sort is stable => the order is ascending position in `indices`, and one
// thread sums its voxel's points in that order, so device and oracle agree bit for bit.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <limits>

#include "internal.cuh"

namespace pclb200 {

static inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

struct MinMaxAcc {
  int lo[3];
  int hi[3];
  unsigned long long count;
};

__device__ __forceinline__ int vf2ord(float f)
{
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
static float vord2f(int i)
{
  int j = i >= 0 ? i : i ^ 0x7fffffff;
  float f;
  memcpy(&f, &j, 4);
  return f;
}

__global__ void k_vg_minmax(const float4* __restrict__ p, size_t n, MinMaxAcc* acc)
{
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  unsigned cnt = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(p + i);
    if (isfinite(v.x) && isfinite(v.y) && isfinite(v.z)) {
      lo[0] = fminf(lo[0], v.x); hi[0] = fmaxf(hi[0], v.x);
      lo[1] = fminf(lo[1], v.y); hi[1] = fmaxf(hi[1], v.y);
      lo[2] = fminf(lo[2], v.z); hi[2] = fmaxf(hi[2], v.z);
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(&acc->lo[d], vf2ord(lo[d]));
      atomicMax(&acc->hi[d], vf2ord(hi[d]));
    }
    atomicAdd(&acc->count, (unsigned long long)cnt);
  }
}

struct VgParams {
  float inv[3];
  int min_b[3];
  int mul[3];
};

__global__ void k_vg_keys(const float4* __restrict__ p, size_t n, VgParams g, unsigned* __restrict__ keys,
                          int32_t* __restrict__ vals)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float4 v = __ldg(p + i);
  unsigned key = 0xffffffffu;  // non-finite points sort to the tail and are cut off
  if (isfinite(v.x) && isfinite(v.y) && isfinite(v.z)) {
    int ijk0 = (int)(floorf(__fmul_rn(v.x, g.inv[0])) - (float)g.min_b[0]);
    int ijk1 = (int)(floorf(__fmul_rn(v.y, g.inv[1])) - (float)g.min_b[1]);
    int ijk2 = (int)(floorf(__fmul_rn(v.z, g.inv[2])) - (float)g.min_b[2]);
    key = (unsigned)(ijk0 * g.mul[0] + ijk1 * g.mul[1] + ijk2 * g.mul[2]);
  }
  keys[i] = key;
  vals[i] = (int32_t)i;
}

__global__ void k_vg_heads(const unsigned* __restrict__ keys, size_t n, unsigned char* __restrict__ head)
{
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j < n)
    head[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}

// keep[m] = run m has >= min_pts points
__global__ void k_vg_keep(const unsigned* __restrict__ starts, size_t n_runs, size_t n_valid, unsigned min_pts,
                          unsigned char* __restrict__ keep)
{
  size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (m >= n_runs)
    return;
  size_t b = starts[m], e = (m + 1 < n_runs) ? starts[m + 1] : n_valid;
  keep[m] = (e - b) >= min_pts ? 1 : 0;
}

struct RunRef {
  unsigned begin, end;
};

__global__ void k_vg_runs(const unsigned* __restrict__ starts, size_t n_runs, size_t n_valid, RunRef* __restrict__ runs)
{
  size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (m >= n_runs)
    return;
  runs[m].begin = starts[m];
  runs[m].end = (m + 1 < n_runs) ? starts[m + 1] : (unsigned)n_valid;
}

__global__ void k_vg_centroids(const float4* __restrict__ p, const int32_t* __restrict__ vals,
                               const RunRef* __restrict__ runs, size_t n_runs, float4* __restrict__ out)
{
  size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (m >= n_runs)
    return;
  const RunRef r = runs[m];
  float cx = 0.f, cy = 0.f, cz = 0.f;
  for (unsigned j = r.begin; j < r.end; ++j) {
    const float4 v = __ldg(p + vals[j]);
    cx = __fadd_rn(cx, v.x);
    cy = __fadd_rn(cy, v.y);
    cz = __fadd_rn(cz, v.z);
  }
  const float fn = (float)(r.end - r.begin);
  out[m] = make_float4(__fdiv_rn(cx, fn), __fdiv_rn(cy, fn), __fdiv_rn(cz, fn), 1.0f);
}

// downsample_all_data_ (voxel_grid.hpp:796-806, CentroidPoint): normals are summed as 4-vectors and normalised
// (AccumulatorNormal, accumulators.hpp:86-116), the curvature is averaged (AccumulatorCurvature, :118-133).
// nc: two float4 per selected point {nx,ny,nz,n4} {curvature,-,-,-}; out: two float4 per voxel, same layout
__global__ void k_vg_load_nc(const unsigned char* __restrict__ src, size_t stride, const int32_t* __restrict__ subset,
                             size_t n, float4* __restrict__ out)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const size_t r = subset ? (size_t)subset[i] : i;
  const float* f = reinterpret_cast<const float*>(src + r * stride);
  out[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
  out[2 * i + 1] = make_float4(f[4], 0.f, 0.f, 0.f);
}

__global__ void k_vg_normals(const float4* __restrict__ nc, const int32_t* __restrict__ vals, const RunRef* __restrict__ runs,
                             size_t n_runs, float4* __restrict__ out)
{
  size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (m >= n_runs)
    return;
  const RunRef r = runs[m];
  float nx = 0.f, ny = 0.f, nz = 0.f, nw = 0.f, cv = 0.f;
  for (unsigned j = r.begin; j < r.end; ++j) {
    const float4 a = __ldg(nc + 2 * (size_t)vals[j]);
    const float4 b = __ldg(nc + 2 * (size_t)vals[j] + 1);
    nx = __fadd_rn(nx, a.x); ny = __fadd_rn(ny, a.y); nz = __fadd_rn(nz, a.z); nw = __fadd_rn(nw, a.w);
    cv = __fadd_rn(cv, b.x);
  }
  const float sq = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fadd_rn(__fmul_rn(nz, nz), __fmul_rn(nw, nw)));
  if (sq > 0.f) {  // Eigen's normalized(): the zero vector stays zero
    const float nrm = __fsqrt_rn(sq);
    nx = __fdiv_rn(nx, nrm); ny = __fdiv_rn(ny, nrm); nz = __fdiv_rn(nz, nrm); nw = __fdiv_rn(nw, nrm);
  }
  out[2 * m] = make_float4(nx, ny, nz, nw);
  out[2 * m + 1] = make_float4(__fdiv_rn(cv, (float)(r.end - r.begin)), 0.f, 0.f, 0.f);
}

struct IsSet {
  __host__ __device__ bool operator()(unsigned char v) const { return v != 0; }
};

size_t voxelgrid(Ctx& c, const void* pts, size_t n, size_t stride, const int32_t* indices, size_t n_idx, int is_dense,
                 const float leaf[3], unsigned min_pts, float* out_xyz1, const void* normals, size_t stride_n,
                 float* out_normal_curv, const float* grid_bounds)
{
  (void)is_dense;  // non-finite points are skipped on either setting (a dense cloud has none)
  cudaStream_t st = c.stream;
  const size_t cnt = indices ? n_idx : n;
  if (cnt == 0)
    return 0;
  PCLB_REQUIRE(leaf[0] > 0 && leaf[1] > 0 && leaf[2] > 0, PCLB200_ERR_INVALID, "leaf size must be positive");
  DevBuf<float4> dense;
  dense.alloc(cnt, st);
  load_xyz_as_float4(c, pts, n, stride, indices, n_idx, dense.p, st);
  // pass 1: min/max
  DevBuf<MinMaxAcc> acc;
  acc.alloc(1, st);
  MinMaxAcc init;
  for (int d = 0; d < 3; ++d) {
    init.lo[d] = 0x7fffffff;
    init.hi[d] = (int)0x80000000;
  }
  init.count = 0;
  PCLB_CUDA(cudaMemcpyAsync(acc.p, &init, sizeof(init), cudaMemcpyHostToDevice, st));
  unsigned g = (unsigned)std::min<size_t>((cnt + 255) / 256, (size_t)c.sm_count * 8);
  k_vg_minmax<<<g, 256, 0, st>>>(dense.p, cnt, acc.p);
  ++c.launches;
  MinMaxAcc h;
  PCLB_CUDA(cudaMemcpyAsync(&h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
  const size_t n_valid = (size_t)h.count;
  if (n_valid == 0)
    return 0;
  float mn[3], mx[3], inv[3];
  for (int d = 0; d < 3; ++d) {
    mn[d] = vord2f(h.lo[d]);
    mx[d] = vord2f(h.hi[d]);
    inv[d] = 1.0f / leaf[d];  // inverse_leaf_size_ = 1 / leaf_size_ (voxel_grid.h:266,282)
  }
  if (grid_bounds) {
    // a spatial tile of a larger cloud: the grid (min_b, div_b) is the WHOLE cloud's, so that the tiles' voxels are the
    // voxels a single VoxelGrid over the whole cloud would form (the caller cuts tiles along voxel boundaries)
    for (int d = 0; d < 3; ++d) {
      PCLB_REQUIRE(grid_bounds[d] <= mn[d] && grid_bounds[3 + d] >= mx[d], PCLB200_ERR_INVALID,
                   "voxelgrid: the given grid bounds do not contain the points");
      mn[d] = grid_bounds[d];
      mx[d] = grid_bounds[3 + d];
    }
  }
  // guard, in the reference's float arithmetic (voxel_grid.hpp:620-629)
  volatile float e0 = (mx[0] - mn[0]) * inv[0], e1 = (mx[1] - mn[1]) * inv[1], e2 = (mx[2] - mn[2]) * inv[2];
  const int64_t dx = (int64_t)e0 + 1, dy = (int64_t)e1 + 1, dz = (int64_t)e2 + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max())
    throw Error(PCLB200_ERR_LEAF_TOO_SMALL,
                "VoxelGrid: leaf size is too small for the input dataset, integer indices would overflow "
                "(voxel_grid.hpp:620-629)");
  VgParams gp;
  int max_b[3], div_b[3];
  for (int d = 0; d < 3; ++d) {
    volatile float lo_s = mn[d] * inv[d], hi_s = mx[d] * inv[d];
    gp.inv[d] = inv[d];
    gp.min_b[d] = (int)std::floor(lo_s);
    max_b[d] = (int)std::floor(hi_s);
    div_b[d] = max_b[d] - gp.min_b[d] + 1;
  }
  gp.mul[0] = 1;
  gp.mul[1] = div_b[0];
  gp.mul[2] = div_b[0] * div_b[1];
  // pass 2: keys ; pass 3: stable radix sort
  DevBuf<unsigned> keys_in, keys;
  DevBuf<int32_t> vals_in, vals;
  keys_in.alloc(cnt, st);
  keys.alloc(cnt, st);
  vals_in.alloc(cnt, st);
  vals.alloc(cnt, st);
  k_vg_keys<<<grid_for(cnt, 256), 256, 0, st>>>(dense.p, cnt, gp, keys_in.p, vals_in.p);
  ++c.launches;
  size_t tmp_bytes = 0;
  PCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.p, keys.p, vals_in.p, vals.p, (int)cnt, 0, 32, st));
  DevBuf<unsigned char> tmp;
  tmp.alloc(tmp_bytes, st);
  PCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.p, keys.p, vals_in.p, vals.p, (int)cnt, 0, 32, st));
  c.launches += 5;
  // pass 4: run starts
  DevBuf<unsigned char> head;
  head.alloc(n_valid, st);
  k_vg_heads<<<grid_for(n_valid, 256), 256, 0, st>>>(keys.p, n_valid, head.p);
  ++c.launches;
  DevBuf<unsigned> starts;
  DevBuf<size_t> d_num;
  starts.alloc(n_valid, st);
  d_num.alloc(1, st);
  cub::CountingInputIterator<unsigned> counting(0);
  size_t tb2 = 0;
  PCLB_CUDA(cub::DeviceSelect::Flagged(nullptr, tb2, counting, head.p, starts.p, d_num.p, (int)n_valid, st));
  DevBuf<unsigned char> tmp2;
  tmp2.alloc(tb2, st);
  PCLB_CUDA(cub::DeviceSelect::Flagged(tmp2.p, tb2, counting, head.p, starts.p, d_num.p, (int)n_valid, st));
  c.launches += 2;
  size_t n_runs = 0;
  PCLB_CUDA(cudaMemcpyAsync(&n_runs, d_num.p, sizeof(size_t), cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
  DevBuf<RunRef> runs_all, runs;
  runs_all.alloc(n_runs, st);
  k_vg_runs<<<grid_for(n_runs, 256), 256, 0, st>>>(starts.p, n_runs, n_valid, runs_all.p);
  ++c.launches;
  size_t n_out = n_runs;
  const RunRef* d_runs = runs_all.p;
  if (min_pts > 1) {
    DevBuf<unsigned char> keep;
    keep.alloc(n_runs, st);
    k_vg_keep<<<grid_for(n_runs, 256), 256, 0, st>>>(starts.p, n_runs, n_valid, min_pts, keep.p);
    runs.alloc(n_runs, st);
    size_t tb3 = 0;
    PCLB_CUDA(cub::DeviceSelect::Flagged(nullptr, tb3, runs_all.p, keep.p, runs.p, d_num.p, (int)n_runs, st));
    DevBuf<unsigned char> tmp3;
    tmp3.alloc(tb3, st);
    PCLB_CUDA(cub::DeviceSelect::Flagged(tmp3.p, tb3, runs_all.p, keep.p, runs.p, d_num.p, (int)n_runs, st));
    c.launches += 3;
    PCLB_CUDA(cudaMemcpyAsync(&n_out, d_num.p, sizeof(size_t), cudaMemcpyDeviceToHost, st));
    PCLB_CUDA(cudaStreamSynchronize(st));
    d_runs = runs.p;
  }
  if (n_out == 0)
    return 0;
  // pass 5: centroids
  DevBuf<float4> out;
  float4* d_out = nullptr;
  const bool out_on_device = is_device_ptr(out_xyz1);
  if (out_on_device)
    d_out = reinterpret_cast<float4*>(out_xyz1);
  else {
    out.alloc(n_out, st);
    d_out = out.p;
  }
  k_vg_centroids<<<grid_for(n_out, 128), 128, 0, st>>>(dense.p, vals.p, d_runs, n_out, d_out);
  ++c.launches;
  PCLB_CUDA(cudaGetLastError());
  if (!out_on_device)
    PCLB_CUDA(cudaMemcpyAsync(out_xyz1, out.p, n_out * sizeof(float4), cudaMemcpyDeviceToHost, st));
  if (normals && out_normal_curv) {
    // the remaining fields of a PointNormal / Normal record, averaged over the same runs in the same order
    PCLB_REQUIRE(stride_n >= 20 && stride_n % 4 == 0, PCLB200_ERR_INVALID, "normal stride must be a multiple of 4 and >= 20");
    DevBuf<unsigned char> staged;
    DevBuf<int32_t> staged_sub;
    const unsigned char* d_src = static_cast<const unsigned char*>(normals);
    if (!is_device_ptr(normals)) {
      const size_t bytes = (n - 1) * stride_n + 20;
      staged.alloc(bytes, st);
      PCLB_CUDA(cudaMemcpyAsync(staged.p, normals, bytes, cudaMemcpyHostToDevice, st));
      d_src = staged.p;
    }
    const int32_t* d_sub = indices;
    if (indices && !is_device_ptr(indices)) {
      staged_sub.alloc(n_idx, st);
      PCLB_CUDA(cudaMemcpyAsync(staged_sub.p, indices, n_idx * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      d_sub = staged_sub.p;
    }
    DevBuf<float4> nc, nout;
    nc.alloc(2 * cnt, st);
    k_vg_load_nc<<<grid_for(cnt, 256), 256, 0, st>>>(d_src, stride_n, d_sub, cnt, nc.p);
    const bool nc_on_device = is_device_ptr(out_normal_curv);
    float4* d_nout = reinterpret_cast<float4*>(out_normal_curv);
    if (!nc_on_device) {
      nout.alloc(2 * n_out, st);
      d_nout = nout.p;
    }
    k_vg_normals<<<grid_for(n_out, 128), 128, 0, st>>>(nc.p, vals.p, d_runs, n_out, d_nout);
    c.launches += 2;
    PCLB_CUDA(cudaGetLastError());
    if (!nc_on_device)
      PCLB_CUDA(cudaMemcpyAsync(out_normal_curv, nout.p, 2 * n_out * sizeof(float4), cudaMemcpyDeviceToHost, st));
    PCLB_CUDA(cudaStreamSynchronize(st));
    return n_out;
  }
  PCLB_CUDA(cudaStreamSynchronize(st));
  return n_out;
}

}  // namespace pclb200
