// device_build.h — the index build of lbvh.cu: build_from_dense, replayed on the host with the REAL kernels
// (pcl_b200/csrc/lbvh_kernels.cuh, compiled for the host and run block by block on the emulator of warp_emu.h) and std::
// algorithms where the driver calls CUB (radix sort of the keys, the two prefix sums).  The sequence, the sizes and the
// host-side decisions (one-leaf clouds, the cell table's finest level and slot count, the margin) follow lbvh.cu line by
// line.  Test infrastructure: lets every host-compiled search test run on an index the device kernels built.
#pragma once
#include <numeric>

#include "../../pcl_b200/csrc/lbvh_kernels.cuh"

namespace device_build {
using namespace pclb200;

template <typename F> inline void launch(std::size_t n_threads_needed, int block, F kernel)
{
  const unsigned grid = static_cast<unsigned>((n_threads_needed + block - 1) / block);
  blockDim.x = block;
  gridDim.x = grid ? grid : 1;
  for (unsigned b = 0; b < (grid ? grid : 1); ++b) {
    blockIdx_storage.x = b;
    warp_emu::run_block(block, kernel);
  }
  blockIdx_storage.x = 0;
  gridDim.x = 1;
}

inline void build_impl(HostIndex& I, const std::vector<float>& xyz, bool build_cell_table);
inline void build(HostIndex& I, const std::vector<float>& xyz, bool build_cell_table)
{
  const EmuDim saved_block = blockDim, saved_grid = gridDim;   // the caller's launch geometry survives the build
  build_impl(I, xyz, build_cell_table);
  blockDim = saved_block;
  gridDim = saved_grid;
  blockIdx_storage.x = 0;
}
inline void build_impl(HostIndex& I, const std::vector<float>& xyz, bool build_cell_table)
{
  I = HostIndex();
  I.xyz = xyz;
  const std::size_t n = xyz.size() / 3;
  std::vector<float4> d_pts(n);
  for (std::size_t i = 0; i < n; ++i) d_pts[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
  // ---- morton_sort ----------------------------------------------------------------------------------------------------
  BBoxAcc acc;
  for (int d = 0; d < 3; ++d) { acc.lo[d] = 0x7fffffff; acc.hi[d] = (int)0x80000000; }
  acc.count = 0;
  {
    blockDim.x = 256;
    const unsigned g = static_cast<unsigned>(std::min<std::size_t>((n + 255) / 256, 8));
    gridDim.x = g ? g : 1;
    for (unsigned b = 0; b < gridDim.x; ++b) { blockIdx_storage.x = b; warp_emu::run_block(256, [&] { k_bbox(d_pts.data(), n, &acc); }); }
    blockIdx_storage.x = 0; gridDim.x = 1;
  }
  const std::size_t n_valid = static_cast<std::size_t>(acc.count);
  float ext = 0.f;
  for (int d = 0; d < 3; ++d) { I.lo[d] = ord2f(acc.lo[d]); I.hi[d] = ord2f(acc.hi[d]); ext = std::max(ext, I.hi[d] - I.lo[d]); }
  I.scale = ext > 0.f ? 2097152.f / ext : 1.f;
  if (!std::isfinite(I.scale)) I.scale = 1.f;
  std::vector<unsigned long long> keys_in(n), keys(n);
  std::vector<int32_t> vals_in(n), vals(n);
  launch(n, 256, [&] { k_morton<false>(d_pts.data(), n, I.lo[0], I.lo[1], I.lo[2], I.scale, keys_in.data(), vals_in.data()); });
  {  // cub::DeviceRadixSort::SortPairs (stable)
    std::vector<std::size_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return keys_in[a] < keys_in[b]; });
    for (std::size_t i = 0; i < n; ++i) { keys[i] = keys_in[order[i]]; vals[i] = vals_in[order[i]]; }
  }
  // ---- build_from_dense ----------------------------------------------------------------------------------------------
  const int nv = static_cast<int>(n_valid);
  if (nv <= kLeafSize) {
    I.root = ~0;
    I.pts.assign(kLeafSize, make_float4(0, 0, 0, 0));
    blockDim.x = 256; gridDim.x = 1;
    warp_emu::run_block(256, [&] { k_gather_sorted(d_pts.data(), vals.data(), nullptr, n_valid, kLeafSize, I.pts.data()); });
    return;
  }
  const int ni = nv - 1;
  std::vector<int2> kchildren(ni), krange(ni);
  std::vector<int> knode_parent(ni), kpoint_parent(nv), keep(ni), new_id(ni), leaf_flag(nv, 0), leaf_incl(nv);
  launch(ni, 256, [&] { k_karras_points(keys.data(), nv, kchildren.data(), knode_parent.data(), kpoint_parent.data(), krange.data()); });
  launch(nv, 256, [&] { k_mark_cells(nv, krange.data(), knode_parent.data(), kpoint_parent.data(), keep.data(), leaf_flag.data()); });
  std::partial_sum(leaf_flag.begin(), leaf_flag.end(), leaf_incl.begin());                   // cub::DeviceScan::InclusiveSum
  { int run = 0; for (int i = 0; i < ni; ++i) { new_id[i] = run; run += keep[i]; } }          // cub::DeviceScan::ExclusiveSum
  unsigned h_hist[64] = {0};
  const bool want_cells = build_cell_table && nv >= 256;
  if (want_cells) {
    blockDim.x = 256;
    const unsigned g = std::min<unsigned>(static_cast<unsigned>((nv + 255) / 256), 8u);
    gridDim.x = g;
    for (unsigned b = 0; b < g; ++b) { blockIdx_storage.x = b; warp_emu::run_block(256, [&] { k_prefix_hist(keys.data(), nv, h_hist); }); }
    blockIdx_storage.x = 0; gridDim.x = 1;
  }
  const int n_leaves = leaf_incl[nv - 1];
  const int n_int = new_id[ni - 1] + keep[ni - 1];
  if (!(n_leaves >= 2 && n_int == n_leaves - 1)) { std::printf("device_build: inconsistent cell cut (%d leaves, %d nodes)\n", n_leaves, n_int); std::exit(4); }
  I.root = 0;
  const std::size_t n_padded = static_cast<std::size_t>(n_leaves) * kLeafSize;
  I.pts.assign(n_padded, make_float4(0, 0, 0, 0));
  I.nodes.assign(n_int, BvhNode());
  I.node_leaves.assign(n_int, make_int2(0, 0));
  std::vector<int> leaf_start(n_leaves), node_parent(n_int), leaf_parent(n_leaves);
  std::vector<int2> children(n_int);
  std::vector<float4> leaf_lo(n_leaves), leaf_hi(n_leaves), node_lo(n_int), node_hi(n_int);
  std::vector<unsigned> flags(n_int, 0u);
  launch(nv, 256, [&] { k_leaf_starts(nv, leaf_flag.data(), leaf_incl.data(), leaf_start.data()); });
  launch(n_padded, 256, [&] { k_fill_sentinels(I.pts.data(), n_padded); });
  launch(nv, 256, [&] { k_scatter_cells(d_pts.data(), vals.data(), nullptr, nv, leaf_incl.data(), leaf_start.data(), I.pts.data()); });
  CellTableW cellw{nullptr, 0, 0, 0};
  std::vector<unsigned long long> slots64;
  if (want_cells) {
    std::uint64_t cum = 1, entries = 0;
    int l = 0, bmax = 0;
    for (int b = 1; b <= kCellMaxBits; ++b) {
      for (; l < 3 * b; ++l) cum += h_hist[l];
      if (cum > static_cast<std::uint64_t>(nv) / 4) break;
      bmax = b;
      entries += cum;
    }
    if (bmax >= 1) {
      unsigned lg = 6;
      while ((std::uint64_t(1) << lg) < 2 * entries) ++lg;
      slots64.assign(std::size_t(1) << lg, 0ull);
      I.log2_slots = lg;
      I.bmax = bmax;
      float m = 0.f;
      for (int d = 0; d < 3; ++d) m = std::max(m, std::max(std::fabs(I.lo[d]), std::fabs(I.hi[d])));
      m = std::max(m, 2097152.f / I.scale);
      I.margin = 2e-6f * m;
      cellw.slots = slots64.data();
      cellw.shift = 32 - lg;
      cellw.mask = (1u << lg) - 1u;
      cellw.bmax = bmax;
    }
  }
  launch(ni, 256, [&] {
    k_link_cells(nv, kchildren.data(), krange.data(), keep.data(), new_id.data(), leaf_incl.data(), keys.data(), children.data(), node_parent.data(),
                 leaf_parent.data(), I.node_leaves.data(), cellw);
  });
  launch(n_leaves, 256, [&] {
    k_refit(I.pts.data(), n_leaves, children.data(), node_parent.data(), leaf_parent.data(), leaf_lo.data(), leaf_hi.data(), node_lo.data(), node_hi.data(),
            flags.data());
  });
  launch(n_int, 256, [&] { k_pack_nodes(n_int, children.data(), leaf_lo.data(), leaf_hi.data(), node_lo.data(), node_hi.data(), I.nodes.data()); });
  I.slots.resize(slots64.size());
  for (std::size_t i = 0; i < slots64.size(); ++i) I.slots[i] = make_uint2(static_cast<unsigned>(slots64[i] & 0xffffffffull), static_cast<unsigned>(slots64[i] >> 32));
}
}  // namespace device_build
