// consumers_host_test.cpp — the kernels behind the stand-alone consumers of the searcher (pcl_b200/csrc/icp_kernels.cuh:
// k_corr plain and reciprocal = pclb200_correspondences, k_fitness = pclb200_fitness_score, k_gicp_cov =
// pclb200_gicp_covariances) compiled for the HOST, run on the emulated thread block, against the CPU oracle
// (oracle/libpcl_oracle.so, linked: test infrastructure): correspondence lists bit for bit (gate, index subsets, non-finite
// source points, duplicates), the fitness score to 1e-12, the regularised covariances to 1e-9.
#define PCLB_HOST_EXTRA_SHIMS "warp_emu.h"
#define PCLB_HOST_EMULATION 1
#include "host_index.h"

#include <cfloat>

#include "../../pcl_b200/csrc/icp_kernels.cuh"

extern "C" {
void* orc_index_build(const float* pts, size_t n, size_t stride, const int32_t* subset, size_t n_subset);
void orc_index_free(void* h);
int orc_knn(void* h, const float* q, size_t nq, size_t qstride, int k, int32_t* out_idx, float* out_d2, int nthreads);
size_t orc_correspondences(void* h_tgt, const float* src, size_t n_src, size_t sstride, const int32_t* indices, size_t n_idx, int is_dense,
                           double max_distance, pclb200_corr* out, int nthreads);
size_t orc_correspondences_reciprocal(void* h_tgt, void* h_src, const float* src, size_t n_src, size_t sstride, const float* tgt, size_t tstride,
                                      const int32_t* indices, size_t n_idx, int is_dense, double max_distance, pclb200_corr* out, int nthreads);
double orc_fitness_score(void* h_tgt, const float* src, size_t n_s, size_t sstride, const int32_t* indices, size_t n_idx, int is_dense,
                         const double* final_T, int scalar_is_double, double max_range, int nthreads);
void orc_gicp_covariances(void* h, const float* cloud, size_t n, size_t stride, int k, double gicp_epsilon, double* out, int nthreads);
}

static long g_checks = 0, g_fail = 0;
#define CHECK(c, ...) do { ++g_checks; if (!(c)) { if (++g_fail <= 20) { std::printf("FAIL %s:%d %s  ", __FILE__, __LINE__, #c); std::printf(__VA_ARGS__); std::printf("\n"); } } } while (0)

static float gate_from_max_dist(double max_dist)   // icp.cu: the largest float not above max_dist^2
{
  const double m2 = max_dist * max_dist;
  if (!(m2 < (double)FLT_MAX)) return FLT_MAX;
  float g = (float)m2;
  if ((double)g > m2) g = std::nextafter(g, -INFINITY);
  return g;
}

template <typename F> static void launch(unsigned grid, int block, F kernel)
{
  blockDim.x = block;
  gridDim.x = grid ? grid : 1;
  for (unsigned b = 0; b < gridDim.x; ++b) { blockIdx_storage.x = b; warp_emu::run_block(block, kernel); }
  blockIdx_storage.x = 0;
  gridDim.x = 1;
}

static std::vector<float> xyz_of(const std::vector<float>& c4)
{
  std::vector<float> v(c4.size() / 4 * 3);
  for (std::size_t i = 0; i < c4.size() / 4; ++i) for (int d = 0; d < 3; ++d) v[3 * i + d] = c4[4 * i + d];
  return v;
}

// icp.cu: correspondences() — queries in slot order (the Morton sort of the real driver only permutes the launch order)
static std::vector<pclb200_corr> device_correspondences(const HostIndex& IT, const HostIndex* IS, const std::vector<float>& src, const std::vector<int32_t>* indices,
                                                        double max_dist)
{
  const std::size_t nq = indices ? indices->size() : src.size() / 4;
  std::vector<float4> q(nq);
  for (std::size_t i = 0; i < nq; ++i) {
    const float* p = &src[4 * (indices ? (std::size_t)(*indices)[i] : i)];
    q[i] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));
  }
  std::vector<pclb200_corr> by_slot(nq), out;
  int d_error = 0;
  const TreeView T = IT.view(true);
  const float gate = gate_from_max_dist(max_dist);
  const unsigned grid = (unsigned)((nq + 127) / 128);
  if (IS)
    launch(grid, 128, [&] { k_corr<true>(T, q.data(), nq, gate, IS->nodes.data(), IS->pts.data(), IS->root, indices ? indices->data() : nullptr, by_slot.data(), &d_error); });
  else
    launch(grid, 128, [&] { k_corr<false>(T, q.data(), nq, gate, nullptr, nullptr, 0, indices ? indices->data() : nullptr, by_slot.data(), &d_error); });
  CHECK(d_error == 0, "traversal stack overflow flag");
  for (const pclb200_corr& c : by_slot) if (c.index_match >= 0) out.push_back(c);   // cub::DeviceSelect::If(CorrValid)
  return out;
}

static void compare_lists(const char* what, const std::vector<pclb200_corr>& got, const std::vector<pclb200_corr>& want)
{
  CHECK(got.size() == want.size(), "%s: %zu pairs, the oracle has %zu", what, got.size(), want.size());
  std::size_t bad = 0;
  for (std::size_t i = 0; i < got.size() && i < want.size(); ++i) bad += std::memcmp(&got[i], &want[i], sizeof(pclb200_corr)) != 0;
  CHECK(bad == 0, "%s: %zu records differ", what, bad);
  std::printf("%-58s %6zu pairs (oracle %zu)\n", what, got.size(), want.size());
}

int main()
{
  std::mt19937 rng(99);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  std::normal_distribution<float> N(0.f, 1.f);
  const int nt = 6000, ns = 4000;
  std::vector<float> tgt(4 * nt, 1.f), src(4 * ns, 1.f);
  for (int i = 0; i < nt; ++i) { tgt[4 * i] = 2.f * U(rng); tgt[4 * i + 1] = 2.f * U(rng); tgt[4 * i + 2] = 0.2f * std::sin(4.f * tgt[4 * i]) + 0.01f * N(rng); }
  for (int i = 0; i < 40; ++i) std::memcpy(&tgt[4 * (nt - 1 - i)], &tgt[4 * i], 12);   // duplicated target points: the smaller index wins
  const float ca = std::cos(0.03f), sa = std::sin(0.03f);
  for (int i = 0; i < ns; ++i) {
    const float* t = &tgt[4 * (int)(U(rng) * nt)];
    src[4 * i] = ca * t[0] - sa * t[1] + 0.01f + 0.002f * N(rng);
    src[4 * i + 1] = sa * t[0] + ca * t[1] - 0.008f + 0.002f * N(rng);
    src[4 * i + 2] = t[2] + 0.004f + 0.002f * N(rng);
  }
  HostIndex IT, IS;
  build_index(IT, xyz_of(tgt), 4);
  build_index(IS, xyz_of(src), 4);
  void* ot = orc_index_build(tgt.data(), nt, 4, nullptr, 0);
  void* os = orc_index_build(src.data(), ns, 4, nullptr, 0);
  std::vector<int32_t> subset;
  for (int i = ns - 1; i >= 0; i -= 3) subset.push_back(i);   // a descending subset: output order is the subset's

  std::vector<pclb200_corr> want(ns);
  for (double md : {std::sqrt(std::numeric_limits<double>::max()), 0.05, 0.01, 0.0}) {
    char name[128];
    std::snprintf(name, sizeof name, "nearest, max distance %.3g", md > 1e10 ? INFINITY : md);
    want.resize(ns);
    want.resize(orc_correspondences(ot, src.data(), ns, 4, nullptr, 0, 1, md, want.data(), 2));
    compare_lists(name, device_correspondences(IT, nullptr, src, nullptr, md), want);
    std::snprintf(name, sizeof name, "reciprocal, max distance %.3g", md > 1e10 ? INFINITY : md);
    want.resize(ns);
    want.resize(orc_correspondences_reciprocal(ot, os, src.data(), ns, 4, tgt.data(), 4, nullptr, 0, 1, md, want.data(), 2));
    compare_lists(name, device_correspondences(IT, &IS, src, nullptr, md), want);
  }
  want.resize(ns);
  want.resize(orc_correspondences(ot, src.data(), ns, 4, subset.data(), subset.size(), 1, 0.03, want.data(), 2));
  compare_lists("nearest, index subset (descending)", device_correspondences(IT, nullptr, src, &subset, 0.03), want);
  want.resize(ns);
  want.resize(orc_correspondences_reciprocal(ot, os, src.data(), ns, 4, tgt.data(), 4, subset.data(), subset.size(), 1, 0.03, want.data(), 2));
  compare_lists("reciprocal, index subset (descending)", device_correspondences(IT, &IS, src, &subset, 0.03), want);
  {
    std::vector<float> holes = src;
    for (int i = 5; i < ns; i += 41) holes[4 * i + (i % 3)] = i % 2 ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity();
    want.resize(ns);
    want.resize(orc_correspondences(ot, holes.data(), ns, 4, nullptr, 0, 0, 0.05, want.data(), 2));
    compare_lists("nearest, non-finite source points skipped", device_correspondences(IT, nullptr, holes, nullptr, 0.05), want);
  }

  // ---- Registration::getFitnessScore: transform (float or double), 1-NN, mean of the squared distances <= max_range ----
  {
    const double a = 0.02, T[16] = {std::cos(a), -std::sin(a), 0, 0.004, std::sin(a), std::cos(a), 0, -0.003, 0, 0, 1, 0.001, 0, 0, 0, 1};
    const TreeView TV = IT.view(true);
    for (int dbl = 0; dbl < 2; ++dbl)
      for (double max_range : {std::numeric_limits<double>::max(), 1e-4, 1e-7, 0.0}) {
        Pending h;
        std::memset(&h, 0, sizeof h);
        for (int i = 0; i < 12; ++i) { h.f[i] = (float)T[i]; h.d[i] = dbl ? T[i] : (double)(float)T[i]; }
        h.apply = 1;
        h.mode = dbl ? 2 : 1;
        std::vector<float4> q(ns);
        for (int i = 0; i < ns; ++i) { q[i] = make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], __int_as_float(i)); apply_pending(h, q[i].x, q[i].y, q[i].z); }
        const unsigned grid = 3;
        std::vector<double> partials(grid * kAccum, 0.0), accum(kAccum, -1.0);
        unsigned counter = 0;
        int d_error = 0;
        IterArgs pub;
        std::memset(&pub, 0, sizeof pub);
        pub.partials = partials.data(); pub.counter = &counter; pub.accum = accum.data(); pub.d_error = &d_error;
        launch(grid, 256, [&] { k_fitness(TV, q.data(), (size_t)ns, max_range, pub); });
        const double got = accum[0] > 0 ? accum[1] / accum[0] : std::numeric_limits<double>::max();
        const double wantf = orc_fitness_score(ot, src.data(), ns, 4, nullptr, ns, 1, T, dbl, max_range, 2);
        CHECK(d_error == 0 && std::fabs(got - wantf) <= 1e-12 * std::fabs(wantf), "fitness %s max_range %g: %.17g, the oracle's %.17g", dbl ? "double" : "float", max_range,
              got, wantf);
        std::printf("fitness score, %-6s max range %-12.3g %.12g over %.0f points (oracle %.12g)\n", dbl ? "double" : "float", max_range, got, accum[0], wantf);
      }
  }

  // ---- GICP covariances over exact k-NN rows -------------------------------------------------------------------------
  {
    const int k = 20;
    std::vector<int32_t> rows((size_t)nt * k);
    std::vector<float> rd((size_t)nt * k);
    orc_knn(ot, tgt.data(), nt, 4, k, rows.data(), rd.data(), 2);
    std::vector<int32_t> pos_of_orig(nt, -1);
    for (std::size_t p = 0; p < IT.pts.size(); ++p) { const int o = __float_as_int(IT.pts[p].w); if (o != kSentinelIndex) pos_of_orig[o] = (int32_t)p; }
    std::vector<float4> q(nt);
    for (int i = 0; i < nt; ++i) q[i] = make_float4(tgt[4 * i], tgt[4 * i + 1], tgt[4 * i + 2], 1.f);
    std::vector<double> got((size_t)nt * 9, -1.0), wantc((size_t)nt * 9, 0.0);
    launch((unsigned)((nt + 127) / 128), 128, [&] { k_gicp_cov(q.data(), (size_t)nt, rows.data(), k, k, IT.pts.data(), pos_of_orig.data(), 0.001, got.data()); });
    orc_gicp_covariances(ot, tgt.data(), nt, 4, k, 0.001, wantc.data(), 2);
    double worst = 0;
    for (std::size_t e = 0; e < got.size(); ++e) worst = std::max(worst, std::fabs(got[e] - wantc[e]));
    CHECK(worst < 1e-9, "GICP covariances: largest difference %.3g", worst);
    std::printf("GICP covariances k = 20: largest |difference| over %d matrices %.3g\n", nt, worst);
  }
  orc_index_free(ot);
  orc_index_free(os);
  std::printf("%ld checks, %ld failures\n%s\n", g_checks, g_fail, g_fail ? "FAILED" : "PASSED");
  return g_fail ? 1 : 0;
}
