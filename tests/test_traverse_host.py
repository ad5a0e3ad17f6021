"""The device traversal header (pcl_b200/csrc/traverse.cuh: walk, nearest1, the cell-table look-ups and their conservative
bounds) compiled for the HOST and run against brute force — tests/host/traverse_host_test.cpp.  CPU only: the CUDA
intrinsics are supplied with the same rounding, the index is built on the host to lbvh.cu's invariants.  ~2e5 checks on
eleven scene families (ties, duplicates beyond a leaf, points on cell boundaries, degenerate frames, gates, far queries, 300 random
small clouds), every query with and without the
cell table, with no / the true / a random seed, and with the TRACK visitor; plus the temporal-coherence chain of k_search
(still_nearest + the TRACK bound over drifting queries): a skipped walk never keeps a match that stopped being the nearest."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")) or shutil.which("g++") is None,
                    reason="needs g++ and the CUDA headers (vector types only; nothing is run on a device)")
def test_device_traversal_header_on_the_host(tmp_path):
    exe = str(tmp_path / "traverse_host_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-frounding-math", "-ffp-contract=off", "-fno-fast-math",
                           "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "host", "traverse_host_test.cpp"), "-o", exe])
    r = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:]


@pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")) or shutil.which("g++") is None,
                    reason="needs g++ and the CUDA headers (vector types only; nothing is run on a device)")
def test_warp_knn_kernel_on_an_emulated_warp(tmp_path):
    """pcl_b200/csrc/knn_warp.cuh (k_knn_warp<false|true>, the shuffled bitonic sort / merge, the ranked insertion, the
    certification rule, the normals epilogue) compiled for the host and run on a lock-step emulation of one warp
    (tests/host/warp_emu.h): every certified row equals brute force bit for bit, every normal equals the host
    computePointNormal on the brute-force list, and the constructed shared-leaf scene — the bug the 10 M-row comparison
    found on the device — passes (it fails on the kernel as it was before the fix)."""
    exe = str(tmp_path / "knn_warp_host_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-frounding-math", "-ffp-contract=off", "-fno-fast-math",
                           "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "host"),
                           "-I" + os.path.join(ROOT, "pcl_b200", "pcl_compat"),
                           os.path.join(ROOT, "tests", "host", "knn_warp_host_test.cpp"), "-o", exe])
    r = subprocess.run([exe, "1"], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:]
    assert "shared leaf, constructed #5" in r.stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")) or shutil.which("g++") is None,
                    reason="needs g++ and the CUDA headers (vector types only; nothing is run on a device)")
def test_whole_icp_iterations_on_an_emulated_block(tmp_path):
    """pcl_b200/csrc/icp_kernels.cuh — k_search, k_accum_dmma (the m8n8k4 fp64 MMA emulated fragment by fragment), k_solve
    with the convergence criteria in its tail — compiled for the host, run on a lock-step emulation of a 256-thread block
    and driven like icp.cu's enqueue-ahead path: every iteration's correspondences equal brute force, the accumulated
    normal equations equal plain fp64 sums, and iterations / state / counts / final transform of twelve aligns (SVD,
    point-to-plane, the symmetric objective, ICPWithNormals, reciprocal correspondences; float and double; gates; tracking
    off / on / automatic) equal the oracle's loop (1e-5 float, 1e-9 double); the four stand-alone estimators against the
    oracle's."""
    import oracle
    oracle.build()
    odir = os.path.join(ROOT, "oracle")
    exe = str(tmp_path / "icp_host_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-frounding-math", "-ffp-contract=off", "-fno-fast-math",
                           "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "host"),
                           os.path.join(ROOT, "tests", "host", "icp_host_test.cpp"), "-o", exe,
                           "-L" + odir, "-lpcl_oracle", "-Wl,-rpath," + odir])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:]
    assert "tracking automatic" in r.stdout and "point-to-plane LLS double" in r.stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")) or shutil.which("g++") is None,
                    reason="needs g++ and the CUDA headers (vector types only; nothing is run on a device)")
def test_per_thread_search_kernels_on_the_host(tmp_path):
    """pcl_b200/csrc/search_kernels.cuh — k_knn<K> for every compiled list size, k_knn_any, k_knn_stats, k_radius_count /
    k_radius_fill, k_normals<K> — compiled for the host and run block by block against brute force (lists bit for bit,
    ties and duplicates included) and, for the normals, against the host computePointNormal."""
    exe = str(tmp_path / "search_host_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-frounding-math", "-ffp-contract=off", "-fno-fast-math",
                           "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "host"),
                           "-I" + os.path.join(ROOT, "pcl_b200", "pcl_compat"),
                           os.path.join(ROOT, "tests", "host", "search_host_test.cpp"), "-o", exe])
    r = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:]
