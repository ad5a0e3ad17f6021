"""The header-only facade (pcl_b200/pcl_compat/pcl/**) on a machine without a GPU — CPU only.

The facade's device calls go through the C-ABI of include/pclb200.h.  Here the facade's own test programs (the same sources
tests/test_facade_gpu.py and tests/test_zz_facade_extra_gpu.py run on a B200 against libpclb200.so) are linked against a TEST
DOUBLE of that C-ABI, tests/host/pclb200_on_oracle.cpp, which answers every call with the CPU oracle.  What this proves is
the HOST side of the boundary: argument marshalling, the state the pcl:: classes keep around each call, the reference's
behaviour in the error paths, and that the expectations written into the programs are the oracle's (hence the reference's)
answers.  It proves nothing about the kernels (tests/test_traverse_host.py and the `-m gpu` tests do that), and the test
double is not a fallback: it is compiled into pytest's temporary directory and nothing in pcl_b200/ can load it."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from test_facade_gpu import FACADE, ROOT, _write_ascii_pcd, _write_binary_pcd, _write_golden

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def double(tmp_path_factory):
    import oracle
    oracle.build()
    d = tmp_path_factory.mktemp("facade_on_oracle")
    odir = os.path.join(ROOT, "oracle")
    lib = str(d / "libpclb200_on_oracle.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-Wno-unused-parameter",
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "pclb200_on_oracle.cpp"),
                           "-o", lib, "-L" + odir, "-lpcl_oracle", "-Wl,-rpath," + odir])

    def program(source):
        exe = str(d / os.path.splitext(os.path.basename(source))[0])
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + FACADE, "-I" + os.path.join(ROOT, "include"),
                               os.path.join(FACADE, source), "-o", exe, lib, "-Wl,-rpath," + str(d), "-pthread"])
        return exe
    return d, lib, program


def test_the_double_exports_the_whole_c_abi(double):
    """Every entry point include/pclb200.h declares — so a facade call that is not answered fails at the call, with a
    message, instead of at link time — and none of the oracle's symbols leak into the product library."""
    import re
    _, lib, _ = double
    declared = set(re.findall(r"\b(pclb200_[a-z0-9_]+)\s*\(", open(os.path.join(ROOT, "include", "pclb200.h")).read()))
    exported = {ln.split()[-1] for ln in subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True).splitlines()}
    assert declared <= exported, sorted(declared - exported)
    product = os.path.join(ROOT, "pcl_b200", "libpclb200.so")
    if os.path.exists(product):
        needed = subprocess.check_output(["readelf", "-d", product], text=True)
        assert "pcl_oracle" not in needed and "on_oracle" not in needed


def _clouds(golden, d):
    _write_ascii_pcd(d / "bun0.pcd", golden["bun0"])
    _write_binary_pcd(d / "bun4.pcd", golden["bun4"])
    return str(d / "bun0.pcd"), str(d / "bun4.pcd")


def test_reference_test_bodies_through_the_facade(double, golden, tmp_path):
    """pcl_compat/tests/test_facade.cpp: the reference's registration / kdtree / filters / features / segmentation test
    bodies written against the drop-in pcl:: classes, 8 600 checks incl. the reference's golden vectors."""
    _, _, program = double
    b0, b4 = _clouds(golden, tmp_path)
    _write_golden(tmp_path / "golden.txt", golden)
    r = subprocess.run([program("tests/test_facade.cpp"), b0, b4, str(tmp_path / "golden.txt")], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-1500:]


def test_extra_facade_program(double, golden, tmp_path):
    """pcl_compat/tests/test_facade_extra.cpp: Search<PointT> overloads, CorrespondenceEstimation under a point
    representation, DefaultConvergenceCriteria thresholds, PinnedCloud, VoxelGrid leaf layout, and the host-side
    CorrespondenceRejectorSampleConsensus: the reference's golden 97 inlier pairs in order and its transform to 1e-4
    (test_registration_api.cpp:225-263), the reference's ICP with median + sample-consensus rejectors under ten random poses
    (test_registration.cpp:336-382)."""
    _, _, program = double
    b0, b4 = _clouds(golden, tmp_path)
    _write_golden(tmp_path / "golden.txt", golden)
    r = subprocess.run([program("tests/test_facade_extra.cpp"), b0, b4, str(tmp_path / "golden.txt")], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-1500:]
    assert int(r.stdout.split(" checks")[0].split()[-1]) > 400   # the golden block ran


def test_icp_command_line_program(double, golden, tmp_path):
    """pcl_compat/examples/iterative_closest_point.cpp (the flow of the reference's tools/iterative_closest_point.cpp):
    blob PCD in, ICP<PointNormal, double> with injected estimators and a one-to-one rejector, concatenateFields, PCD out."""
    _, _, program = double
    b0, b4 = _clouds(golden, tmp_path)
    out = tmp_path / "aligned.pcd"
    r = subprocess.run([program("examples/iterative_closest_point.cpp"), b0, b4, str(out), "50", "0.05"], capture_output=True, text=True)
    assert r.returncode in (0, 1) and "has converged" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
    text = out.read_text().splitlines()
    hdr = {ln.split()[0]: ln.split()[1:] for ln in text[:11] if ln and not ln.startswith("#")}
    assert hdr["FIELDS"][:3] == ["x", "y", "z"] and "normal_x" in hdr["FIELDS"] and hdr["POINTS"] == ["397"]
    body = np.array([[float(t) for t in ln.split()] for ln in text[11:] if ln.strip()])
    assert body.shape[0] == 397 and np.isfinite(body[:, :3]).all()
    tgt = np.asarray(golden["bun4"], dtype=np.float64)[:, :3]

    def mean_nn(a):
        return np.sqrt(((a[:, None, :] - tgt[None, :, :]) ** 2).sum(-1).min(1)).mean()
    assert mean_nn(body[:, :3]) < 0.5 * mean_nn(np.asarray(golden["bun0"], dtype=np.float64)[:, :3])


TUTORIALS = "/root/reference/doc/tutorials/content/sources"


@pytest.mark.skipif(not os.path.isdir(TUTORIALS), reason="the reference checkout (its tutorial sources) is not on this machine")
def test_reference_tutorials_compile_unchanged(double, tmp_path):
    """The reference's own tutorial programs for the path — iterative_closest_point, kdtree_search, voxel_grid (the
    PCLPointCloud2 form), pcd_read, pcd_write, concatenate_clouds / _fields / _points, statistical_removal,
    radius_outlier_removal — and seven of its examples/ (normal estimation, extract indices, remove NaN, copy point cloud,
    min / max coordinates, point validity, organised clouds) compiled UNCHANGED, from where they lie in the reference checkout, against the facade headers;
    the ones that need no data file are run (on the test double): the ICP tutorial converges and prints the 0.7 shift."""
    _, lib, _ = double
    built = {}
    for name in ("iterative_closest_point", "kdtree_search", "voxel_grid", "pcd_read", "pcd_write", "concatenate_clouds",
                 "concatenate_fields", "concatenate_points", "statistical_removal", "radius_outlier_removal"):
        src = os.path.join(TUTORIALS, name, name + ".cpp")
        exe = str(tmp_path / name)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + FACADE, "-I" + os.path.join(ROOT, "include"), src, "-o", exe, lib,
                               "-Wl,-rpath," + os.path.dirname(lib), "-pthread"])
        built[name] = exe
    examples = "/root/reference/examples"
    for rel in ("features/example_normal_estimation", "filters/example_extract_indices", "filters/example_remove_nan_from_point_cloud",
                "common/example_copy_point_cloud", "common/example_get_max_min_coordinates", "common/example_check_if_point_is_valid",
                "common/example_organized_point_cloud"):
        name = os.path.basename(rel)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + FACADE, "-I" + os.path.join(ROOT, "include"), os.path.join(examples, rel + ".cpp"),
                               "-o", str(tmp_path / name), lib, "-Wl,-rpath," + os.path.dirname(lib), "-pthread"])
        built[name] = str(tmp_path / name)
    r = subprocess.run([built["example_check_if_point_is_valid"]], capture_output=True, text=True)
    assert r.returncode == 0 and "Is p_valid valid? 1" in r.stdout and "Is p_invalid valid? 0" in r.stdout
    r = subprocess.run([built["example_extract_indices"]], capture_output=True, text=True)
    assert r.returncode == 0 and "Cloud has 5 points." in r.stdout and "Output has 2 points." in r.stdout
    r = subprocess.run([built["iterative_closest_point"]], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "has converged" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]
    rows = [ln.split() for ln in r.stdout.strip().splitlines()[-4:]]
    T = np.array([[float(v) for v in row] for row in rows])
    assert T.shape == (4, 4) and abs(T[0, 3] - 0.7) < 1e-5 and np.allclose(T[:3, :3], np.eye(3), atol=1e-5) and np.allclose(T[1:3, 3], 0, atol=1e-5)
    r = subprocess.run([built["kdtree_search"]], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "K nearest neighbor search" in r.stdout and "Neighbors within radius search" in r.stdout
    d2 = [float(ln.split("squared distance:")[1].strip(" )\n")) for ln in r.stdout.splitlines() if "squared distance" in ln][:10]
    assert len(d2) == 10 and d2 == sorted(d2)
    assert subprocess.run([built["pcd_write"]], capture_output=True, text=True, cwd=tmp_path).returncode == 0
    r = subprocess.run([built["pcd_read"]], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "Loaded" in r.stdout and len(r.stdout.strip().splitlines()) >= 6
    r = subprocess.run([built["concatenate_clouds"], "-f"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and len(r.stderr.strip().splitlines()[-1].split()) == 6     # x y z + the three normal components


TOOLS = "/root/reference/tools"


@pytest.mark.skipif(not os.path.isdir(TOOLS), reason="the reference checkout (its tools/) is not on this machine")
def test_reference_command_line_tools_compile_unchanged_and_run(double, golden, tmp_path):
    """The reference's own command-line tools for the path — tools/iterative_closest_point.cpp, voxel_grid.cpp, outlier_removal.cpp,
    cluster_extraction.cpp — compiled UNCHANGED from the reference checkout against the facade headers (pcl/console/{print,parse,
    time}.h, blob PCD I/O, VoxelGrid<PCLPointCloud2>, ExtractIndices, the rejector family, TransformationEstimationLM, ...) and run on
    bun0 / bun4 (on the test double): ICP converges and moves bun0 onto bun4, VoxelGrid at 0.02 gives the reference's 103 points."""
    _, lib, _ = double
    b0, b4 = _clouds(golden, tmp_path)
    exe = {}
    for name in ("iterative_closest_point", "voxel_grid", "outlier_removal", "cluster_extraction"):
        exe[name] = str(tmp_path / ("tool_" + name))
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + FACADE, "-I" + os.path.join(ROOT, "include"), os.path.join(TOOLS, name + ".cpp"),
                               "-o", exe[name], lib, "-Wl,-rpath," + os.path.dirname(lib), "-pthread"])

    def points_of(path):
        lines = open(path, "rb").read().split(b"\n")
        n = [int(ln.split()[1]) for ln in lines[:12] if ln.startswith(b"POINTS")][0]
        return n

    out = tmp_path / "aligned.pcd"
    r = subprocess.run([exe["iterative_closest_point"], b0, b4, str(out), "-debug", "1"], capture_output=True, text=True)
    assert r.returncode == 0 and "has converged: 1" in r.stdout, r.stdout[-1500:] + r.stderr[-800:]
    rows = r.stdout.split("Transformation is:")[1].strip().splitlines()[:4]
    T = np.array([[float(v) for v in row.split()] for row in rows])
    assert T.shape == (4, 4) and abs(np.linalg.det(T[:3, :3]) - 1) < 1e-4 and 0.4 < T[2, 0] < 0.7   # the ~30 degree turn about y between the two scans
    text = out.read_text().splitlines()
    body = np.array([[float(t) for t in ln.split()] for ln in text[11:] if ln.strip()])
    tgt = np.asarray(golden["bun4"], dtype=np.float64)[:, :3]

    def mean_nn(a):
        return np.sqrt(((a[:, None, :] - tgt[None, :, :]) ** 2).sum(-1).min(1)).mean()
    assert body.shape[0] == 397 and mean_nn(body[:, :3]) < 0.5 * mean_nn(np.asarray(golden["bun0"], dtype=np.float64)[:, :3])

    r = subprocess.run([exe["voxel_grid"], b0, str(tmp_path / "vg.pcd"), "-leaf", "0.02,0.02,0.02"], capture_output=True, text=True)
    assert r.returncode == 0 and points_of(tmp_path / "vg.pcd") == 103, r.stdout[-800:] + r.stderr[-400:]
    r = subprocess.run([exe["outlier_removal"], b0, str(tmp_path / "sor.pcd"), "-method", "statistical", "-mean_k", "8", "-std_dev_mul", "1.0"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and 200 < points_of(tmp_path / "sor.pcd") < 397, r.stdout[-800:] + r.stderr[-400:]
    r = subprocess.run([exe["outlier_removal"], b0, str(tmp_path / "ror.pcd"), "-method", "radius", "-radius", "0.01", "-min_pts", "4"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and 100 < points_of(tmp_path / "ror.pcd") < 397
    r = subprocess.run([exe["cluster_extraction"], b0, str(tmp_path / "cl.pcd"), "-tolerance", "0.01", "-min", "5"], capture_output=True, text=True)
    assert r.returncode == 0 and "clusters]" in r.stdout and os.path.exists(tmp_path / "cl0.pcd") and points_of(tmp_path / "cl0.pcd") > 300
