"""Host-only behaviour of the facade (pcl_b200/pcl_compat/tests/test_host_api.cpp): the pcl::PointCloud container's
width / height bookkeeping, the scenarios of the reference's test/common/test_pointcloud.cpp.  CPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACADE = os.path.join(ROOT, "pcl_b200", "pcl_compat")


def test_point_cloud_container_and_host_classes():
    subprocess.check_call(["make", "-C", FACADE, "-s", "tests/test_host_api"])
    r = subprocess.run([os.path.join(FACADE, "tests", "test_host_api")], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-2000:]


def test_reference_centroid_tests(golden, tmp_path):
    """The bodies of the reference's test/common/test_centroid.cpp for the moment functions on the path: compute3DCentroid
    (float, double; empty / all-NaN inputs leave the caller's centroid untouched), computeMeanAndCovarianceMatrix,
    demeanPointCloud on bun0 with the reference's expected values; and of test/common/test_transforms.cpp: transformPointCloud
    dense / indexed / with a NaN point, transformPointCloudWithNormals dense / indexed, Matrix4f and Matrix4d, 10 epsilon."""
    import numpy as np
    from test_facade_gpu import _write_ascii_pcd
    subprocess.check_call(["make", "-C", FACADE, "-s", "tests/test_host_api"])
    _write_ascii_pcd(tmp_path / "bun0.pcd", np.asarray(golden["bun0"], dtype=np.float32))
    r = subprocess.run([os.path.join(FACADE, "tests", "test_host_api"), "centroid", str(tmp_path / "bun0.pcd")], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-2000:]


def test_facade_extra_program_builds():
    """the device-side extra program compiles against the C-ABI here (it runs under -m gpu)"""
    from pcl_b200 import build
    build.build()
    subprocess.check_call(["make", "-C", FACADE, "-s", "tests/test_facade_extra"])


def test_host_point_normal_matches_oracle_and_golden(golden, orc, tmp_path):
    """pcl::computePointNormal / computeMeanAndCovarianceMatrix / eigen33 / solvePlaneParameters /
    flipNormalTowardsViewpoint of the facade (host side, pcl/common/*.h, pcl/features/*.h): 200 index subsets of bun0
    bit for bit against the oracle's restatement, and the reference's golden values of
    test/features/test_normal_estimation.cpp:103-138."""
    import numpy as np
    from test_facade_gpu import _write_ascii_pcd
    subprocess.check_call(["make", "-C", FACADE, "-s", "tests/test_host_api"])
    bun0 = np.asarray(golden["bun0"], dtype=np.float32)
    _write_ascii_pcd(tmp_path / "bun0.pcd", bun0)
    out = tmp_path / "normals.bin"
    subprocess.check_call([os.path.join(FACADE, "tests", "test_host_api"), "normals", str(tmp_path / "bun0.pcd"), str(out)])
    v = np.fromfile(out, dtype=np.float32)
    n = bun0.shape[0]
    cloud = orc.to_xyz1(bun0)
    for s in range(200):
        idx = np.array([(s * 37 + j * (s % 5 + 1)) % n for j in range(3 + s % 40)], dtype=np.int32)
        o, ok = orc.point_normal(cloud, idx)
        g = v[5 * s: 5 * s + 5]
        assert ok
        assert np.array_equal(g[[0, 1, 2, 4]].view(np.uint32), o.view(np.uint32)), (s, g, o)
    t = v[1000:]
    gold = np.asarray(golden["normal_bun0"], dtype=np.float64)      # nx ny nz d curvature
    assert np.allclose(np.abs(t[0:3]), gold[0:3], atol=1e-4) and abs(abs(t[3]) - abs(gold[3])) < 1e-4 and abs(t[4] - gold[4]) < 1e-4
    assert np.allclose(t[5:10], gold, atol=1e-4)                     # computePointNormal(cloud, plane, curvature)
    assert np.allclose(t[10:14], [-0.035592, -0.369596, -0.928511, 0.0799743], atol=1e-4)   # flipped towards the origin
    assert abs(t[14] + 0.035592) < 1e-4
    assert np.allclose(np.abs(t[15:18]), gold[0:3], atol=1e-4) and abs(t[18] - gold[4]) < 1e-4
