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


def test_facade_extra_program_builds():
    """the device-side extra program compiles against the C-ABI here (it runs under -m gpu)"""
    from pcl_b200 import build
    build.build()
    subprocess.check_call(["make", "-C", FACADE, "-s", "tests/test_facade_extra"])
