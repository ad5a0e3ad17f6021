"""Runs the C++ facade test program (pcl_b200/pcl_compat/tests/test_facade.cpp): the reference's own
registration / kdtree / filters / features tests written against the drop-in pcl:: classes.  GPU only."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACADE = os.path.join(ROOT, "pcl_b200", "pcl_compat")


def _write_ascii_pcd(path, xyz):
    with open(path, "w") as f:
        f.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n"
                f"COUNT 1 1 1\nWIDTH {len(xyz)}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(xyz)}\nDATA ascii\n")
        for p in xyz:
            f.write("%.9g %.9g %.9g\n" % (p[0], p[1], p[2]))


def _write_binary_pcd(path, xyz):
    with open(path, "wb") as f:
        f.write(("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n"
                 f"COUNT 1 1 1\nWIDTH {len(xyz)}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(xyz)}\nDATA binary\n").encode())
        f.write(np.ascontiguousarray(xyz, dtype=np.float32).tobytes())


def test_facade_builds_on_cpu():
    """-m 'not gpu': the header-only facade and its test program compile against the C-ABI."""
    from pcl_b200 import build
    build.build()
    subprocess.check_call(["make", "-C", FACADE, "-s"])
    assert os.path.exists(os.path.join(FACADE, "tests", "test_facade"))


def _write_golden(path, golden):
    """The golden vectors the facade programs compare against, as text: name, count, values."""
    with open(path, "w") as f:
        for k in ("corr_original", "corr_reciprocal", "icp_bun0_bun4", "svd_Tref", "normal_bun0", "corr_rej_dist",
                  "corr_rej_median", "corr_rej_one_to_one", "corr_rej_trimmed", "corr_rej_sac", "sac_transform"):
            v = np.asarray(golden[k], dtype=np.float64).ravel()
            f.write(f"{k} {v.size}\n" + " ".join("%.17g" % x for x in v) + "\n")


@pytest.mark.gpu
def test_facade_reference_tests(golden, tmp_path):
    subprocess.check_call(["make", "-C", FACADE, "-s"])
    _write_ascii_pcd(tmp_path / "bun0.pcd", golden["bun0"])      # both PCD encodings go through the reader
    _write_binary_pcd(tmp_path / "bun4.pcd", golden["bun4"])
    _write_golden(tmp_path / "golden.txt", golden)
    r = subprocess.run([os.path.join(FACADE, "tests", "test_facade"), str(tmp_path / "bun0.pcd"),
                        str(tmp_path / "bun4.pcd"), str(tmp_path / "golden.txt")], capture_output=True, text=True)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "PASSED" in r.stdout
