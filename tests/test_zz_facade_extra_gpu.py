"""Second facade test program (pcl_b200/pcl_compat/tests/test_facade_extra.cpp): the pcl::search::Search overloads,
CorrespondenceEstimation::setPointRepresentation[Reciprocal] and the DefaultConvergenceCriteria thresholds — facade
surface added after the main facade program's last run on hardware.  Named to run last."""
import os
import subprocess

import pytest

from test_facade_gpu import FACADE, _write_ascii_pcd, _write_binary_pcd


@pytest.mark.gpu
def test_facade_extra_api(golden, tmp_path):
    subprocess.check_call(["make", "-C", FACADE, "-s", "tests/test_facade_extra"])
    _write_ascii_pcd(tmp_path / "bun0.pcd", golden["bun0"])
    _write_binary_pcd(tmp_path / "bun4.pcd", golden["bun4"])
    r = subprocess.run([os.path.join(FACADE, "tests", "test_facade_extra"), str(tmp_path / "bun0.pcd"),
                        str(tmp_path / "bun4.pcd")], capture_output=True, text=True)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "PASSED" in r.stdout
