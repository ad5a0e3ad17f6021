"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header declares,
and refuses to run without a GPU (no CPU fallback).  No compute calls."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from pcl_b200 import build
    return build.build()


def _header_symbols():
    h = open(os.path.join(ROOT, "include", "pclb200.h")).read()
    return sorted(set(re.findall(r"PCLB200_API\s+[\w\s\*]+?\b(pclb200_\w+)\s*\(", h)))


def test_header_and_binding_agree():
    import pcl_b200
    assert _header_symbols() == sorted(pcl_b200.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = set(re.findall(r"\bT (pclb200_\w+)", out))
    missing = [s for s in _header_symbols() if s not in exported]
    assert not missing, missing
    import pcl_b200
    L = pcl_b200.lib()
    assert L.pclb200_version() == 100


def test_library_is_sm100a_native(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(built):
    import torch
    import pcl_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pcl_b200.Pclb200Error) as e:
        pcl_b200.Context(0)
    assert e.value.code == pcl_b200.ERR_CUDA


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pcl_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "pcl_oracle" not in txt, f


def test_default_params_layout(built):
    """pclb200_icp_default_params is pure host code: its values, read back through the ctypes mirror of the struct,
    prove that the Python layout of pclb200_icp_params matches the header (field order, padding, trailing fields)."""
    import numpy as np
    import pcl_b200 as P
    p = P.default_params()
    assert p.max_iterations == 10 and p.estimator == P.EST_SVD and p.is_dense == 1
    assert p.enforce_same_direction_normals == 1 and p.correspondence_kind == P.CORR_NEAREST
    assert p.max_correspondence_distance == np.sqrt(np.finfo(np.float64).max)
    assert p.euclidean_fitness_epsilon == -np.finfo(np.float64).max
    assert p.mse_threshold_absolute == 1e-12
    assert p.correspondence_k == 10 and p.track_mode == 0


def test_library_sources_read_no_environment_variable():
    """A drop-in library must not change behaviour on environment variables: every switch is a params field or an explicit
    call.  (The .so still imports getenv — the statically linked CUDA runtime reads its own CUDA_* variables.)"""
    import glob
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pcl_b200", "csrc")
    offenders = [f for f in glob.glob(os.path.join(root, "*.cu*")) if "getenv" in open(f).read()]
    assert not offenders, offenders
