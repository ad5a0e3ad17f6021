"""Builds pcl_b200/libpclb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pcl_b200.build [--force] [--verbose]

-fmad=false: device arithmetic must round like the reference's scalar C++ (no fused multiply-add), because
parity with the CPU path is bit-exact for indices and distances (see DESIGN.md "numerics").
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libpclb200.so")
SOURCES = ["alloc.cu", "lbvh.cu", "search.cu", "icp.cu", "voxel.cu", "reject.cu", "normals_corr.cu", "cluster.cu", "comm.cu", "capi.cu"]
HEADERS = ["internal.cuh", "traverse.cuh", "knn_warp.cuh", "search_kernels.cuh", "lbvh_kernels.cuh", "voxel_kernels.cuh", "reject_kernels.cuh", "normals_corr_kernels.cuh", "cluster_kernels.cuh", "icp_kernels.cuh", "corr_select.cuh", os.path.join("..", "..", "include", "pclb200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v", "--expt-relaxed-constexpr",
         "-ccbin", "/usr/bin/g++"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, leaf=None, tag=None, defines=()):
    """leaf/tag: build an experimental variant (libpclb200_<tag>.so with -DPCLB_LEAF=<leaf>) next to the default."""
    global OBJ, SO
    flags = list(FLAGS)
    if tag:
        OBJ = os.path.join(HERE, "build_" + tag)
        SO = os.path.join(HERE, f"libpclb200_{tag}.so")
    if leaf:
        flags += [f"-DPCLB_LEAF={int(leaf)}"]
    flags += [f"-D{d}" for d in defines]
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _newer(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([NVCC] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        else:
            with open(obj + ".ptxas.log", "w") as f:
                f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on " + src)
        return obj

    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _newer(SO, objs):
        r = subprocess.run([NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                           "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC", "-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return SO


if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        if a.startswith("--leaf="):
            kw["leaf"] = int(a.split("=")[1])
        if a.startswith("--tag="):
            kw["tag"] = a.split("=")[1]
        if a.startswith("--define="):
            kw.setdefault("defines", []).append(a.split("=", 1)[1])
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, **kw))
