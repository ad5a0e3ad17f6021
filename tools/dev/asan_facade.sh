#!/bin/bash
# Builds the test double of the C-ABI and the facade's programs (and, when /root/reference is present, four of the reference's
# tools) with -fsanitize=address,undefined and runs them on bun0 / bun4.  CPU only.  usage: tools/dev/asan_facade.sh [workdir]
set -eu
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
W=${1:-/tmp/pclb_asan}
mkdir -p "$W" && cd "$W"
F=$ROOT/pcl_b200/pcl_compat
SAN="-O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer"
python - "$ROOT" <<'PY'
import sys, numpy as np
root = sys.argv[1]
sys.path.insert(0, root + "/tests"); sys.path.insert(0, root)
from test_facade_gpu import _write_ascii_pcd, _write_binary_pcd, _write_golden
import pathlib
g = np.load(root + "/tests/golden/pcl_golden.npz")
_write_ascii_pcd(pathlib.Path("bun0.pcd"), g["bun0"]); _write_binary_pcd(pathlib.Path("bun4.pcd"), g["bun4"]); _write_golden("golden.txt", g)
PY
g++ $SAN -fPIC -shared -I$ROOT/include $ROOT/tests/host/pclb200_on_oracle.cpp -o libpclb200_on_oracle.so -L$ROOT/oracle -lpcl_oracle -Wl,-rpath,$ROOT/oracle
for t in tests/test_facade tests/test_facade_extra tests/test_host_api tests/test_pcd_io examples/iterative_closest_point; do
  g++ $SAN -I$F -I$ROOT/include $F/$t.cpp -o $(basename $t) ./libpclb200_on_oracle.so -Wl,-rpath,$W -pthread
done
export ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=print_stacktrace=1
run() { echo "== $*"; "$@" > run.log 2>&1 || { tail -20 run.log; echo "FAILED: $*"; exit 1; }; grep -i -E "sanitizer|runtime error" run.log && { echo "SANITIZER REPORT in: $*"; exit 1; } || true; tail -1 run.log; }
run ./test_facade bun0.pcd bun4.pcd golden.txt
run ./test_facade_extra bun0.pcd bun4.pcd golden.txt
run ./test_host_api
run ./test_host_api centroid bun0.pcd
mkdir -p d && run ./test_pcd_io selftest d
run ./iterative_closest_point bun0.pcd bun4.pcd out.pcd 50 0.05
if [ -d /root/reference/tools ]; then
  for t in iterative_closest_point voxel_grid outlier_removal cluster_extraction; do
    g++ $SAN -I$F -I$ROOT/include /root/reference/tools/$t.cpp -o tool_$t ./libpclb200_on_oracle.so -Wl,-rpath,$W -pthread
  done
  run ./tool_iterative_closest_point bun0.pcd bun4.pcd t1.pcd
  run ./tool_voxel_grid bun0.pcd t2.pcd -leaf 0.02,0.02,0.02
  run ./tool_outlier_removal bun0.pcd t3.pcd -method statistical -mean_k 8 -std_dev_mul 1.0
  run ./tool_outlier_removal bun0.pcd t4.pcd -method radius -radius 0.01 -min_pts 4
  run ./tool_cluster_extraction bun0.pcd t5.pcd -tolerance 0.01 -min 5
fi
echo "no sanitizer report, no leak"
