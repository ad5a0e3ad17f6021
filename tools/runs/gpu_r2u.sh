python tools/debug_cfg2.py 1000000 50 > gpurun_out/r2u_cfg2_debug.jsonl 2> gpurun_out/r2u_cfg2_debug.err; tail -3 gpurun_out/r2u_cfg2_debug.err; head -c 6000 gpurun_out/r2u_cfg2_debug.jsonl | cut -c1-330
for v in _acc3 _acc4; do
  PCLB200_LIB=pcl_b200/libpclb200$v.so python tools/iter_times.py 10000000 10 > gpurun_out/r2u_iter$v.jsonl 2> gpurun_out/r2u_iter$v.err; tail -1 gpurun_out/r2u_iter$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
rows=[json.loads(l) for l in open(f"gpurun_out/r2u_iter{v}.jsonl") if '"iter"' in l and '"rep": 1' in l]
print(v, [r["accum_ms"] for r in rows], rows[-1]["n_corr"], rows[-1]["mse"])
PY
done
PCLB200_LIB=pcl_b200/libpclb200_acc3.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_normals_corr.py -x -q -m gpu 2>&1 | tail -3
