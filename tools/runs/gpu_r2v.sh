python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "config2_1M" 2>&1 | tail -8
for v in _acb4 _acb5 _acb6; do
  PCLB200_LIB=pcl_b200/libpclb200$v.so python tools/iter_times.py 10000000 10 > gpurun_out/r2v_iter$v.jsonl 2> gpurun_out/r2v_iter$v.err; tail -1 gpurun_out/r2v_iter$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
rows=[json.loads(l) for l in open(f"gpurun_out/r2v_iter{v}.jsonl") if '"iter"' in l and '"rep": 1' in l]
print(v, [r["accum_ms"] for r in rows], "sum=%.3f"%sum(r["accum_ms"] for r in rows), rows[-1]["n_corr"], rows[-1]["mse"])
PY
done
