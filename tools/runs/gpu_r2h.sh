python -m pytest tests/test_gpu_parity.py tests/test_gpu_consumers.py tests/test_gpu_api_edges.py -x -q -m gpu 2>&1 | tail -12
for v in default mb5 mb6; do
  if [ $v = default ]; then L=pcl_b200/libpclb200.so; else L=pcl_b200/libpclb200_$v.so; fi
  PCLB200_LIB=$L python tools/iter_times.py 10000000 12 > gpurun_out/r2h_$v.jsonl 2> gpurun_out/r2h_$v.err
done
python - <<'PY'
import json
for f in ("default","mb5","mb6"):
    rows=[json.loads(l) for l in open(f"gpurun_out/r2h_{f}.jsonl") if '"iter"' in l]
    print(f, [r["search_ms"] for r in rows], [r["accum_ms"] for r in rows][-2:], "sum10=%.2f"%sum(r["search_ms"]+r["accum_ms"] for r in rows[:10]))
PY
