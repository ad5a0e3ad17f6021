for v in default hn; do
  if [ $v = default ]; then L=pcl_b200/libpclb200.so; else L=pcl_b200/libpclb200_$v.so; fi
  PCLB200_LIB=$L python tools/iter_times.py 10000000 12 > gpurun_out/r2q_$v.jsonl 2> gpurun_out/r2q_$v.err
done
python - <<'PY'
import json
for f in ("default","hn"):
    rows=[json.loads(l) for l in open(f"gpurun_out/r2q_{f}.jsonl") if '"iter"' in l]
    print(f, [r["search_ms"] for r in rows], "sum10=%.2f"%sum(r["search_ms"]+r["accum_ms"] for r in rows[:10]))
PY
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "per_iteration or icp" 2>&1 | tail -3
PCLB200_LIB=pcl_b200/libpclb200_hn.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "per_iteration or icp" 2>&1 | tail -3
