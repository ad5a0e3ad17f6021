python -m pytest tests/test_gpu_parity.py tests/test_gpu_consumers.py tests/test_multirank_gloo.py -x -q -m gpu 2>&1 | tail -12
for v in default wq mb5 pf; do
  if [ $v = default ]; then L=pcl_b200/libpclb200.so; else L=pcl_b200/libpclb200_$v.so; fi
  PCLB200_LIB=$L python tools/iter_times.py 10000000 12 > gpurun_out/r2g_$v.jsonl 2> gpurun_out/r2g_$v.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_icp_wq -s 22 -c 1 -o gpurun_out/r2g_wq_late -f env PCLB200_LIB=pcl_b200/libpclb200_wq.so python tools/iter_times.py 10000000 12 > gpurun_out/r2g_ncu.log 2>&1
python - <<'PY'
import json
for f in ("default","wq","mb5","pf"):
    rows=[json.loads(l) for l in open(f"gpurun_out/r2g_{f}.jsonl") if '"iter"' in l]
    print(f, [r["search_ms"] for r in rows], "sum10=%.2f"%sum(r["search_ms"]+r["accum_ms"] for r in rows[:10]))
PY
