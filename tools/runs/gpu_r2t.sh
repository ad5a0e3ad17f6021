for v in "" _defer; do
  lib=pcl_b200/libpclb200$v.so
  PCLB200_LIB=$lib python tools/iter_times.py 10000000 10 > gpurun_out/r2t_iter$v.jsonl 2> gpurun_out/r2t_iter$v.err; tail -1 gpurun_out/r2t_iter$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
rows=[json.loads(l) for l in open(f"gpurun_out/r2t_iter{v}.jsonl") if '"iter"' in l and '"rep": 1' in l]
print(v or "default", [r["search_ms"] for r in rows], "sum=%.2f"%sum(r["search_ms"] for r in rows), rows[-1]["n_corr"], rows[-1]["mse"])
PY
done
