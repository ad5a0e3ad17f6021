PCLB200_LIB=pcl_b200/libpclb200_stats.so python tools/iter_times.py 10000000 12 > gpurun_out/r2r_stats.jsonl 2> gpurun_out/r2r_stats.err; tail -2 gpurun_out/r2r_stats.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open("gpurun_out/r2r_stats.jsonl") if '"iter"' in l]
for r in rows: print({k:r.get(k) for k in ("iter","nodes","leaves","lookups","node_lane_utilisation","node_visits_p50_p90_p99")})
PY
