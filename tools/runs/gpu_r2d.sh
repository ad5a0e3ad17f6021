python tools/iter_times.py 10000000 12 > gpurun_out/r2d_b0.jsonl 2> gpurun_out/r2d_b0.err
PCLB200_LIB=pcl_b200/libpclb200_b1.so python tools/iter_times.py 10000000 12 > gpurun_out/r2d_b1.jsonl 2> gpurun_out/r2d_b1.err
PCLB200_LIB=pcl_b200/libpclb200_stats.so python tools/iter_times.py 10000000 12 > gpurun_out/r2d_b1_stats.jsonl 2> gpurun_out/r2d_b1_stats.err
python - <<'PY'
import json
for f in ("r2d_b0","r2d_b1","r2d_b1_stats"):
    rows=[json.loads(l) for l in open(f"gpurun_out/{f}.jsonl") if '"iter"' in l]
    print(f, [r["search_ms"] for r in rows], "sum10=%.2f"%sum(r["search_ms"] for r in rows[:10]))
    if "lookups" in rows[0]:
        for r in rows: print({k:r[k] for k in ("iter","lookups","nodes","leaves","pushes","home_seeds","cells_pushed")})
PY
