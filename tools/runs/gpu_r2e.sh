python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
python tools/iter_times.py 10000000 12 > gpurun_out/r2e_iters.jsonl 2> gpurun_out/r2e_iters.err; tail -2 gpurun_out/r2e_iters.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open("gpurun_out/r2e_iters.jsonl") if '"iter"' in l]
print([r["search_ms"] for r in rows], "sum10=%.2f"%sum(r["search_ms"] for r in rows[:10]))
print([r["accum_ms"] for r in rows])
print(rows[-1])
PY
