python bench.py --workload cfg4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_cfg4_1gpu.json 2> gpurun_out/r2n_cfg4_1gpu.err; tail -2 gpurun_out/r2n_cfg4_1gpu.err
python bench.py --workload cfg5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_cfg5_1gpu.json 2> gpurun_out/r2n_cfg5_1gpu.err; tail -2 gpurun_out/r2n_cfg5_1gpu.err
python - <<'PY'
import json
for f in ("r2n_cfg4_1gpu","r2n_cfg5_1gpu"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["breakdown_ms_per_step"], d["setup_ms"], d["config"])
    except Exception as e: print(f, "ERR", e)
PY
