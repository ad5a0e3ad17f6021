python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -3 gpurun_out/r2w_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2w_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"], d["breakdown_ms_per_step"], d["setup_ms"], d["roofline"]["iteration"])
PY
