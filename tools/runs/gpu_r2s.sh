python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_edges.py tests/test_multirank_gloo.py tests/test_facade_gpu.py tests/test_gpu_normals_corr.py -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -2 gpurun_out/r2s_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["breakdown_ms_per_step"], d["gpu_launches"])
PY
python bench.py --workload cfg5 --points 20000000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('cfg5small', d['ms_per_step'], d['breakdown_ms_per_step'])"
