set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json
