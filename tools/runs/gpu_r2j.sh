python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
python bench.py --workload cfg4 --points 4000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_cfg4_small.json 2> gpurun_out/r2j_cfg4_small.err; tail -3 gpurun_out/r2j_cfg4_small.err; cut -c1-900 gpurun_out/r2j_cfg4_small.json
python bench.py --workload cfg5 --points 20000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_cfg5_small.json 2> gpurun_out/r2j_cfg5_small.err; tail -3 gpurun_out/r2j_cfg5_small.err; cut -c1-1200 gpurun_out/r2j_cfg5_small.json
python tools/iter_times.py 10000000 10 2>/dev/null | tail -3
