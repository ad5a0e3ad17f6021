python -m pytest tests/test_gpu_parity.py tests/test_gpu_normals_corr.py tests/test_gpu_consumers.py -x -q -m gpu 2>&1 | tail -6
python tools/bench_stages.py > gpurun_out/r2l_stages.jsonl 2> gpurun_out/r2l_stages.err; tail -2 gpurun_out/r2l_stages.err
grep -E "knn_k|normals" gpurun_out/r2l_stages.jsonl | cut -c1-160
python bench.py --workload cfg5 --points 20000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_cfg5_small.json 2> gpurun_out/r2l_cfg5_small.err; tail -3 gpurun_out/r2l_cfg5_small.err; cut -c1-1600 gpurun_out/r2l_cfg5_small.json
