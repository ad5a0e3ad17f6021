python -m pytest tests/test_gpu_parity.py tests/test_gpu_normals_corr.py tests/test_facade_gpu.py -x -q -m gpu 2>&1 | tail -6
python tools/bench_stages.py > gpurun_out/r2o_stages.jsonl 2> gpurun_out/r2o_stages.err; tail -2 gpurun_out/r2o_stages.err
grep -E "knn_k|normals_knn" gpurun_out/r2o_stages.jsonl | cut -c1-130
