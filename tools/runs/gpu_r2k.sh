python -m pytest tests/test_gpu_parity.py tests/test_gpu_normals_corr.py tests/test_gpu_consumers.py tests/test_gpu_api_edges.py -x -q -m gpu 2>&1 | tail -8
python tools/bench_stages.py > gpurun_out/r2k_stages.jsonl 2> gpurun_out/r2k_stages.err; tail -2 gpurun_out/r2k_stages.err
grep -E "knn_k|normals|index_build" gpurun_out/r2k_stages.jsonl | cut -c1-200
