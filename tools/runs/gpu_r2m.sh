python -m pytest tests/test_gpu_parity.py tests/test_gpu_normals_corr.py -x -q -m gpu 2>&1 | tail -4
python tools/bench_stages.py > gpurun_out/r2m_stages.jsonl 2> gpurun_out/r2m_stages.err; tail -2 gpurun_out/r2m_stages.err
grep -E "knn_k|normals_knn" gpurun_out/r2m_stages.jsonl | cut -c1-130
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_knn_warp -c 1 -o gpurun_out/r2m_knn_warp -f python tools/iter_times.py 10000000 1 > gpurun_out/r2m_ncu.log 2>&1; tail -2 gpurun_out/r2m_ncu.log
