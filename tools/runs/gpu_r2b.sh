python tools/iter_times.py 10000000 12 > gpurun_out/r2b_iters.jsonl 2> gpurun_out/r2b_iters.err; tail -2 gpurun_out/r2b_iters.err
PCLB200_LIB=pcl_b200/libpclb200_stats.so python tools/iter_times.py 10000000 12 > gpurun_out/r2b_iters_stats.jsonl 2> gpurun_out/r2b_iters_stats.err; tail -2 gpurun_out/r2b_iters_stats.err
cat gpurun_out/r2b_iters.jsonl gpurun_out/r2b_iters_stats.jsonl
