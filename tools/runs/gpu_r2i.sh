python bench.py --steps 5 --warmup 3 > gpurun_out/r2i_bench_1gpu.json 2> gpurun_out/r2i_bench_1gpu.err; tail -2 gpurun_out/r2i_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; tail -2 gpurun_out/r2i_bench_ref.err
# ncu: DRAM traffic of one converged-regime k_search launch + one k_accum_dmma launch, and the kNN/normals kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_search|k_accum_dmma' -s 36 -c 2 -o gpurun_out/r2i_search_accum -f python tools/iter_times.py 10000000 10 > gpurun_out/r2i_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_normals -c 1 -o gpurun_out/r2i_normals16 -f python tools/iter_times.py 10000000 1 > gpurun_out/r2i_ncu2.log 2>&1
cat gpurun_out/r2i_bench_1gpu.json; cat gpurun_out/r2i_bench_ref.json
