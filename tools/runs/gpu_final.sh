# full GPU suite the driver runs at round end, smoke(), the default bench, and the ncu launch list of the same command
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2p_bench_1gpu.json 2> gpurun_out/r2p_bench_1gpu.err; tail -2 gpurun_out/r2p_bench_1gpu.err; cut -c1-400 gpurun_out/r2p_bench_1gpu.json
python tools/config_table.py > gpurun_out/r2p_config_table.jsonl 2> gpurun_out/r2p_config_table.err; cat gpurun_out/r2p_config_table.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_b_ncu.log 2>&1; tail -1 gpurun_out/r2p_b_ncu.log | cut -c1-200
