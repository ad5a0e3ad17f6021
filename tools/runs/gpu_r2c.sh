# ncu full capture of k_search at a converged iteration (launch 22 of 24) and an early one (launch 13 = rep1 iter1)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_search -s 13 -c 1 -o gpurun_out/r2c_search_early -f python tools/iter_times.py 10000000 12 > gpurun_out/r2c_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_search -s 22 -c 1 -o gpurun_out/r2c_search_late -f python tools/iter_times.py 10000000 12 > gpurun_out/r2c_ncu2.log 2>&1
tail -3 gpurun_out/r2c_ncu2.log
ls -la gpurun_out/*.ncu-rep
