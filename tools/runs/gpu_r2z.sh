for v in "" _ef0d15; do
  PCLB200_LIB=pcl_b200/libpclb200$v.so python tools/knn_check.py 10000000 10000000 16 2>gpurun_out/r2z_chk$v.err | cut -c1-2500 | tee -a gpurun_out/r2z_chk.jsonl; tail -1 gpurun_out/r2z_chk$v.err | cut -c1-200
done
