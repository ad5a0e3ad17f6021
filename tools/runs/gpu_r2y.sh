for v in "" _d15 _ef0d15; do
  PCLB200_LIB=pcl_b200/libpclb200$v.so python tools/knn_check.py 2>gpurun_out/r2y_chk$v.err | cut -c1-1500 | tee -a gpurun_out/r2y_chk.jsonl; tail -1 gpurun_out/r2y_chk$v.err | cut -c1-200
done
