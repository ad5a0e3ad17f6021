for v in "" _ef0 _d15 _ef0d15; do
  PCLB200_LIB=pcl_b200/libpclb200$v.so python tools/knn_times.py 2>gpurun_out/r2x_knn$v.err | tee -a gpurun_out/r2x_knn.jsonl; tail -1 gpurun_out/r2x_knn$v.err | cut -c1-200
done
