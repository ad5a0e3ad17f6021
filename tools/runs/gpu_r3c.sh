for v in "" _l10 _l14; do
  PCLB200_LIB=pcl_b200/libpclb200$v.so python tools/knn_times.py 2>gpurun_out/r3c_knn$v.err | tee -a gpurun_out/r3c_knn.jsonl; tail -1 gpurun_out/r3c_knn$v.err | cut -c1-200
done
