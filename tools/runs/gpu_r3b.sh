for v in "" _t12 _t17; do
  PCLB200_LIB=pcl_b200/libpclb200$v.so python tools/knn_times.py 2>gpurun_out/r3b_knn$v.err | tee -a gpurun_out/r3b_knn.jsonl; tail -1 gpurun_out/r3b_knn$v.err | cut -c1-200
done
python tools/knn_check.py 10000000 10000000 16,32 2>gpurun_out/r3b_chk.err | cut -c1-1500 | tee gpurun_out/r3b_chk.jsonl; tail -1 gpurun_out/r3b_chk.err | cut -c1-200
python -m pytest tests/test_gpu_parity.py tests/test_gpu_normals_corr.py -x -q -m gpu 2>&1 | tail -3
