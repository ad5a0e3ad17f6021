PCLB200_LIB=pcl_b200/libpclb200_ef0d15.so python tools/knn_check.py 10000000 10000000 16 2>gpurun_out/r3a_chk.err | cut -c1-1500 | tee gpurun_out/r3a_chk.jsonl; tail -1 gpurun_out/r3a_chk.err | cut -c1-200
PCLB200_LIB=pcl_b200/libpclb200_ef0d15.so python tools/knn_times.py 2>>gpurun_out/r3a_chk.err | tee gpurun_out/r3a_knn.jsonl
PCLB200_LIB=pcl_b200/libpclb200_ef0d15.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "knn or normals" 2>&1 | tail -3
