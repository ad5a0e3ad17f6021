# 2 GPUs: the 2-GPU product test + strong/weak scaling sanity of cfg3
python -m pytest tests/test_multirank_gloo.py -x -q -m gpu 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 4 --warmup 3 --scaling strong > gpurun_out/r2n_cfg3_strong_2gpu.json 2> gpurun_out/r2n_cfg3_strong_2gpu.err; tail -2 gpurun_out/r2n_cfg3_strong_2gpu.err
$TR bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2n_cfg3_weak_2gpu.json 2> gpurun_out/r2n_cfg3_weak_2gpu.err; tail -2 gpurun_out/r2n_cfg3_weak_2gpu.err
python - <<'PY'
import json
for f in ("r2n_cfg3_strong_2gpu","r2n_cfg3_weak_2gpu"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["breakdown_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
