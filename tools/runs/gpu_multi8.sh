TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
$TR bench.py --gpus 8 --workload cfg5 --scaling strong --steps 4 --warmup 3 > gpurun_out/r2n_cfg5_strong_8gpu.json 2> gpurun_out/r2n_cfg5_strong_8gpu.err; tail -2 gpurun_out/r2n_cfg5_strong_8gpu.err
$TR bench.py --gpus 8 --steps 4 --warmup 3 --scaling strong > gpurun_out/r2n_cfg3_strong_8gpu.json 2> gpurun_out/r2n_cfg3_strong_8gpu.err; tail -2 gpurun_out/r2n_cfg3_strong_8gpu.err
$TR bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2n_cfg3_weak_8gpu.json 2> gpurun_out/r2n_cfg3_weak_8gpu.err; tail -2 gpurun_out/r2n_cfg3_weak_8gpu.err
python - <<'PY'
import json
for f in ("r2n_cfg5_strong_8gpu","r2n_cfg3_strong_8gpu","r2n_cfg3_weak_8gpu"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["breakdown_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
