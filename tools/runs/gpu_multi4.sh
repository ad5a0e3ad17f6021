TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 4 --workload cfg4 --scaling strong --steps 4 --warmup 3 > gpurun_out/r2n_cfg4_strong_4gpu.json 2> gpurun_out/r2n_cfg4_strong_4gpu.err; tail -2 gpurun_out/r2n_cfg4_strong_4gpu.err
$TR bench.py --gpus 4 --steps 4 --warmup 3 --scaling strong > gpurun_out/r2n_cfg3_strong_4gpu.json 2> gpurun_out/r2n_cfg3_strong_4gpu.err; tail -2 gpurun_out/r2n_cfg3_strong_4gpu.err
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
$TR2 bench.py --gpus 2 --steps 4 --warmup 3 --scaling strong > gpurun_out/r2n_cfg3_strong_2gpu.json 2> gpurun_out/r2n_cfg3_strong_2gpu.err; tail -2 gpurun_out/r2n_cfg3_strong_2gpu.err
python - <<'PY'
import json
for f in ("r2n_cfg4_strong_4gpu","r2n_cfg3_strong_4gpu","r2n_cfg3_strong_2gpu"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["breakdown_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
