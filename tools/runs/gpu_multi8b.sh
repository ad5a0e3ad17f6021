TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR bench.py --gpus 8 --workload cfg5 --scaling strong --steps 4 --warmup 3 > gpurun_out/r3n_cfg5_strong_8gpu.json 2> gpurun_out/r3n_cfg5_strong_8gpu.err; tail -2 gpurun_out/r3n_cfg5_strong_8gpu.err
timeout 300 $TR bench.py --gpus 8 --steps 4 --warmup 3 --scaling strong > gpurun_out/r3n_cfg3_strong_8gpu.json 2> gpurun_out/r3n_cfg3_strong_8gpu.err; tail -2 gpurun_out/r3n_cfg3_strong_8gpu.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $TR4 bench.py --gpus 4 --workload cfg4 --scaling strong --steps 4 --warmup 3 > gpurun_out/r3n_cfg4_strong_4gpu.json 2> gpurun_out/r3n_cfg4_strong_4gpu.err; tail -2 gpurun_out/r3n_cfg4_strong_4gpu.err
python - <<'PY'
import json
for f in ("r3n_cfg5_strong_8gpu","r3n_cfg3_strong_8gpu","r3n_cfg4_strong_4gpu"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["breakdown_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
