#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/final_2gpu.json 2> gpurun_out/final_2gpu.err; python - <<'PY'
import json
j=json.loads(open("gpurun_out/final_2gpu.json").read().strip().splitlines()[-1]); print("2gpu value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms", j["ms_per_step"], j["config"]["workload"], j["config"]["class_counts"])
PY
tail -2 gpurun_out/final_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample 32768 2>/dev/null | tail -1 | cut -c1-300
