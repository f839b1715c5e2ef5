#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 2-GPU bench"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --replicas 240 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1800 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample 16384 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; tail -c 600 gpurun_out/bench_2gpu_ref.json
echo "== 1-GPU same config"
timeout 900 python bench.py --steps 3 --warmup 3 --replicas 240 --no-cpu-baseline > gpurun_out/bench_1gpu_240.json 2> gpurun_out/bench_1gpu_240.err; tail -c 1500 gpurun_out/bench_1gpu_240.json
