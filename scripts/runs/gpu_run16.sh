#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_cyclic7.csv python bench.py --steps 2 --warmup 1 --replicas 160 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -5 gpurun_out/launches_cyclic7.csv | cut -c1-300
echo "== cyclooctane engines"; timeout 900 python scripts/gpu_sweep2.py cyclooctane_td "ENGINE=group@1" "TPP_MAX_N=20,BLOCK=64,BLOCKS_PER_SM=2@1" "TPP_MAX_N=20,BLOCK=64,BLOCKS_PER_SM=4@1" 2>&1 | tee gpurun_out/sweep16_cyclo.txt | cut -c1-330
echo "== tritangents parity"; timeout 1500 python tests/tools/gpu_parity_large.py tritangents 16384 2>&1 | tee gpurun_out/parity_trit.txt
echo "== katsura parity"; timeout 600 python tests/tools/gpu_parity_large.py katsura8 256 2>&1 | tee gpurun_out/parity_k8.txt
