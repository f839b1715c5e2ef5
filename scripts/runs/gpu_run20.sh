#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "BLOCK=128,BLOCKS_PER_SM=1@592" "BLOCK=128,BLOCKS_PER_SM=1,CARVEOUT=0@592" "BLOCK=96,BLOCKS_PER_SM=1@592" "BLOCK=160,BLOCKS_PER_SM=1@592" "BLOCK=64,BLOCKS_PER_SM=2,CARVEOUT=0@592" "BLOCK=128,BLOCKS_PER_SM=1@2368" "BLOCK=128,BLOCKS_PER_SM=2,CARVEOUT=0@592" \
  2>&1 | tee gpurun_out/sweep20_katsura.txt | cut -c1-200
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "BLOCK=128,BLOCKS_PER_SM=1@160" "BLOCK=128,BLOCKS_PER_SM=1,CARVEOUT=0@160" "BLOCK=64,BLOCKS_PER_SM=4,CARVEOUT=0@160" "BLOCK=96,BLOCKS_PER_SM=1@160" "BLOCK=160,BLOCKS_PER_SM=1@160" "BLOCK=64,BLOCKS_PER_SM=4@160" "BLOCK=128,BLOCKS_PER_SM=2,CARVEOUT=0@160" "BLOCK=128,BLOCKS_PER_SM=1@640" \
  2>&1 | tee gpurun_out/sweep20_cyclic7.txt | cut -c1-200
timeout 600 python scripts/gpu_sweep2.py tritangents "BLOCK=128,BLOCKS_PER_SM=1@1" "BLOCK=64,BLOCKS_PER_SM=4@1" 2>&1 | tee gpurun_out/sweep20_trit.txt | cut -c1-200
