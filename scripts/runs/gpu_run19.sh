#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "BLOCK=64,BLOCKS_PER_SM=4@592" "BLOCK=64,BLOCKS_PER_SM=4,CARVEOUT=100@592" "BLOCK=64,BLOCKS_PER_SM=4,CARVEOUT=0@592" \
  "BLOCK=128,BLOCKS_PER_SM=2@592" "BLOCK=256,BLOCKS_PER_SM=1@592" "BLOCK=128,BLOCKS_PER_SM=1@592" "BLOCK=192,BLOCKS_PER_SM=1@592" \
  2>&1 | tee gpurun_out/sweep19_katsura.txt | cut -c1-200
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "BLOCK=64,BLOCKS_PER_SM=4@160" "BLOCK=64,BLOCKS_PER_SM=4,CARVEOUT=100@160" "BLOCK=256,BLOCKS_PER_SM=1@160" "BLOCK=128,BLOCKS_PER_SM=2@160" \
  2>&1 | tee gpurun_out/sweep19_cyclic7.txt | cut -c1-200
timeout 300 python scripts/gpu_sweep2.py biochem_sweep "BLOCK=64,BLOCKS_PER_SM=8@256" "BLOCK=256,BLOCKS_PER_SM=2@256" "BLOCK=256,BLOCKS_PER_SM=1@256" 2>&1 | tee gpurun_out/sweep19_bio.txt | cut -c1-200
