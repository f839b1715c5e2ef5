#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export HC_B200_LIB=$PWD/homotopycontinuation.jl_b200/libhc_b200_dbg.so
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@592" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@2368" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@592" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@2368" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=2@592" "ENGINE=local,BLOCK=32,BLOCKS_PER_SM=4@592" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=16@2368" "ENGINE=local,BLOCK=128,BLOCKS_PER_SM=2@592" \
  "ENGINE=local,BLOCK=32,BLOCKS_PER_SM=2@592" "ENGINE=local,BLOCK=32,BLOCKS_PER_SM=1@592" \
  2>&1 | tee gpurun_out/sweep10_katsura.txt
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@160" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@160" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@640" \
  2>&1 | tee gpurun_out/sweep10_cyclic7.txt
timeout 300 python scripts/gpu_sweep2.py biochem_sweep \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@64" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=16@256" "ENGINE=tpp,BLOCK=64,BLOCKS_PER_SM=16@256" \
  2>&1 | tee gpurun_out/sweep10_bio.txt
