#!/bin/bash
# round-1 (session 4): local-memory thread-per-path engine vs global slabs, occupancy and paths-per-lane sweeps
mkdir -p gpurun_out
HC_B200_ENGINE=local timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_local.log 2>&1; echo "pytest local exit $?" ; tail -3 gpurun_out/pytest_local.log
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "ENGINE=tpp,BLOCK=64,BLOCKS_PER_SM=8@592" "ENGINE=tpp,BLOCK=64,BLOCKS_PER_SM=8@2368" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@592" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@2368" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@592" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@2368" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=2@592" "ENGINE=local,BLOCK=32,BLOCKS_PER_SM=4@592" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=16@2368" "ENGINE=local,BLOCK=128,BLOCKS_PER_SM=2@592" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=6@592" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=3@592" \
  2>&1 | tee gpurun_out/sweep8_katsura.txt
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "ENGINE=tpp,BLOCK=64,BLOCKS_PER_SM=8@160" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@160" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@160" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@640" \
  2>&1 | tee gpurun_out/sweep8_cyclic7.txt
timeout 300 python scripts/gpu_sweep2.py biochem_sweep \
  "ENGINE=tpp,BLOCK=64,BLOCKS_PER_SM=8@64" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@64" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=16@256" \
  2>&1 | tee gpurun_out/sweep8_bio.txt
