#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "BLOCK=64,BLOCKS_PER_SM=2@592" "BLOCK=64,BLOCKS_PER_SM=4@592" "BLOCK=64,BLOCKS_PER_SM=3@592" "BLOCK=32,BLOCKS_PER_SM=2@592" "BLOCK=64,BLOCKS_PER_SM=1@592" \
  "BLOCK=64,BLOCKS_PER_SM=2,SEG_WINDOW=32@592" "BLOCK=64,BLOCKS_PER_SM=2,SEG_WINDOW=128@592" "BLOCK=64,BLOCKS_PER_SM=2@2368" \
  2>&1 | tee gpurun_out/sweep14_katsura.txt
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "BLOCK=64,BLOCKS_PER_SM=2@160" "BLOCK=64,BLOCKS_PER_SM=4@160" "BLOCK=64,BLOCKS_PER_SM=2@640" \
  2>&1 | tee gpurun_out/sweep14_cyclic7.txt
timeout 300 python scripts/gpu_sweep2.py biochem_sweep \
  "BLOCK=64,BLOCKS_PER_SM=2@256" "BLOCK=64,BLOCKS_PER_SM=4@256" "BLOCK=64,BLOCKS_PER_SM=8@256" \
  2>&1 | tee gpurun_out/sweep14_bio.txt
