#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "BLOCK=64,BLOCKS_PER_SM=4@592" "BLOCK=64,BLOCKS_PER_SM=4@2368" \
  "BLOCK=64,BLOCKS_PER_SM=6@592" "BLOCK=128,BLOCKS_PER_SM=2@592" "BLOCK=128,BLOCKS_PER_SM=3@592" "BLOCK=64,BLOCKS_PER_SM=2@592" \
  "BLOCK=64,BLOCKS_PER_SM=4,SEG_WINDOW=16@592" "BLOCK=64,BLOCKS_PER_SM=4,SEG_WINDOW=64@592" \
  2>&1 | tee gpurun_out/sweep12_katsura.txt
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "BLOCK=64,BLOCKS_PER_SM=4@160" "BLOCK=64,BLOCKS_PER_SM=4@640" "BLOCK=64,BLOCKS_PER_SM=6@160" \
  "BLOCK=64,BLOCKS_PER_SM=4,SEG_WINDOW=64@160" \
  2>&1 | tee gpurun_out/sweep12_cyclic7.txt
timeout 300 python scripts/gpu_sweep2.py biochem_sweep \
  "BLOCK=64,BLOCKS_PER_SM=4@256" "BLOCK=64,BLOCKS_PER_SM=6@256" \
  2>&1 | tee gpurun_out/sweep12_bio.txt
timeout 600 python scripts/gpu_sweep2.py tritangents "BLOCK=64,BLOCKS_PER_SM=4@1" 2>&1 | tee gpurun_out/sweep12_trit.txt
