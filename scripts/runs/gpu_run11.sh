#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
HC_B200_ENGINE=local timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_local.log 2>&1; echo "pytest local exit $?" ; tail -3 gpurun_out/pytest_local.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "ENGINE=tpp,BLOCK=64,BLOCKS_PER_SM=8@592" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@592" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@2368" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@592" "ENGINE=local,BLOCK=128,BLOCKS_PER_SM=2@592" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4,SEG_WINDOW=16@592" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4,SEG_WINDOW=64@592" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4,SEG_WINDOW=8@592" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=6@592" \
  2>&1 | tee gpurun_out/sweep11_katsura.txt
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "ENGINE=tpp,BLOCK=64,BLOCKS_PER_SM=8@160" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@160" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4@640" \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=4,SEG_WINDOW=64@160" \
  2>&1 | tee gpurun_out/sweep11_cyclic7.txt
timeout 300 python scripts/gpu_sweep2.py biochem_sweep \
  "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=8@256" "ENGINE=local,BLOCK=64,BLOCKS_PER_SM=16@256" \
  2>&1 | tee gpurun_out/sweep11_bio.txt
