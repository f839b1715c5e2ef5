#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "REFILL_MIN=8@592" "REFILL_MIN=1@592" "REFILL_MIN=4@592" "REFILL_MIN=16@592" "REFILL_MIN=32@592" "SEG_WINDOW=16@592" "SEG_WINDOW=64@592" \
  2>&1 | tee gpurun_out/sweep21_katsura.txt | cut -c1-200
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "REFILL_MIN=8@160" "REFILL_MIN=1@160" "REFILL_MIN=16@160" "BLOCK=160@160" "SEG_WINDOW=64@160" "SEG_WINDOW=16@160" \
  2>&1 | tee gpurun_out/sweep21_cyclic7.txt | cut -c1-200
timeout 300 python scripts/gpu_sweep2.py biochem_sweep "REFILL_MIN=8@256" "BLOCK=128@256" "BLOCK=256,BLOCKS_PER_SM=2@256" 2>&1 | tee gpurun_out/sweep21_bio.txt | cut -c1-200
