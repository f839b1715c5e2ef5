#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python scripts/gpu_sweep2.py katsura8 \
  "SEG_WINDOW_EVAL=1@592" "SEG_WINDOW_EVAL=32@592" "SEG_WINDOW_EVAL=1,SEG_WINDOW=64@592" "SEG_WINDOW_EVAL=1,BLOCKS_PER_SM=2@592" "SEG_WINDOW_EVAL=1,BLOCK=96,BLOCKS_PER_SM=1@592" "SEG_WINDOW_EVAL=1,BLOCK=96,BLOCKS_PER_SM=2@592" \
  2>&1 | tee gpurun_out/sweep17_katsura.txt | cut -c1-200
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral \
  "SEG_WINDOW_EVAL=1@160" "SEG_WINDOW_EVAL=32@160" "SEG_WINDOW_EVAL=1,SEG_WINDOW=64@160" "SEG_WINDOW_EVAL=1,BLOCKS_PER_SM=6@160" \
  2>&1 | tee gpurun_out/sweep17_cyclic7.txt | cut -c1-200
timeout 300 python scripts/gpu_sweep2.py biochem_sweep "SEG_WINDOW_EVAL=1@256" "SEG_WINDOW_EVAL=32@256" 2>&1 | tee gpurun_out/sweep17_bio.txt | cut -c1-200
timeout 600 python scripts/gpu_sweep2.py tritangents "SEG_WINDOW_EVAL=1@1" "SEG_WINDOW_EVAL=32@1" 2>&1 | tee gpurun_out/sweep17_trit.txt | cut -c1-300
