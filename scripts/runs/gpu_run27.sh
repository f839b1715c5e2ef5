#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python scripts/gpu_sweep2.py cyclic7_polyhedral "@160" "BLOCK=160@160" "BLOCK=192@160" "SEG_WINDOW=64@160" "SEG_WINDOW=1000@160" "BLOCK=160,SEG_WINDOW=1000@160" "SEG_WINDOW=16@160" "BLOCK=160@640" 2>&1 | tee gpurun_out/sweep27_cyclic7.txt | cut -c1-110
timeout 900 python scripts/gpu_sweep2.py katsura8 "@592" "BLOCK=160@592" "BLOCK=96@592" 2>&1 | tee gpurun_out/sweep27_katsura.txt | cut -c1-110
