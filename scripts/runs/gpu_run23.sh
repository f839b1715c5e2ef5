#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
timeout 900 python scripts/gpu_sweep2.py katsura8 "@592" "BLOCK=160@592" "@2368" 2>&1 | tee gpurun_out/sweep23_katsura.txt | cut -c1-200
timeout 600 python scripts/gpu_sweep2.py cyclic7_polyhedral "@160" "BLOCK=160@160" "@640" 2>&1 | tee gpurun_out/sweep23_cyclic7.txt | cut -c1-200
timeout 300 python scripts/gpu_sweep2.py biochem_sweep "@256" 2>&1 | tee gpurun_out/sweep23_bio.txt | cut -c1-200
timeout 600 python scripts/gpu_sweep2.py tritangents "@1" 2>&1 | tee gpurun_out/sweep23_trit.txt | cut -c1-300
