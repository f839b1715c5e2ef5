#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
for wl in "katsura8 @592" "katsura8 REFILL_MIN=4@592" "katsura8 REFILL_MIN=16@592" "cyclic7_polyhedral @160" "cyclic7_polyhedral REFILL_MIN=4@160" "cyclic7_polyhedral REFILL_MIN=16@160" "cyclic7_polyhedral REFILL_MIN=24@160" "biochem_sweep @256" "biochem_sweep REFILL_MIN=16@256" "tritangents @1"; do
  set -- $wl
  timeout 600 python scripts/gpu_sweep2.py $1 "$2" 2>&1 | cut -c1-100
done | tee gpurun_out/sweep30.txt
