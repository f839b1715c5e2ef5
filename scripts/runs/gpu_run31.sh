#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
for wl in "katsura8 POST_FRAC=0@592" "katsura8 POST_FRAC=16@592" "katsura8 POST_FRAC=24@592" "katsura8 POST_FRAC=28@592" "katsura8 POST_FRAC=32@592" "cyclic7_polyhedral POST_FRAC=0@160" "cyclic7_polyhedral POST_FRAC=16@160" "cyclic7_polyhedral POST_FRAC=24@160" "cyclic7_polyhedral POST_FRAC=32@160" "biochem_sweep POST_FRAC=0@256" "biochem_sweep POST_FRAC=24@256" "tritangents POST_FRAC=24@1"; do
  set -- $wl
  timeout 600 python scripts/gpu_sweep2.py $1 "$2" 2>&1 | cut -c1-100
done | tee gpurun_out/sweep31.txt
