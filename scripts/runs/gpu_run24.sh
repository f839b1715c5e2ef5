#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
U2=$PWD/homotopycontinuation.jl_b200/libhc_b200_u2.so
for rep in 1 2; do
for wl in "katsura8 @592" "cyclic7_polyhedral @160" "biochem_sweep @256"; do
  set -- $wl
  echo -n "A(x4) "; timeout 600 python scripts/gpu_sweep2.py $1 "$2" 2>&1 | cut -c1-90
  echo -n "B(x2) "; HC_B200_LIB=$U2 timeout 600 python scripts/gpu_sweep2.py $1 "$2" 2>&1 | cut -c1-90
done; done | tee gpurun_out/ab24.txt
