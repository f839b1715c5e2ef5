#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
HC_B200_BLOCKS_PER_SM=2 timeout 1200 ncu --set full --import-source on --clock-control none -k regex:hc_track -s 1 -c 1 -o gpurun_out/ncu_tpl_k8 -f python bench.py --steps 1 --warmup 1 --workload katsura8 --replicas 592 --no-cpu-baseline > gpurun_out/ncu_tpl_k8.log 2>&1; tail -2 gpurun_out/ncu_tpl_k8.log; ls -la gpurun_out/*.ncu-rep
