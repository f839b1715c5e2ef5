#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== local (assume)"; HC_B200_ENGINE=local timeout 300 python tests/tools/gpu_debug3.py 2>&1 | tail -16
echo "=== local (generic ld/st)"; HC_B200_LIB=$PWD/homotopycontinuation.jl_b200/libhc_b200_dbg.so HC_B200_ENGINE=local timeout 300 python tests/tools/gpu_debug3.py 2>&1 | tail -16
echo "=== memcheck local (assume)"; HC_B200_ENGINE=local timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tests/tools/gpu_debug3.py 2>&1 | grep -v "^$" | tail -40
echo "=== ncu tpp katsura8"
HC_B200_ENGINE=tpp timeout 900 ncu --set full --import-source on --clock-control none -k regex:hc_track -s 1 -c 1 -o gpurun_out/ncu_tpp_k8 -f python bench.py --steps 1 --warmup 1 --workload katsura8 --replicas 148 --no-cpu-baseline > gpurun_out/ncu_tpp_k8.log 2>&1; tail -2 gpurun_out/ncu_tpp_k8.log; ls -la gpurun_out/*.ncu-rep
