#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
export PYTHONUNBUFFERED=1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:hc_track -s 1 -c 1 -o /tmp/ncu/k8 -f python bench.py --steps 1 --warmup 1 --workload katsura8 --replicas 592 --no-cpu-baseline > gpurun_out/ncu_v3_k8.log 2>&1; tail -2 gpurun_out/ncu_v3_k8.log
ncu -i /tmp/ncu/k8.ncu-rep --page raw --csv > gpurun_out/ncu_v3_k8_raw.csv 2>/dev/null
ncu -i /tmp/ncu/k8.ncu-rep --page source --csv --print-source cuda,sass > /tmp/ncu/k8_src.csv 2>/dev/null
python scripts/ncu_by_function.py /tmp/ncu/k8_src.csv | cut -c1-170 > gpurun_out/ncu_v3_k8_by_function.txt
gzip -c /tmp/ncu/k8_src.csv > gpurun_out/ncu_v3_k8_src.csv.gz
timeout 1200 ncu --set full --clock-control none -k regex:hc_track -s 1 -c 1 -o /tmp/ncu/c7 -f python bench.py --steps 1 --warmup 1 --replicas 160 --no-cpu-baseline > gpurun_out/ncu_v3_c7.log 2>&1; tail -2 gpurun_out/ncu_v3_c7.log
ncu -i /tmp/ncu/c7.ncu-rep --page raw --csv > gpurun_out/ncu_v3_c7_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
