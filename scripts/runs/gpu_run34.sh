#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
export PYTHONUNBUFFERED=1
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:hc_track -c 1 -o /tmp/ncu/tr -f python tests/tools/gpu_parity_large.py tritangents 16384 > gpurun_out/ncu_v5_trit.log 2>&1; tail -3 gpurun_out/ncu_v5_trit.log
ncu -i /tmp/ncu/tr.ncu-rep --page raw --csv > gpurun_out/ncu_v5_trit_raw.csv 2>/dev/null
ncu -i /tmp/ncu/tr.ncu-rep --page source --csv --print-source cuda,sass > /tmp/ncu/tr_src.csv 2>/dev/null
python scripts/ncu_by_function.py /tmp/ncu/tr_src.csv | cut -c1-170 > gpurun_out/ncu_v5_trit_by_function.txt
head -40 gpurun_out/ncu_v5_trit_by_function.txt
