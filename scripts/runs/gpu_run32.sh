#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
export PYTHONUNBUFFERED=1
for spec in "c7 cyclic7_polyhedral 160" "k8 katsura8 592"; do
  set -- $spec
  timeout 1200 ncu --set full --import-source on --clock-control none -k regex:hc_track -s 1 -c 1 -o /tmp/ncu/$1 -f python bench.py --steps 1 --warmup 1 --workload $2 --replicas $3 --no-cpu-baseline > gpurun_out/ncu_v5_$1.log 2>&1; tail -1 gpurun_out/ncu_v5_$1.log
  ncu -i /tmp/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu_v5_$1_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$1.ncu-rep --page source --csv --print-source cuda,sass > /tmp/ncu/$1_src.csv 2>/dev/null
  python scripts/ncu_by_function.py /tmp/ncu/$1_src.csv | cut -c1-170 > gpurun_out/ncu_v5_$1_by_function.txt
  rm -f /tmp/ncu/$1.ncu-rep /tmp/ncu/$1_src.csv
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_v5_cyclic7.csv python bench.py --steps 2 --warmup 1 --replicas 160 --no-cpu-baseline > gpurun_out/launches_v5.log 2>&1; tail -4 gpurun_out/launches_v5_cyclic7.csv | cut -c1-250
ls -la gpurun_out | tail -8
