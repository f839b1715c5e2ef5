#!/bin/bash
# ncu --set full of the lane-group engine on the first 4096 cyclooctane (total degree) paths + class-count parity vs the oracle
mkdir -p gpurun_out /tmp/ncu
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hc_track -c 1 -o /tmp/ncu/cy -f python tests/tools/gpu_parity_large.py cyclooctane_td 4096 > gpurun_out/ncu_s5_cyclo.log 2>&1; tail -8 gpurun_out/ncu_s5_cyclo.log
ncu -i /tmp/ncu/cy.ncu-rep --page raw --csv > gpurun_out/ncu_s5_cyclo_raw.csv 2>/dev/null
ncu -i /tmp/ncu/cy.ncu-rep --page source --csv --print-source cuda,sass > /tmp/ncu/cy_src.csv 2>/dev/null
python scripts/ncu_by_function.py /tmp/ncu/cy_src.csv | cut -c1-170 > gpurun_out/ncu_s5_cyclo_by_function.txt
head -45 gpurun_out/ncu_s5_cyclo_by_function.txt
