mkdir -p gpurun_out
(timeout 900 python tests/tools/gpu_tail_probe.py tritangents 110592 200; timeout 600 python tests/tools/gpu_tail_probe.py cyclooctane_td 32768 300) 2>&1 | tee gpurun_out/r2b_tail_probe.txt
