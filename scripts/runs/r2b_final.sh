mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu ) > gpurun_out/r2b_gpu_tests.txt 2>&1; tail -5 gpurun_out/r2b_gpu_tests.txt
bash scripts/runs/r2b_bench1.sh
