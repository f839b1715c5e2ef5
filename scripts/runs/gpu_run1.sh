#!/bin/bash
# first GPU pass of a session: parity tests, bench (both arms), launch list, one full ncu capture
set -x
mkdir -p gpurun_out
nproc; nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_track_kernel -s 1 -c 1 -o gpurun_out/prof_r01b \
   python bench.py --steps 1 --warmup 1 --replicas 148 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
