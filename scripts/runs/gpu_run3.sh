#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms", round(j["ms_per_step"],1), "grid", j["config"]["grid"], j["config"]["block"], "frac", round(j["roofline"]["frac"],4), "ok", j["config"]["success_paths"])
except Exception as e:
    print("$tag failed", e, open("gpurun_out/bench_$tag.err").read()[-600:])
PY
}
run g8 HC_B200_GROUP=8
run g8c8 HC_B200_GROUP=8 HC_B200_ROUND_CAP=8
run g8c32 HC_B200_GROUP=8 HC_B200_ROUND_CAP=32
run g16 HC_B200_GROUP=16
run g16c16 HC_B200_GROUP=16 HC_B200_ROUND_CAP=16
run g32 HC_B200_GROUP=32
run g32c32 HC_B200_GROUP=32 HC_B200_ROUND_CAP=32
HC_B200_GROUP=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_track_kernel -s 1 -c 1 -o gpurun_out/prof_v2_g8 \
   python bench.py --steps 1 --warmup 1 --replicas 32 --no-cpu-baseline > gpurun_out/ncu_v2.log 2>&1
tail -2 gpurun_out/ncu_v2.log
