#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
run() {
  tag=$1; rep=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --replicas $rep > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag rep $rep value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms", round(j["ms_per_step"],1), "grid", j["config"]["grid"], j["config"]["block"], "frac", round(j["roofline"]["frac"],4), "ok", j["config"]["success_paths"])
except Exception as e:
    print("$tag failed", e, open("gpurun_out/bench_$tag.err").read()[-600:])
PY
}
run tpp_base 148 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=4
run tpp_16w 148 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=8
run tpp_r592_l4 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=4
run tpp_r592_l8 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=8
run tpp_r592_l8_rf1 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=8 HC_B200_REFILL_MIN=1
run tpp_r592_l8_rf16 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=8 HC_B200_REFILL_MIN=16
run tpp_r592_b128 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=128 HC_B200_BLOCKS_PER_SM=4
run tpp_r592_b32 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=32 HC_B200_BLOCKS_PER_SM=16
run grp8_r592 592 HC_B200_ENGINE=group HC_B200_GROUP=8
