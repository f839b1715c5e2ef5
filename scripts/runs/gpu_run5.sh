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
run tpp_2w 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=1
run tpp_4w 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=2
run tpp_8w 592 HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=4
for bps in 4 8; do
HC_B200_ENGINE=tpp HC_B200_BLOCK=64 HC_B200_BLOCKS_PER_SM=$bps timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_track -s 1 -c 1 -o gpurun_out/prof_tpp_bps$bps \
   python bench.py --steps 1 --warmup 1 --replicas 592 --no-cpu-baseline > gpurun_out/ncu_tpp_$bps.log 2>&1
tail -1 gpurun_out/ncu_tpp_$bps.log
done
