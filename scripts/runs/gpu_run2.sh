#!/bin/bash
# v2 smoke: parity tests + bench sweep over group sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for G in 32 16 8; do
  HC_B200_GROUP=$G timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_g$G.json 2> gpurun_out/bench_g$G.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_g$G.json").read().strip().splitlines()[-1])
    print("G=$G value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms", j["ms_per_step"], "grid", j["config"]["grid"], j["config"]["block"], "frac", j["roofline"]["frac"], "ok", j["config"]["success_paths"])
except Exception as e:
    print("G=$G failed", e, open("gpurun_out/bench_g$G.err").read()[-800:])
PY
done
