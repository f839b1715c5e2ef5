#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; wl=$2; rep=$3; shift; shift; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload $wl --replicas $rep > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms", round(j["ms_per_step"],1), j["config"]["engine"], "grid", j["config"]["grid"], j["config"]["block"], "frac", round(j["roofline"]["frac"],4), j["config"]["class_counts"], "flops/path", round(j["roofline"]["flops_per_path"]))
except Exception as e:
    print("$tag failed", e, open("gpurun_out/bench_$tag.err").read()[-800:])
PY
}
run cyc7poly_tpp cyclic7_polyhedral 160 HC_B200_ENGINE=tpp
run cyc7poly_g8 cyclic7_polyhedral 40 HC_B200_ENGINE=group HC_B200_GROUP=8
run trit_tpp tritangents 1 HC_B200_ENGINE=tpp
run trit_g32 tritangents 1 HC_B200_ENGINE=group HC_B200_GROUP=32
run cyclo_tpp cyclooctane_td 1 HC_B200_ENGINE=tpp
run cyclo_g32 cyclooctane_td 1 HC_B200_ENGINE=group HC_B200_GROUP=32
run bio_tpp biochem_sweep 64 HC_B200_ENGINE=tpp
run bio_g8 biochem_sweep 16 HC_B200_ENGINE=group HC_B200_GROUP=8
