#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
timeout 900 python bench.py --workload biochem_sweep --replicas 512 --steps 3 --no-cpu-baseline > gpurun_out/final_bio2.json 2> gpurun_out/final_bio2.err; python - <<'PY'
import json
j=json.loads(open("gpurun_out/final_bio2.json").read().strip().splitlines()[-1]); print("bio value", round(j["value"]), "e2e", round(j["e2e"]["value"]))
PY
