#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_default.log 2>&1; echo "pytest default exit $?" ; tail -3 gpurun_out/pytest_default.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 700 gpurun_out/final_ref.json
timeout 1200 python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err; tail -c 2600 gpurun_out/final_default.json; tail -2 gpurun_out/final_default.err
timeout 900 python bench.py --workload katsura8 --replicas 1184 --steps 3 > gpurun_out/final_katsura8.json 2> gpurun_out/final_katsura8.err; tail -c 900 gpurun_out/final_katsura8.json
timeout 900 python bench.py --workload biochem_sweep --replicas 512 --steps 3 > gpurun_out/final_bio.json 2> gpurun_out/final_bio.err; tail -c 900 gpurun_out/final_bio.json
