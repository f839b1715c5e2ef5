mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -3; }
{
run X=1 cyclooctane_polyhedral 1 3; run HC_B200_HANDOFF_SYNC=0 cyclooctane_polyhedral 1 3; run X=1 cyclooctane_td 1 3; run X=1 tritangents 1 3
} 2>&1 | tee gpurun_out/r2b_final_check.txt
