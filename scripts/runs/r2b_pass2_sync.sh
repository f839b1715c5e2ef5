mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -1; }
{
run HC_B200_HANDOFF_SYNC=1 tritangents 1 1; run HC_B200_HANDOFF_SYNC=0 tritangents 1 1
run HC_B200_HANDOFF_SYNC=1 cyclooctane_td 1 1; run HC_B200_HANDOFF_SYNC=0 cyclooctane_td 1 1
run HC_B200_HANDOFF_SYNC=1 cyclooctane_polyhedral 1 1
run HC_B200_HANDOFF_SYNC=1 HC_B200_HANDOFF_GROUP=32 tritangents 1 1; run HC_B200_HANDOFF_SYNC=1 HC_B200_HANDOFF_GROUP=8 cyclooctane_td 1 1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_pass" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2b_pass2_sync.txt
