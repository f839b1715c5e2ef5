mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -2; }
{
run HC_B200_VERBOSE=1 tritangents 1 1; run HC_B200_HANDOFF_SOLO=0 tritangents 1 1
run HC_B200_VERBOSE=1 HC_B200_HANDOFF_GROUP=8 cyclooctane_td 1 1; run X=1 cyclooctane_td 1 1

timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_pass or group_engine or polyhedral_starts" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2b_solo.txt
