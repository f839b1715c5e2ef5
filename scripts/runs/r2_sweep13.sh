mkdir -p gpurun_out
export HC_B200_JIT=1
run() { echo "== $*"; env "$@" python tests/tools/gpu_run_once.py cyclic7_polyhedral 480 2 2>&1 | grep -E "paths/s" | tail -1; }
run HC_B200_JIT_HOIST=0
run HC_B200_JIT_HOIST=24
run HC_B200_JIT_HOIST=48
run HC_B200_JIT_HOIST=96
run HC_B200_JIT_HOIST=1000
