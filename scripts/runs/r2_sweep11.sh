mkdir -p gpurun_out
export HC_B200_JIT=1
run() { echo "== $*"; env "$@" python tests/tools/gpu_run_once.py cyclic7_polyhedral 480 3 2>&1 | grep -E "paths/s" | tail -2; }
run HC_B200_JIT_PREFETCH=0
run HC_B200_JIT_PREFETCH=2
run HC_B200_JIT_PREFETCH=0
run HC_B200_JIT_PREFETCH=3
