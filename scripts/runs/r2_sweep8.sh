mkdir -p gpurun_out
export HC_B200_JIT=1
REPS=8 timeout 600 python tests/tools/gpu_jit_check.py cyclic7_polyhedral katsura8 2>&1 | grep -v "^\[hc_b200\] program" | grep "jit=1"
run() { echo "== $*"; env "$@" python tests/tools/gpu_run_once.py cyclic7_polyhedral 320 2 2>&1 | grep -v "^\[hc_b200\] program" | tail -1; }
run HC_B200_JIT_PREFETCH=0
run HC_B200_JIT_PREFETCH=1
echo "== katsura prefetch 0/1"
HC_B200_JIT_PREFETCH=0 python tests/tools/gpu_run_once.py katsura8 800 2 2>&1 | tail -1
HC_B200_JIT_PREFETCH=1 python tests/tools/gpu_run_once.py katsura8 800 2 2>&1 | tail -1
