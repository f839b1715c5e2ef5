mkdir -p gpurun_out
export HC_B200_JIT=1
run() { echo "== $W $*"; env "$@" python tests/tools/gpu_run_once.py $W $R 2 2>&1 | grep -E "paths/s" | tail -1; }
W=cyclic7_polyhedral R=480
run A=1
W=katsura8 R=1184
run A=1
W=biochem_sweep R=256
run A=1
REPS=8 timeout 600 python tests/tools/gpu_jit_check.py cyclic7_polyhedral katsura8 2>&1 | grep "jit=1" | head -4
