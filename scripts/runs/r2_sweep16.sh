mkdir -p gpurun_out
export HC_B200_JIT=1
run() { echo "== $W $*"; env "$@" python tests/tools/gpu_run_once.py $W $R 2 2>&1 | grep -E "paths/s" | tail -1; }
W=cyclic7_polyhedral R=480
run HC_B200_JIT_FLAGS=
run HC_B200_JIT_FLAGS=--fmad=false
W=katsura8 R=1184
run HC_B200_JIT_FLAGS=
run HC_B200_JIT_FLAGS=--fmad=false
W=biochem_sweep R=256
run HC_B200_JIT_FLAGS=
run HC_B200_JIT_FLAGS=--fmad=false
