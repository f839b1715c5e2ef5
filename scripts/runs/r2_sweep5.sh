mkdir -p gpurun_out
export HC_B200_JIT=1 HC_B200_JIT_LU_SMEM=0
run() { echo "== $*"; env "$@" python tests/tools/gpu_run_once.py cyclic7_polyhedral 320 2 2>&1 | grep -v "^\[hc_b200\] program" | tail -1; }
run HC_B200_JIT_BLOCK=128 HC_B200_BLOCKS_PER_SM=2
run HC_B200_JIT_BLOCK=64 HC_B200_BLOCKS_PER_SM=4
run HC_B200_JIT_BLOCK=256 HC_B200_REFILL_MIN=4
run HC_B200_JIT_BLOCK=256 HC_B200_REFILL_MIN=12
run HC_B200_JIT_BLOCK=256 HC_B200_REFILL_MIN=2
run HC_B200_JIT_BLOCK=256 HC_B200_CARVEOUT=0
