mkdir -p gpurun_out
export HC_B200_JIT=1
run() { echo "== $*"; env "$@" timeout 300 python tests/tools/gpu_run_once.py tritangents 1 1 2>&1 | grep -E "paths/s" | tail -1; }
run HC_B200_JIT_BLOCK=256
run HC_B200_JIT_BLOCK=128
run HC_B200_JIT_BLOCK=64
run HC_B200_JIT_BLOCK=32
run HC_B200_JIT_BLOCK=64 HC_B200_BLOCKS_PER_SM=2
run HC_B200_JIT_BLOCK=128 HC_B200_JIT_SYNC=0
run HC_B200_JIT_BLOCK=64 HC_B200_JIT_SYNC=0 HC_B200_BLOCKS_PER_SM=2
