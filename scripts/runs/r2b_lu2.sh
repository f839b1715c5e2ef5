# two-column register LU (n >= 10) vs the one-column one: first pass of the heavy-tailed configs
mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -2; }
{
run HC_B200_JIT_LU_ONE_COLUMN=0 tritangents 1 2; run HC_B200_JIT_LU_ONE_COLUMN=1 tritangents 1 2
run HC_B200_JIT_LU_ONE_COLUMN=0 cyclooctane_td 1 2; run HC_B200_JIT_LU_ONE_COLUMN=1 cyclooctane_td 1 2
run HC_B200_JIT_LU_ONE_COLUMN=0 cyclooctane_polyhedral 1 2
timeout 900 python -m pytest tests/test_jit.py tests/test_gpu_parity.py -q -m gpu -k "bit_identical or two_pass" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2b_lu2.txt
