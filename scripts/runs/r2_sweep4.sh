mkdir -p gpurun_out
export HC_B200_JIT=1
REPS=8 timeout 600 python tests/tools/gpu_jit_check.py cyclic7_polyhedral katsura8 2>&1 | grep -v "^\[hc_b200\] program" | grep "jit=1"
for cfg in "256 1" "256 0" "192 1" "128 1"; do
  set -- $cfg
  echo "== block $1 lu_smem $2"
  HC_B200_JIT_BLOCK=$1 HC_B200_JIT_LU_SMEM=$2 python tests/tools/gpu_run_once.py cyclic7_polyhedral 320 2 2>&1 | grep -v "^\[hc_b200\] program" | tail -1
done
for cfg in "256 1" "256 0"; do
  set -- $cfg
  echo "== katsura block $1 lu_smem $2"
  HC_B200_JIT_BLOCK=$1 HC_B200_JIT_LU_SMEM=$2 python tests/tools/gpu_run_once.py katsura8 800 2 2>&1 | tail -1
done
