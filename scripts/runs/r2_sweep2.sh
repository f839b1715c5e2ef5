mkdir -p gpurun_out
export HC_B200_JIT=1
REPS=160 timeout 900 python tests/tools/gpu_jit_check.py cyclic7_polyhedral katsura8 2>&1 | grep -v "^\[hc_b200\] program"
for cfg in "128 0" "128 1" "256 1" "256 0" "192 1" "64 1"; do
  set -- $cfg
  echo "== block $1 jit_sync $2"
  HC_B200_JIT_BLOCK=$1 HC_B200_JIT_SYNC=$2 python tests/tools/gpu_run_once.py cyclic7_polyhedral 160 2 2>&1 | grep -v "^\[hc_b200\] program" | tail -1
done
