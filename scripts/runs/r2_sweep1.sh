mkdir -p gpurun_out
export HC_B200_JIT=1
for cfg in "128 0" "128 1" "64 0" "64 1" "96 1" "256 1" "32 1"; do
  set -- $cfg
  echo "== block $1 sync $2" 
  HC_B200_JIT_BLOCK=$1 HC_B200_SYNC_CTA=$2 python tests/tools/gpu_run_once.py cyclic7_polyhedral 160 2 2>&1 | grep -v "^\[hc_b200\] program" | tail -1
done
