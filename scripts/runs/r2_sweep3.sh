mkdir -p gpurun_out
export HC_B200_JIT=1
for cfg in "128" "192" "256"; do
  echo "== block $cfg"
  HC_B200_JIT_BLOCK=$cfg python tests/tools/gpu_run_once.py cyclic7_polyhedral 320 2 2>&1 | grep -v "^\[hc_b200\] program" | tail -1
done
HC_B200_JIT_BLOCK=256 python tests/tools/gpu_run_once.py katsura8 800 2 2>&1 | tail -1
HC_B200_JIT_BLOCK=128 python tests/tools/gpu_run_once.py katsura8 800 2 2>&1 | tail -1
