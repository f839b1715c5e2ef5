# compute-sanitizer on the lockstep specialised kernel and the interpreter kernels (small batches)
mkdir -p gpurun_out
export HC_B200_JIT=1
for tool in memcheck racecheck synccheck; do
  echo "== $tool, specialised kernel (katsura(8) x 8 = 2048 paths)"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tests/tools/gpu_run_once.py katsura8 8 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|paths/s|Error|hazard" | head -8
done
echo "== memcheck, specialised kernel, polyhedral (cyclic-7 x 2)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tests/tools/gpu_run_once.py cyclic7_polyhedral 2 1 2>&1 | grep -E "ERROR SUMMARY|paths/s|Error" | head -5
export HC_B200_JIT=0
echo "== memcheck, interpreter kernel (katsura(8) x 2)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tests/tools/gpu_run_once.py katsura8 2 1 2>&1 | grep -E "ERROR SUMMARY|paths/s|Error" | head -5
