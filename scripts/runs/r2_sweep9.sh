mkdir -p gpurun_out
export HC_B200_VERBOSE=1
for w in "tritangents 1" "cyclooctane_td 1" "biochem_sweep 256" "katsura8 1184" "cyclic7_polyhedral 480"; do
  set -- $w
  for jit in 0 1; do
    echo "== $1 x$2 jit=$jit"
    HC_B200_JIT=$jit timeout 900 python tests/tools/gpu_run_once.py $1 $2 2 2>&1 | grep -E "specialised kernel|paths/s" | tail -3
  done
done
