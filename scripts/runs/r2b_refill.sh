mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -1; }
{
for r in 3 5 8 11 16; do run HC_B200_REFILL_MIN=$r cyclic7_polyhedral 160 2; run HC_B200_REFILL_MIN=$r katsura8 592 2; run HC_B200_REFILL_MIN=$r biochem_sweep 512 2; done
} 2>&1 | tee gpurun_out/r2b_refill.txt
