mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -1; }
{
for k in 0.5 1 2; do run HC_B200_REFILL_K=$k cyclic7_polyhedral 160 2; run HC_B200_REFILL_K=$k katsura8 592 2; run HC_B200_REFILL_K=$k biochem_sweep 512 2; done
run HC_B200_REFILL_MIN=8 cyclic7_polyhedral 160 2; run HC_B200_REFILL_MIN=8 katsura8 592 2; run HC_B200_REFILL_MIN=8 biochem_sweep 512 2
run X=1 tritangents 1 1; run X=1 cyclooctane_td 1 1; run X=1 cyclooctane_polyhedral 1 1
timeout 900 python -m pytest tests/test_jit.py tests/test_gpu_parity.py -q -m gpu -k "bit_identical or deterministic or two_pass or refill" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2b_refill3.txt
