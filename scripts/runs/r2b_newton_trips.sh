# Newton trips per lockstep round: 0 = unlimited (the previous behaviour), 2 (default), 3
mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -${@: -1}; }
{
for v in 0 2 3; do run HC_B200_JIT_NEWTON_TRIPS=$v cyclic7_polyhedral 160 2; run HC_B200_JIT_NEWTON_TRIPS=$v katsura8 592 2; done
run X=1 biochem_sweep 512 2
run X=1 tritangents 1 1
run X=1 cyclooctane_td 1 1
timeout 1500 python -m pytest tests/test_jit.py tests/test_gpu_parity.py tests/test_monodromy.py -x -q -m gpu -k "jit or two_pass or bit_identical or polyhedral or monodromy" 2>&1 | tail -5
} 2>&1 | tee gpurun_out/r2b_newton_trips.txt
