# lane-strided loops with a warp-uniform trip count (guarded body): second pass and the regular configs again
mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -${@: -1}; }
{
run X=1 tritangents 1 1
run X=1 cyclooctane_td 1 1
run X=1 cyclooctane_polyhedral 1 1
run X=1 katsura8 592 2
run X=1 cyclic7_polyhedral 160 2
run X=1 biochem_sweep 512 2
timeout 1500 python -m pytest tests/test_jit.py tests/test_gpu_parity.py -q -m gpu -k "bit_identical or two_pass or group_engine or system_sizes" 2>&1 | tail -4
} 2>&1 | tee gpurun_out/r2b_uniform_loops.txt
