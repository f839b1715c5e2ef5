# two-pass batches with the endgame-step criterion; GPU tests of the two-pass path; katsura / cyclic-7 with and without
mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -${@: -1}; }
{
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_pass" 2>&1 | tail -5
for e in 80 120 160; do run HC_B200_HANDOFF_EG_STEPS=$e tritangents 1 1; done
run HC_B200_HANDOFF_EG_STEPS=120 HC_B200_HANDOFF_GROUP=32 tritangents 1 1
for e in 80 120 160; do run HC_B200_HANDOFF_EG_STEPS=$e cyclooctane_td 1 1; done
run HC_B200_HANDOFF_EG_STEPS=120 HC_B200_HANDOFF_GROUP=8 cyclooctane_td 1 1
for e in 80 120 160; do run HC_B200_HANDOFF_EG_STEPS=$e cyclooctane_polyhedral 1 1; done
run HC_B200_HANDOFF=0 katsura8 592 3
run HC_B200_HANDOFF=1 katsura8 592 3
run HC_B200_HANDOFF=1 cyclic7_polyhedral 160 3
run HC_B200_HANDOFF=1 biochem_sweep 512 3
} 2>&1 | tee gpurun_out/r2b_handoff2.txt
