# two-pass batches: thread-per-path specialised kernel + lane-group second pass for the paths it hands over
mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-2}" timeout 600 python tests/tools/gpu_run_once.py "${@: -2}" 1 2>&1 | tail -1; }
{
run HC_B200_HANDOFF=0 tritangents 1
for st in 150 200 300 500; do for g in 8 32; do run HC_B200_HANDOFF_STEPS=$st HC_B200_HANDOFF_GROUP=$g tritangents 1; done; done
run HC_B200_HANDOFF_STEPS=300 HC_B200_HANDOFF_EXT=0 tritangents 1
run HC_B200_JIT=0 cyclooctane_td 1
run HC_B200_HANDOFF=0 cyclooctane_td 1
for st in 250 350 500; do for g in 8 32; do run HC_B200_HANDOFF_STEPS=$st HC_B200_HANDOFF_GROUP=$g cyclooctane_td 1; done; done
run HC_B200_JIT=0 cyclooctane_polyhedral 1
run HC_B200_JIT=1 cyclooctane_polyhedral 1
run HC_B200_HANDOFF=0 katsura8 592
run HC_B200_HANDOFF=1 katsura8 592
run HC_B200_HANDOFF=0 cyclic7_polyhedral 160
run HC_B200_HANDOFF=1 cyclic7_polyhedral 160
} 2>&1 | tee gpurun_out/r2b_handoff1.txt
