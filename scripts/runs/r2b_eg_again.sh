mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -1; }
{
for e in 60 90 120 200; do run HC_B200_HANDOFF_EG_STEPS=$e tritangents 1 1; done
for e in 90 200; do run HC_B200_HANDOFF_EG_STEPS=$e cyclooctane_td 1 1; done
} 2>&1 | tee gpurun_out/r2b_eg_again.txt
