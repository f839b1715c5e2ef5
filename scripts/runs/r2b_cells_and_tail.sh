# device-made polyhedral starts (GPU tests) + how fast the lane-group engine walks the tritangents tail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "polyhedral" 2>&1 | tail -5
for cfg in "HC_B200_JIT=1" "HC_B200_ENGINE=group HC_B200_GROUP=8" "HC_B200_ENGINE=group HC_B200_GROUP=32" "HC_B200_JIT=0"; do
  echo "== $cfg (tritangents, first 16384 paths)"
  env $cfg timeout 600 python tests/tools/gpu_run_once.py tritangents 16384 1 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r2b_tail_engines.txt
