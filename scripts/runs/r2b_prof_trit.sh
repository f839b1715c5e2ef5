mkdir -p gpurun_out
HC_B200_JIT=1 timeout 1200 ncu --set full --import-source on --clock-control none -k regex:hc_jit_track -c 1 -f -o gpurun_out/r2b_full_trit16k python tests/tools/gpu_run_once.py tritangents 16384 1 > gpurun_out/r2b_full_trit16k.out 2>&1
tail -3 gpurun_out/r2b_full_trit16k.out
