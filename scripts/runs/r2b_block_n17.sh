# lanes per SM of the specialised kernel at n = 17 (25 KB of lane state each); full capture of the second-pass kernel
mkdir -p gpurun_out
run() { echo "== $*"; env "${@:1:$#-3}" timeout 600 python tests/tools/gpu_run_once.py "${@: -3}" 2>&1 | tail -${@: -1}; }
{
for b in 256 192 128; do run HC_B200_JIT_BLOCK=$b cyclooctane_td 1 1; run HC_B200_JIT_BLOCK=$b cyclooctane_polyhedral 1 1; done
} 2>&1 | tee gpurun_out/r2b_block_n17.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hc_track_kernel -c 1 -f -o gpurun_out/r02b_full_pass2_cyclooctane python tests/tools/gpu_run_once.py cyclooctane_td 1 1 > gpurun_out/r02b_full_pass2_cyclooctane.out 2>&1
tail -2 gpurun_out/r02b_full_pass2_cyclooctane.out
