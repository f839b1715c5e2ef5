# all paths of C3 / C4 (both start systems) on the final two-pass build vs the oracle on the box's cores
mkdir -p gpurun_out
for w in tritangents cyclooctane_td cyclooctane_polyhedral; do
  timeout 900 python tests/tools/gpu_full_config.py $w 2>&1 | grep -v Warning | tee gpurun_out/r02b_full_$w.txt
done
