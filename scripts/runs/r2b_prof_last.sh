# evidence of the final build: launch lists (bench headline, a two-pass batch) and one full capture of the dominant kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu-baseline > gpurun_out/r02c_launches_bench.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 10 --csv --log-file gpurun_out/r02c_launches_tritangents.csv python tests/tools/gpu_run_once.py tritangents 1 1 > gpurun_out/r02c_launches_tritangents.out 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hc_jit_track -c 1 -f -o gpurun_out/r02c_full_cyclic7 python tests/tools/gpu_run_once.py cyclic7_polyhedral 160 1 > gpurun_out/r02c_full_cyclic7.out 2>&1
tail -3 gpurun_out/r02c_launches_tritangents.csv | cut -c1-200
