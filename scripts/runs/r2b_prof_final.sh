# round 2 (second session) evidence: full GPU test suite, launch list of the bench command and of a two-pass batch, DRAM traffic
# per workload (both kernels of a two-pass batch), one full capture of the dominant kernel on the final build
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2b_gpu_tests.txt 2>&1; tail -4 gpurun_out/r2b_gpu_tests.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu-baseline > gpurun_out/r02b_launches_bench.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 10 --csv --log-file gpurun_out/r02b_launches_tritangents.csv python tests/tools/gpu_run_once.py tritangents 1 1 > gpurun_out/r02b_launches_tritangents.out 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for w in "tritangents 1" "cyclooctane_polyhedral 1" "cyclooctane_td 1"; do
  set -- $w
  timeout 900 ncu --metrics $M --clock-control none -k regex:hc_ -c 2 --csv --log-file gpurun_out/r02b_metrics_$1.csv python tests/tools/gpu_run_once.py $1 $2 1 > gpurun_out/r02b_metrics_$1.out 2>&1
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hc_jit_track -c 1 -f -o gpurun_out/r02b_full_cyclic7 python tests/tools/gpu_run_once.py cyclic7_polyhedral 160 1 > gpurun_out/r02b_full_cyclic7.out 2>&1
ls -la gpurun_out | tail -12
