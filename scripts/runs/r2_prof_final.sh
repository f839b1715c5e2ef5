# round-2 evidence: launch list of the bench command, one full capture of the dominant kernel, DRAM traffic per workload
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu-baseline > gpurun_out/r02_launches_bench.out 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hc_jit_track -c 1 -f -o gpurun_out/r02_full_cyclic7 python tests/tools/gpu_run_once.py cyclic7_polyhedral 160 1 > gpurun_out/r02_full_cyclic7.out 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for w in "katsura8 592" "tritangents 1" "biochem_sweep 256" "cyclooctane_td 1"; do
  set -- $w
  timeout 900 ncu --metrics $M --clock-control none -k regex:hc_ -c 1 --csv --log-file gpurun_out/r02_traffic_$1.csv python tests/tools/gpu_run_once.py $1 $2 1 > gpurun_out/r02_traffic_$1.out 2>&1
done
