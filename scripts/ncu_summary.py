"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) into the metrics DESIGN.md quotes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct', 'lts__t_bytes.sum', 'lts__throughput.avg.pct',
        'l1tex__throughput.avg.pct', 'sm__throughput.avg.pct', 'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'sm__inst_executed_pipe_fp64',
        'smsp__inst_executed.sum', 'pipe_fp64', 'smsp__issue_active.avg.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'op_dfma_pred_on.sum', 'op_dadd_pred_on.sum', 'op_dmul_pred_on.sum', 'launch__occupancy', 'thread_inst_executed_per_inst_executed',
        'issue_stalled', 'launch__grid_size', 'launch__block_size', 'local', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg ', 'smsp__inst_executed_op_', 'l1tex__data_bank']
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in keys):
            print(f"{h} [{u}] {v}")
