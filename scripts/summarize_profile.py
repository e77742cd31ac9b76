"""ncu raw CSV -> compact per-kernel summary CSV (the metrics DESIGN.md / bench.py quote)."""
import csv, sys
KEEP = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__issue_active.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h]
cols = [k for k in KEEP if k in idx] + stall
w = csv.writer(open(sys.argv[2], 'w', newline=''))
w.writerow(cols)
w.writerow([units[idx[c]] for c in cols])
for r in rows[2:]:
    w.writerow([r[idx[c]] for c in cols])
print(f"{len(rows) - 2} kernels -> {sys.argv[2]}")
