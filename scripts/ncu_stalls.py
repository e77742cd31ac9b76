"""Print the warp-stall breakdown and a few headline metrics of every kernel in an ncu raw CSV."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:70])
    for h in hdr:
        if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h:
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v > 0.05:
                print(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
    for h in ["gpu__time_duration.sum", "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
              "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_shared_mem",
              "launch__occupancy_limit_registers", "launch__registers_per_thread", "dram__bytes_write.sum", "dram__bytes_read.sum",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]:
        if h in idx:
            print(f"   {h:70s} {r[idx[h]]} {units[idx[h]]}")
