"""Prints the metrics we track from an .ncu-rep (first kernel): python tools/ncu_summary.py rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum", "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for r in rows[2:]:
    print("=" * 60)
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"{h:70s} {units[i]:12s} {r[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("stalls:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:7]))
