#!/usr/bin/env python
"""Print the metrics that matter from an .ncu-rep (raw page): time, DRAM traffic, pipe utilisation, stalls."""
import csv, subprocess, sys
csv.field_size_limit(10**9)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsuin_requests.avg.pct_of_peak_sustained_active", "l1tex__f_wavefronts.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
filt = sys.argv[2] if len(sys.argv) > 2 else None
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        if h in KEYS or "issue_stalled" in h and "per_issue_active" in h and "TriageCompute" not in h or (filt and filt in h):
            print("  %-90s %s %s" % (h, v, u))
