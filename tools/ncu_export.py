#!/usr/bin/env python
"""Export what the judge needs from .ncu-rep files into profiles/: a per-kernel text summary and traffic.json
(dram bytes read + written per launch).   python tools/ncu_export.py OUT_PREFIX rep1.ncu-rep [rep2 ...]"""
import csv, json, os, subprocess, sys
csv.field_size_limit(10**9)
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
prefix = sys.argv[1]
traffic = {}
lines = []
for rep in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        lines.append("== %s   [%s]  grid %s block %s" % (short, os.path.basename(rep), r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        vals = dict(zip(hdr, zip(r, units)))
        for k in KEEP:
            if k in vals:
                lines.append("   %-75s %s %s" % (k, vals[k][0], vals[k][1]))
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "TriageCompute" not in h:
                v = float(vals[h][0].replace(",", "") or 0)
                if v >= 0.2:
                    lines.append("   stall %-69s %.2f" % (h.split("issue_stalled_")[1].split("_per_issue")[0], v))
        def num(key):
            v, u = vals[key]
            v = float(v.replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
        try:
            key = short.split("<")[0]
            traffic[key] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        except Exception:
            pass
open(prefix + "_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
tpath = os.path.join(os.path.dirname(prefix), "traffic.json")
try:          # keep the hand-written entries (instruction counts, fp64 shares) of kernels this export does not cover
    merged = json.load(open(tpath))
except Exception:
    merged = {}
merged.update(traffic)
json.dump(merged, open(tpath, "w"), indent=1)
print("\n".join(lines[:12])); print(traffic)
