"""Markdown summary of an `ncu --set full` report: python profiles/ncu_summary.py <report.ncu-rep> > profiles/<name>.md
(runs `ncu -i ... --page raw --csv` and keeps the metrics the roofline discussion uses)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = [("time", "gpu__time_duration.sum"), ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
        ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 hit rate", "lts__t_sector_hit_rate.pct"),
        ("regs/thread", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("blocks/SM limit (regs)", "launch__occupancy_limit_registers"),
        ("blocks/SM limit (smem)", "launch__occupancy_limit_shared_mem"),
        ("waves per SM", "launch__waves_per_multiprocessor"),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("eligible warps / cycle / scheduler", "smsp__warps_eligible.avg.per_cycle_active"),
        ("warp instructions", "smsp__inst_executed.sum"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("xu pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("alu pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("fma pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("lsu pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("tensor pipe %", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
          or h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
names = [r[ix["Kernel Name"]].replace("void ", "").split("(")[0] for r in data]
print("ncu --set full --clock-control none, one launch per kernel (cold-cache, serialised; compare shares)\n")
print("| metric | unit | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(names))
for label, key in want:
    if key in ix:
        print("| %s | %s | %s |" % (label, units[ix[key]], " | ".join(r[ix[key]] for r in data)))
for key in sorted(stalls):
    short = key.replace("smsp__average_warps_issue_stalled_", "stall ").replace("smsp__average_warp_latency_issue_stalled_", "stall ").replace("_per_issue_active.ratio", "").replace(".ratio", "")
    vals = [r[ix[key]] for r in data]
    try:
        if max(float(v) for v in vals) < 0.15:
            continue
    except ValueError:
        pass
    print("| %s | %s | %s |" % (short, units[ix[key]], " | ".join(vals)))
