"""Turns an .ncu-rep (ncu --set full) into a small markdown table + traffic.json.
Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_ncu_full.md [traffic.json]"""
import csv
import io
import json
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
     ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
     ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
     ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
     ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
     ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected")]
names = [d[ix["Kernel Name"]].split("(")[0].replace("void ", "") for d in data]
with open(out_md, "w") as f:
    f.write("ncu --set full --clock-control none, one launch per kernel (cold-cache, serialised; compare shares)\n\n")
    f.write("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---|" * len(names) + "\n")
    for key, label in M:
        if key in ix:
            f.write("| %s | %s | %s |\n" % (label, units[ix[key]], " | ".join(d[ix[key]][:12] for d in data)))
print(open(out_md).read())
if len(sys.argv) > 3:
    traffic = {}
    for n, d in zip(names, data):
        def val(k):
            v = float(d[ix[k]])
            u = units[ix[k]].lower()
            return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
        b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        if "gauss_pass_x" in n:
            traffic["gauss_pass_x"] = b
        elif "features_kernel" in n or "features_march_kernel" in n:
            traffic["features_fused"] = b
        elif "gauss_pass_strided" in n:
            traffic["gauss_pass_y" if "gauss_pass_z" in traffic else "gauss_pass_z"] = b
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    print(traffic)
