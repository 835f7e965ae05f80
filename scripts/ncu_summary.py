"""Summarise an .ncu-rep (raw page) into a small markdown table under profiles/.  Usage: ncu_summary.py REP OUT.md"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst (warp execution efficiency x32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__occupancy_limit_registers", "occupancy limit regs (blocks)"), ("launch__occupancy_limit_shared_mem", "occupancy limit smem (blocks)"),
]
with open(out, "w") as f:
    f.write(f"# ncu --set full summary of `{rep}`\n\n")
    for r in rows[2:]:
        f.write(f"## {r[idx['Kernel Name']]} (launch id {r[idx['ID']]})\n\n| metric | value | unit |\n|---|---|---|\n")
        for k, label in want:
            if k in idx:
                f.write(f"| {label} (`{k}`) | {r[idx[k]]} | {units[idx[k]]} |\n")
        stalls = sorted(((float(r[idx[h]]) if r[idx[h]] not in ("", "n/a") else 0.0, h) for h in hdr
                         if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")), reverse=True)[:6]
        if stalls:
            f.write("\nTop warp stall reasons (cycles per issued instruction):\n\n")
            for v, h in stalls:
                f.write(f"* `{h}` = {v:.2f}\n")
        f.write("\n")
print("wrote", out)
