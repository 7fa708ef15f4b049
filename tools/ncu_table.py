#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> a markdown table of the metrics the profiles/ summaries quote.
Usage: ncu_table.py raw.csv [kernel-substring]"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 wavefronts %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp instr"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard / issue"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
sub = sys.argv[2] if len(sys.argv) > 2 else ""
cols = [(k, n) for k, n in COLS if k in ix]
print("| kernel | " + " | ".join(f"{n} ({units[ix[k]]})" if units[ix[k]] else n for k, n in cols) + " |")
print("|---|" + "---:|" * len(cols))
for r in rows[2:]:
    if len(r) < len(hdr) or sub not in r[ix["Kernel Name"]]:
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("svfsi::", "").replace("void ", "").strip()

    def fmt(v):
        try:
            f = float(v.replace(",", ""))
            return f"{f:.4g}" if abs(f) < 1e6 else f"{f:.4g}"
        except ValueError:
            return v
    print(f"| `{name}` | " + " | ".join(fmt(r[ix[k]]) for k, _ in cols) + " |")
