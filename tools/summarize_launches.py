#!/usr/bin/env python
"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> per-kernel table (markdown).
Usage: summarize_launches.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as fh:
    lines = [ln for ln in fh if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
tot = defaultdict(float); cnt = defaultdict(int)
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).strip()
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {tot[k] / cnt[k]:.1f} | {100 * tot[k] / T:.1f}% |")
print(f"\ntotal {T / 1e3:.2f} ms over {sum(cnt.values())} launches")
