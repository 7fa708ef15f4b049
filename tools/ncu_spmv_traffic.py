#!/usr/bin/env python
"""Turns an `ncu --set full ... --page raw --csv` dump into profiles/r01_spmv_traffic.json: DRAM bytes read
+ written per launch of the SPARMULVV dof=4 kernel, which bench.py reports as roofline.traffic.
Usage: ncu_spmv_traffic.py raw.csv nnz nNo out.json [source note]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def val(r, key, table):
    return float(r[ix[key]].replace(",", "")) * table[units[ix[key]]]


sel = [r for r in rows[2:] if len(r) > 5 and "spmv_vv4" in r[ix["Kernel Name"]]]
if not sel:
    sys.exit("no spmv_vv4 launch in the capture")
r = sel[0]
name = r[ix["Kernel Name"]].split("(")[0].replace("svfsi::", "").strip()
out = dict(kernel=name, nnz=int(sys.argv[2]), nNo=int(sys.argv[3]),
           dram_bytes_read=int(val(r, "dram__bytes_read.sum", scale)),
           dram_bytes_write=int(val(r, "dram__bytes_write.sum", scale)),
           gpu_time_us=val(r, "gpu__time_duration.sum", tscale), launches_in_capture=len(sel),
           source=(sys.argv[5] if len(sys.argv) > 5 else "ncu --set full --clock-control none, first captured launch"))
json.dump(out, open(sys.argv[4], "w"), indent=1)
print(json.dumps(out))
