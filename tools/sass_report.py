#!/usr/bin/env python
"""profiles/r02_sass_ptxas.md: the ptxas register / spill table of every kernel of libsvfsi_b200.so (from the
`-Xptxas -v` logs the Makefile keeps next to the objects) and the memory / FP64 instruction mix of the hot kernels
from `cuobjdump -sass` (proof of sm_100a code with 256-bit loads, FP64 FMAs, no local memory).
    python tools/sass_report.py > profiles/r02_sass_ptxas.md"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "svfsi_b200", "csrc")
HOT = ["spmv_vv4_quad_kernel", "spmv_vv4_fused_kernel", "spmv_vv4_quad_scale_kernel", "fluid_record6_kernel",
       "fluid_gather_quad_kernel", "fluid_gather_r_kernel", "multidot_fused_kernel", "multi_axpy_scale_kernel",
       "spmv_small_kernelILi15ELi3ELi3ELi8ELi2", "spmv_small_kernelILi0ELi1ELi1ELi4ELi4", "spmv_small_kernelILi15ELi3ELi1ELi8ELi2",
       "spmv_small_kernelILi0ELi1ELi3ELi4ELi4"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except OSError:
        return n


print("# Round 2 - ptxas resources and SASS instruction mix (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`)\n")
print("`tools/sass_report.py`; register / spill numbers from `-Xptxas -v`, instruction mix from `cuobjdump -sass` of the objects the library links.\n")
print("## Registers, spills, shared memory of every kernel\n\n| file | kernel | registers | spill stores (B) | static smem (B) |\n|---|---|---:|---:|---:|")
for log in sorted(glob.glob(os.path.join(CS, "*.ptxas.log"))):
    L = open(log).read().split("\n")
    for i, l in enumerate(L):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", l)
        if not m:
            continue
        info = " ".join(L[i + 1:i + 5])
        regs = re.search(r"Used (\d+) registers", info)
        sp = re.search(r"(\d+) bytes spill stores", info)
        sm = re.search(r"(\d+) bytes smem", info)
        name = demangle(m.group(1)).replace("svfsi::", "").replace("void ", "")
        print(f"| {os.path.basename(log)[:-10]} | `{name}` | {regs.group(1) if regs else '?'} | {sp.group(1) if sp else 0} | {sm.group(1) if sm else 0} |")
print("\n## Instruction mix of the hot kernels (static SASS counts)\n")
print("| kernel | arch | LDG.256 | LDG.128 | LDG.64 | LDG.32 | STG | DFMA | DMUL | DADD | SHFL | LDS/STS | LDL/STL | MUFU.RSQ64H | total |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for obj in sorted(glob.glob(os.path.join(CS, "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    arch = re.search(r"arch = (\S+)", out)
    cur, cnt = None, None
    fns = {}
    for l in out.split("\n"):
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1)
            cnt = fns.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m and cur:
            op = m.group(1)
            cnt["total"] += 1
            if op.startswith("LDG"):
                cnt["LDG.256" if ".256" in op else "LDG.128" if ".128" in op else "LDG.64" if ".64" in op else "LDG.32"] += 1
            elif op.startswith("STG"):
                cnt["STG"] += 1
            elif op.startswith(("DFMA", "DMUL", "DADD", "SHFL")):
                cnt[op.split(".")[0]] += 1
            elif op.startswith(("LDS", "STS")):
                cnt["LDS/STS"] += 1
            elif op.startswith(("LDL", "STL")):
                cnt["LDL/STL"] += 1
            elif op.startswith("MUFU.RSQ64H"):
                cnt["MUFU.RSQ64H"] += 1
    for fn, c in fns.items():
        if any(h in fn for h in HOT):
            name = demangle(fn).replace("svfsi::", "").replace("void ", "")
            cols = ["LDG.256", "LDG.128", "LDG.64", "LDG.32", "STG", "DFMA", "DMUL", "DADD", "SHFL", "LDS/STS", "LDL/STL", "MUFU.RSQ64H", "total"]
            print(f"| `{name}` | {arch.group(1) if arch else '?'} | " + " | ".join(str(c[k]) for k in cols) + " |")
print("\nFirst 256-bit loads of `spmv_vv4_quad_kernel` (the dominant kernel), verbatim from `cuobjdump -sass`:\n\n```")
out = subprocess.run(["cuobjdump", "-sass", os.path.join(CS, "la_kernels.o")], capture_output=True, text=True).stdout
on = False
n = 0
for l in out.split("\n"):
    if "Function :" in l:
        on = "spmv_vv4_quad_kernel" in l
    if on and (".256" in l or "DFMA" in l) and "/*" in l:
        print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip())
        n += 1
        if n >= 14:
            break
print("```")
