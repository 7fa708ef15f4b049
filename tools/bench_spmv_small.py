#!/usr/bin/env python
"""Small-block SpMV shapes (NSSOLVER: K 3x3, G 3x1, D 1x3, L 1x1; heat CG: 1x1; L/SPARMUL.f:135-297): every kernel
family (gpu_set_spmv_small_ / SVFSI_SPMV_SMALL modes 0..6) timed alone on the bench pipe's pattern, after a
result check of each family against the lane-per-block kernel (mode 0, oracle-checked by the parity tests) on a
1M-tet pattern.  GPU box only:  python tools/bench_spmv_small.py [nz] [--ncu] > gpurun_out/spmv_small.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from svfsi_b200 import api, mesh  # noqa: E402

SHAPES = (("SS", 3, 1, 1, 1), ("VV", 0, 3, 3, 3), ("SV", 2, 3, 3, 1), ("VS", 1, 3, 1, 3))   # name, kind, dof, BR, BC
MODES = tuple(int(m) for m in os.environ.get("SMALL_MODES", "0,1,2,3,4,5,6").split(","))
nz = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 408
peak = bench.measured_peak()[0]
out = dict(peak_GBps=peak, modes=list(MODES))
api.init(device=0, rank=0, nranks=1)


def lhs(nzz):
    """FSILS_LHS_CREATE on the pipe's pattern (cached under /tmp: the --ncu pass of the same box reuses it)"""
    cache = f"/tmp/spmv_small_pattern_{nzz}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        gnNo, ltg, rowPtr, colPtr = int(z["gnNo"]), z["ltg"], z["rowPtr"], z["colPtr"]
    else:
        gnNo, p = mesh.build_rank_problem(64, 64, nzz, rank=0, nparts=1, R=bench.R_PIPE,
                                          L=bench.L_PIPE * nzz / bench.DIMS[2])
        ltg, rowPtr, colPtr = p.rm.ltg, p.rowPtr, p.colPtr
        np.savez(cache, gnNo=gnNo, ltg=ltg, rowPtr=rowPtr, colPtr=colPtr)
    api.FSILS_LHS_CREATE(gnNo, ltg.size, colPtr.size, ltg, rowPtr, colPtr, 0)
    return int(colPtr.size), int(ltg.size)


NCU = "--ncu" in sys.argv
if NCU:      # one pair of launches per shape and family for `ncu -k regex:spmv_` (no result check, no timing table)
    nnz, nNo = lhs(nz)
    for name, kind, dof, br, bc in SHAPES:
        for md in MODES:
            api.time_kernel(6, dof, kind, 0, md)     # the warm-up launch only
    api.finalize()
    sys.exit(0)

# ---- results: every family against the lane-per-block kernel, 40 axial cells (1M tets)
nnz, nNo = lhs(40)
rng = np.random.default_rng(3)
errs = {}
for name, kind, dof, br, bc in SHAPES:
    K = rng.standard_normal((nnz, br * bc)); U = rng.standard_normal((nNo, bc))
    api.set_spmv_small(0)
    ref = api.FSILS_SPARMUL(name, dof, K, U)
    for md in MODES:
        if md == 0:
            continue
        api.set_spmv_small(md)
        got = api.FSILS_SPARMUL(name, dof, K, U)
        errs[f"{name}{dof}_mode{md}"] = float(np.abs(got - ref).max() / np.abs(ref).max())
api.set_spmv_small(-1)
api.FSILS_LHS_FREE()
out["check_1M_max_rel_diff_vs_mode0"] = errs

# ---- timing on the bench pattern
nnz, nNo = lhs(nz)
out.update(nz=nz, nnz=nnz, nNo=nNo)
tab = {}
for name, kind, dof, br, bc in SHAPES:
    nbytes = nnz * (8.0 * br * bc + 4.0) + nNo * (8.0 + 8.0 * br + 8.0 * bc)
    for md in MODES:
        api.time_kernel(6, dof, kind, 3, md)
        ms = api.time_kernel(6, dof, kind, 20, md) / 20
        tab[f"{name}{dof}_mode{md}"] = dict(ms=ms, GBps=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / peak)
out["timing"] = tab
print(json.dumps(out, indent=1))
api.finalize()
