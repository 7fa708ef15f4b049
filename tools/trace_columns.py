#!/usr/bin/env python
"""Reads the per-rank column traces written with SVFSI_TRACE_FILE=<path> (device globaltimer stamps of the SpMV
+ fused column kernels, svfsi_b200/csrc/core.cu trace_slot) and prints, per rank, the mean split of a
Gram-Schmidt column: where the time between two columns goes (GPU clocks are not compared across ranks).
    python tools/trace_columns.py <path> <nranks> [first_col]"""
import sys

import numpy as np

path, n = sys.argv[1], int(sys.argv[2])
first = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rows = []
for r in range(n):
    a = np.loadtxt(f"{path}.{r}", dtype=np.uint64, ndmin=2).astype(np.int64)
    a = a[first:]
    ok = (a[:, 0] > 0) & (a[:, 2] > 0) & (a[:, 4] > 0)
    a = a[ok]
    t0, t1, t2, t3, t4, t5, t6, t7 = (a[:, i].astype(np.float64) for i in range(8))
    per = np.diff(t0)
    good = per < 5e5          # drop the gaps between solves
    d = dict(rank=r, cols=int(a.shape[0]),
             period_us=float(np.median(per[good]) / 1e3) if good.any() else 0.0,
             spmv_to_dot_us=float(np.median((t0 - t5)[t5 > 0]) / 1e3) if (t5 > 0).any() else 0.0,
             halo_pub_after_spmv_start_us=float(np.median((t6 - t5)[(t6 > 0) & (t5 > 0)]) / 1e3) if (t6 > 0).any() else 0.0,
             recv_wait_us=float(np.median((t1 - t0)[t1 > 0]) / 1e3) if (t1 > 0).any() else 0.0,
             dot_compute_us=float(np.median(t2 - np.where(t1 > 0, t1, t0)) / 1e3),
             own_flag_us=float(np.median((t7 - t2)[t7 > 0]) / 1e3) if (t7 > 0).any() else 0.0,
             allreduce_wait_us=float(np.median((t3 - t7)[(t3 > 0) & (t7 > 0)]) / 1e3) if (t3 > 0).any() else 0.0,
             tail_us=float(np.median(t4 - np.where(t3 > 0, t3, t2)) / 1e3),
             dot_kernel_us=float(np.median(t4 - t0) / 1e3))
    rows.append(d)
    print(" ".join(f"{k}={v:.1f}" if isinstance(v, float) else f"{k}={v}" for k, v in d.items()))
