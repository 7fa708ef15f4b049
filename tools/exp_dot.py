#!/usr/bin/env python
"""Experiment: the multi-dot kernels at the bench size (k vectors of 4 x 1.73M doubles), old vs fused"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from svfsi_b200 import api, mesh
api.init(device=0, rank=0, nranks=1)
gnNo, p = bench.setup_rank(api, mesh, (64, 64, int(os.environ.get("NZ", "408"))), 0, 1)
n = p.rm.nNo * 4
for k in (1, 2, 8, 9, 16, 17, 34, 50):
    r = dict(k=k)
    for v, name in ((0, "old"), (1, "fused")):
        api.time_kernel(3, 4, k, 3, v)
        ms = api.time_kernel(3, 4, k, 20, v) / 20
        r[name + "_us"] = ms * 1e3
        r[name + "_GBps"] = 8.0 * n * (k + 1) / (ms * 1e-3) / 1e9
    print(json.dumps(r), flush=True)
api.finalize()
