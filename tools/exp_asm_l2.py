#!/usr/bin/env python
"""Experiment: per-element time of the three gather-assembly kernels as a function of the mesh size, i.e. of
whether the 640-byte element records stay resident in the 126 MB L2 between the records kernel and the
gather kernels (64 x 64 x nz lattice: nz = 4 -> 63 MB of records, 8 -> 126 MB, 408 -> 6.4 GB)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from svfsi_b200 import api, mesh  # noqa: E402

api.init(device=0, rank=0, nranks=1)
tune = int(os.environ.get("SVFSI_ASM_TUNE", "42733696"))
res = []
for nz in [int(a) for a in (sys.argv[1:] or ["2", "4", "8", "16", "64"])]:
    gnNo, p = bench.setup_rank(api, mesh, (64, 64, nz), 0, 1)
    api.state_upload(4, p.Ag, p.Yg, None)
    api.construct_fluid_dev(bench.RHO, bench.MU, bench.F_BODY, bench.DT, bench.GA["af"], bench.GA["am"],
                            bench.GA["gam"], api.ASM_GATHER)
    api.sync()
    r = dict(nz=nz, nEl=int(p.rm.nEl), rec_MB=p.rm.nEl * 640 / 1e6)
    for part, name in ((1, "A"), (2, "B"), (4, "C"), (7, "ABC")):
        api.time_kernel(5, 4, part, 3, tune)
        ms = api.time_kernel(5, 4, part, 20, tune) / 20
        r[name + "_ms"] = ms
        r[name + "_ns_per_elem"] = ms * 1e6 / p.rm.nEl
    res.append(r)
    print(json.dumps(r), flush=True)
    api.FSILS_LHS_FREE()
api.finalize()
