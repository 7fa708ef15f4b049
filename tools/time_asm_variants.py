#!/usr/bin/env python
"""Times the three kernels of the gather assembly (records / tangent gather / residual gather)
for every kernel variant on the bench pipe (GPU box only):  python tools/time_asm_variants.py [nz]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from svfsi_b200 import api, mesh  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 408
api.init(device=0, rank=0, nranks=1)
gnNo, p = bench.setup_rank(api, mesh, (64, 64, nz), 0, 1)
api.state_upload(4, p.Ag, p.Yg, None)
bench.newton_step_dev(api, api.ASM_GATHER)
api.sync()
out = dict(nEl=int(p.rm.nEl), nnz=int(p.colPtr.size))
# tune bits: asm_kernels.cu asm_tune().  8 = block-owner gather (128-thread CTAs); 32 = row-owner gather
# (B + C in one launch), +64 / +512 = two / four visits in flight, +256 = 8 warps per CTA; 1024 = pair-owner
# gather (B + C in one launch), +2048 = 256-thread CTAs, +4096 / +8192 = 48 / 64-register cap;
# +32768 / +65536 = record prefetch into L1 / L2 (block-owner kernels 8 and 16384);
# 262144 = quad gather (4 lanes per block, 256-bit loads), +2048 = 256-thread CTAs, +8192 / +4096 = 48 / 64-register cap;
#   +524288 = block descriptors (one load at group start);
# 131072 = wide-load (256-bit) block-owner gather, +2048 = 256-thread CTAs, +8192 / +4096 = 48 / 64-register cap;
# 16384 = lean block-owner gather, +2048 = 256-thread CTAs, +8192 / +4096 = 48 / 32-register cap;
# records: 0 = v1, 128 = v3 (pair staging, rsqrt arithmetic)
tunes_val = tuple(int(t) for t in os.environ.get("ASM_TUNES", "8,266240,270336,786432,790528,794624").split(","))
tunes_rec = tuple(int(t) for t in os.environ.get("ASM_TUNES_REC", "0,128").split(","))
for part, name, tunes in ((1, "record", tunes_rec), (2, "gather_val", tunes_val), (4, "gather_r", (0,))):
    for tune in tunes:
        api.time_kernel(5, 4, part, 2, tune)
        out[f"{name}_tune{tune}_ms"] = api.time_kernel(5, 4, part, 10, tune) / 10
# SPARMULVV dof=4: 8 lanes per row (0) vs 4 lanes per row with 256-bit loads (1)
for v in (0, 1):
    api.time_kernel(0, 4, 0, 3, v)
    out[f"spmv_vv4_variant{v}_ms"] = api.time_kernel(0, 4, 0, 20, v) / 20
print(json.dumps(out))
api.finalize()
