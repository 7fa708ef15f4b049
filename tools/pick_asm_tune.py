#!/usr/bin/env python
# NOTE: the deterministic-repeatability part of the variant test (second run bitwise equal) holds for
# every variant; variants differ from each other in the last bits (different FMA contraction).
"""Picks the fastest records-kernel and gather-kernel variants out of tools/time_asm_variants.py's
JSON, restricted to the variants whose parity test passed (one tune per line in the ok file).
Prints the combined SVFSI_ASM_TUNE value.  Usage: pick_asm_tune.py variants.json ok.txt"""
import json
import sys

j = json.load(open(sys.argv[1]))
try:
    ok = {int(x) for x in open(sys.argv[2]).read().split()}
except OSError:
    ok = set()
rec = {int(k[len("record_tune"):-3]): v for k, v in j.items() if k.startswith("record_tune")}
gat = {int(k[len("gather_val_tune"):-3]): v for k, v in j.items() if k.startswith("gather_val_tune")}
# a records variant r is trusted if some tested combination containing it passed; same for gathers
REC_BITS = 1 | 128
rec_ok = {r: v for r, v in rec.items() if any((t & REC_BITS) == r for t in ok)} or {0: rec.get(0, 0.0)}
gat_ok = {g: v for g, v in gat.items() if any((t & ~REC_BITS) == g for t in ok)} or {8: gat.get(8, 0.0)}
# the row-owner (32) and pair-owner (1024) kernels also do the residual gather; the block-owner
# ones need kernel C on top
c_ms = j.get("gather_r_tune0_ms", 0.0)
gat_ok = {g: v + (0.0 if g & (32 | 1024) else c_ms) for g, v in gat_ok.items()}
r = min(rec_ok, key=rec_ok.get)
g = min(gat_ok, key=gat_ok.get)
print(r | g)
