#!/usr/bin/env python
"""How much would staging element records once per CTA save in the gather assembly?  (CPU analysis for DESIGN.md
section 10, item 1 -- no GPU needed.)

The default gather kernel reads ~5 sectors of a 640-byte element record per 4x4 contribution, 16 contributions per
element, from L2 (profiles/r02_asm_experiments.md).  If a CTA owns a CLUSTER of rows (nodes) and stages the records
of every element that touches the cluster in shared memory once, the L2 -> SM traffic becomes
    (elements touching the cluster) x record bytes            per cluster,
and a staged element serves  reuse = (its nodes inside the cluster)  of its 4 nodes.  Total traffic relative to
reading every record exactly once = 4 / mean reuse.  This script measures mean reuse for
  * chunks of consecutive rows in the mesh's own numbering (what a CTA gets without any reordering), and
  * clusters grown breadth-first over the node graph (a greedy spatial clustering)
on the bench's Kuhn-lattice pipe and on the irregular test meshes, for cluster sizes that fit in shared memory
(a record is 440 - 640 B: 200 KB hold ~320 - 450 records).
    python tools/cluster_reuse.py > profiles/r02_cluster_reuse.md"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from svfsi_b200 import mesh  # noqa: E402


def reuse_of(labels, IEN):
    """labels[node] = cluster id.  Returns (mean nodes-in-cluster per staged (element, cluster) pair,
    max staged elements per cluster, mean staged elements per cluster)"""
    lab = labels[IEN]                                   # (nEl, 4)
    lab_sorted = np.sort(lab, axis=1)
    distinct = 1 + (np.diff(lab_sorted, axis=1) != 0).sum(axis=1)      # clusters an element is staged in
    staged = distinct.sum()
    nclu = labels.max() + 1
    # staged elements per cluster
    first = np.ones_like(lab_sorted, dtype=bool)
    first[:, 1:] = np.diff(lab_sorted, axis=1) != 0
    per = np.bincount(lab_sorted[first], minlength=nclu)
    return 4.0 * IEN.shape[0] / staged, int(per.max()), float(per.mean())


def chunks(nNo, size):
    return (np.arange(nNo) // size).astype(np.int64)


def bfs_clusters(nNo, IEN, size):
    """greedy: grow a cluster breadth-first from the lowest unassigned node until it holds `size` nodes"""
    import scipy.sparse as sp
    r = np.repeat(IEN, 4, axis=1).ravel()
    c = np.tile(IEN, (1, 4)).ravel()
    A = sp.csr_matrix((np.ones(r.size, dtype=np.int8), (r, c)), shape=(nNo, nNo))
    indptr, indices = A.indptr, A.indices
    lab = -np.ones(nNo, dtype=np.int64)
    cur = 0
    nxt = 0
    from collections import deque
    while True:
        while nxt < nNo and lab[nxt] >= 0:
            nxt += 1
        if nxt >= nNo:
            break
        q = deque([nxt])
        lab[nxt] = cur
        n = 1
        while q and n < size:
            u = q.popleft()
            for v in indices[indptr[u]:indptr[u + 1]]:
                if lab[v] < 0:
                    lab[v] = cur
                    n += 1
                    q.append(v)
                    if n >= size:
                        break
        cur += 1
    return lab


def main():
    import unstructured as un
    cases = []
    m = mesh.make_cylinder(32, 32, 48, R=2.0, L=7.0)
    cases.append(("Kuhn-lattice pipe 32 x 32 x 48 (295k tets; the bench mesh is the same lattice, 64 x 64 x 408)",
                  m.nNo, m.IEN.astype(np.int64) - 1))
    x, IEN = un.delaunay_box(n=6000, seed=3)
    cases.append((f"Delaunay box, {IEN.shape[0]} tets, elements shuffled", x.shape[0], IEN.astype(np.int64) - 1))
    print("# Round 2 - record reuse of a cluster-owner assembly (CPU analysis, `tools/cluster_reuse.py`)\n")
    print(__doc__.split("\n\n")[1] + "\n")
    for name, nNo, ien in cases:
        print(f"\n## {name}\n")
        print("| rows per CTA | clustering | nodes served per staged record (of 4) | L2 -> SM record traffic vs reading every record once | staged records per CTA (mean / max) |")
        print("|---:|---|---:|---:|---:|")
        for size in (32, 64, 128, 256):
            for cname, lab in (("consecutive rows", chunks(nNo, size)), ("breadth-first clusters", bfs_clusters(nNo, ien, size))):
                reuse, mx, mean = reuse_of(lab, ien)
                print(f"| {size} | {cname} | {reuse:.2f} | {4.0 / reuse:.2f} x | {mean:.0f} / {mx} |")
    print("""
Reading: the default kernel moves ~25 GB from L2 to the SMs per assembly of the 10M-tet mesh (~2.5 KB per element, i.e. ~4 x the
640-byte record).  A cluster-owner kernel moves `record bytes x (4 / reuse)`.  The records need not all sit in shared memory at once
(they can stream through it in tiles, each consumed once by the CTA); what must stay resident is the cluster's OUTPUT: rows x ~15 blocks
x 128 B = 61 / 123 / 245 KB for 32 / 64 / 128 rows, so 64-row clusters are the largest that fit next to a record tile in 227 KB.
With 64-row breadth-first clusters a staged record serves 2.1 - 2.3 of its 4 nodes: 1.8 - 1.9 x the records = ~12 GB at 640 B per record
(~8 GB with the 440-byte minimal record) instead of ~25 GB, all of it as coalesced bulk copies instead of 32-byte gathers.
Consecutive rows of the mesh's own numbering are lines / planes (lattice) or unrelated nodes (shuffled Delaunay mesh), not compact blocks:
they reuse a record 1.0 - 1.7 times and would not pay -- the cluster builder is the enabling piece.""")


if __name__ == "__main__":
    main()
