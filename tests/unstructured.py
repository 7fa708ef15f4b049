"""Irregular tetrahedral meshes for the parity tests (test infrastructure): every other test uses the
Kuhn lattice, whose rows all have <= 15 blocks and whose nodes have <= 24 elements.  These meshes have
what a real svFSI mesh has -- arbitrary valence, arbitrary element numbering -- and one deliberately
extreme case (a node shared by > 64 elements / a row of > 64 blocks), which is where the row-owner
assembly kernel hands over to the block-owner one (asm_kernels.cu launch_fluid_gather_parts)."""
import numpy as np

from svfsi_b200 import mesh


def _orient(x, tets):
    """svFSI reorders the nodes of an element so that the Jacobian is positive (S/READMSH.f:1010-1185 CHECKIEN);
    do the same: swap the first two nodes where det < 0, drop slivers."""
    t = tets.astype(np.int64)
    X = x[t[:, :3]] - x[t[:, 3:4]]
    det = np.linalg.det(X)
    vol_scale = np.abs(det).mean()
    keep = np.abs(det) > 1e-3 * vol_scale
    t, det = t[keep], det[keep]
    neg = det < 0
    t[neg, 0], t[neg, 1] = t[neg, 1].copy(), t[neg, 0].copy()
    return t


def _compact(x, tets):
    used = np.unique(tets)
    new = np.full(x.shape[0], -1, dtype=np.int64)
    new[used] = np.arange(used.size)
    return np.ascontiguousarray(x[used]), new[tets]


def delaunay_box(n=350, seed=11):
    """Delaunay tetrahedralisation of random points in a 1 x 1 x 2 box, elements shuffled."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 1.0, (n, 3)) * np.array([1.0, 1.0, 2.0])
    t = _orient(x, Delaunay(x).simplices)
    x, t = _compact(x, t)
    t = t[rng.permutation(t.shape[0])]
    return x, (t + 1).astype(np.int32)


def fan(nsurf=90, seed=5):
    """A ball: one centre node joined to every triangle of the convex hull of `nsurf` points on the
    unit sphere -> the centre node belongs to ~2*nsurf elements and its row has nsurf + 1 blocks."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(seed)
    p = rng.standard_normal((nsurf, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    tri = ConvexHull(p).simplices
    x = np.vstack([p, np.zeros((1, 3))])
    t = np.hstack([tri, np.full((tri.shape[0], 1), nsurf)])
    t = _orient(x, t)
    # the centre is NOT the last row of the pattern: renumber it into the middle
    perm = rng.permutation(x.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return np.ascontiguousarray(x[perm]), (inv[t] + 1).astype(np.int32)


def problem(x, IEN, seed=3):
    """(rowPtr, colPtr, Ag, Yg) for a one-rank run on the mesh: LHSA pattern + a smooth velocity /
    pressure state with 1 % noise and a random acceleration."""
    nNo = x.shape[0]
    rowPtr, colPtr = mesh.csr_pattern(nNo, IEN)
    rng = np.random.default_rng(seed)
    Yg = np.zeros((nNo, 4))
    Yg[:, 0] = 3.0 * x[:, 1] * x[:, 2]
    Yg[:, 1] = -2.0 * x[:, 0] + x[:, 2] ** 2
    Yg[:, 2] = 5.0 + x[:, 0] * x[:, 1]
    Yg[:, 3] = 10.0 - 4.0 * x[:, 2]
    Yg *= 1.0 + 0.01 * rng.standard_normal(Yg.shape)
    Ag = rng.standard_normal((nNo, 4))
    Ag[:, 3] = 0.0
    return rowPtr, colPtr, Ag, Yg
