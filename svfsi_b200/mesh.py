"""Synthetic tet-cylinder problems for the svFSI fluid hot path (host-side input generator).

Produces exactly the arrays svFSI holds after READFILES/DISTRIBUTE/INITIALIZE for a one-mesh,
one-domain TET4 case (reference: Code/Source/svFSI):

* ``x(nsd,tnNo)``, ``IEN(eNoN,nEl)`` 1-based with svFSI's orientation (positive Jacobian, i.e.
  ``((x2-x1) x (x3-x2)).(x4-x3) < 0``, READMSH.f:1133-1149),
* faces inlet / outlet / wall as global node lists (+ boundary triangles), like ``msh%fa(:)%gN``,
* a k-way element partition with per-rank first-touch local numbering and ``ltg`` exactly as
  DISTRIBUTE.f:1455-1529 builds them (elements disjoint, cut nodes replicated, no ghosts),
* generalised-alpha constants (INITIALIZE.f:78-139) and a Poiseuille + perturbation state.

The lattice is Kuhn's 6-tet subdivision of a structured (nx, ny, nz) grid whose square cross
section is mapped radially onto a disc.  The subdivision is mirrored per quadrant so that the
cube diagonal always passes through the outer corner of the four corner columns (keeps the
corner tets well shaped); nx and ny must be even.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np

SEED = 1234  # seed for numpy default_rng


@dataclass
class Face:
    name: str
    gN: np.ndarray            # global node ids (1-based, int32, ascending)
    tri: np.ndarray           # (nTri, 3) global node ids (1-based) of boundary triangles
    parent: np.ndarray        # (nTri,) 0-based parent element id


@dataclass
class Mesh:
    x: np.ndarray             # (nNo, 3) float64
    IEN: np.ndarray           # (nEl, 4) int32, 1-based
    faces: dict = field(default_factory=dict)
    cell_k: np.ndarray | None = None   # (nEl,) axial cell index of every element (for slabs)
    dims: tuple = (0, 0, 0)
    R: float = 1.0
    L: float = 1.0

    @property
    def nNo(self):
        return self.x.shape[0]

    @property
    def nEl(self):
        return self.IEN.shape[0]


def _kuhn_paths():
    """Six monotone lattice paths (0,0,0)->(1,1,1); each is one tet of the Kuhn subdivision."""
    tets = []
    for perm in itertools.permutations(range(3)):
        v = np.zeros(3, dtype=np.int64)
        verts = [v.copy()]
        for ax in perm:
            v[ax] += 1
            verts.append(v.copy())
        tets.append(np.array(verts))
    return np.array(tets)  # (6, 4, 3)


def make_cylinder(nx: int, ny: int, nz: int, R: float = 2.0, L: float = 30.0) -> Mesh:
    """Kuhn-lattice tet cylinder with 6*nx*ny*nz elements and (nx+1)(ny+1)(nz+1) nodes."""
    if nx % 2 or ny % 2:
        raise ValueError("nx and ny must be even (quadrant-mirrored Kuhn subdivision)")
    gx = np.linspace(-1.0, 1.0, nx + 1)
    gy = np.linspace(-1.0, 1.0, ny + 1)
    gz = np.linspace(0.0, L, nz + 1)
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    rad = np.sqrt(X * X + Y * Y)
    cheb = np.maximum(np.abs(X), np.abs(Y))
    scale = np.where(rad > 0, cheb / np.where(rad > 0, rad, 1.0), 1.0)
    x = np.stack([R * X * scale, R * Y * scale, Z], axis=-1).reshape(-1, 3)

    def nid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ci = ci.ravel(); cj = cj.ravel(); ck = ck.ravel()
    fx = ci < nx // 2           # mirror in x on the -x half
    fy = cj < ny // 2
    paths = _kuhn_paths()
    ien = np.empty((ci.size, 6, 4), dtype=np.int64)
    for t in range(6):
        for v in range(4):
            dx, dy, dz = paths[t, v]
            ii = ci + np.where(fx, 1 - dx, dx)
            jj = cj + np.where(fy, 1 - dy, dy)
            kk = ck + dz
            ien[:, t, v] = nid(ii, jj, kk)
    cell_k = np.repeat(ck, 6)
    ien = ien.reshape(-1, 4)
    # svFSI orientation: swap nodes 1,2 where the Jacobian is negative (READMSH.f:1133-1149)
    a = x[ien[:, 0]] - x[ien[:, 3]]
    b = x[ien[:, 1]] - x[ien[:, 3]]
    c = x[ien[:, 2]] - x[ien[:, 3]]
    jac = np.einsum("ij,ij->i", np.cross(a, b), c)
    neg = jac < 0
    ien[neg, 0], ien[neg, 1] = ien[neg, 1].copy(), ien[neg, 0].copy()
    m = Mesh(x=np.ascontiguousarray(x), IEN=(ien + 1).astype(np.int32), cell_k=cell_k,
             dims=(nx, ny, nz), R=R, L=L)
    _make_faces(m, nx, ny, nz)
    return m


def _make_faces(m: Mesh, nx, ny, nz):
    """Boundary triangles = tet faces whose three nodes lie on the same boundary surface."""
    idx = np.arange(m.nNo)
    k = idx % (nz + 1)
    j = (idx // (nz + 1)) % (ny + 1)
    i = idx // ((nz + 1) * (ny + 1))
    on = {
        "inlet": k == 0,
        "outlet": k == nz,
        "wall": (i == 0) | (i == nx) | (j == 0) | (j == ny),
    }
    ien0 = m.IEN.astype(np.int64) - 1
    local_faces = [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]
    for name, mask in on.items():
        tris, parents = [], []
        for lf in local_faces:
            f = ien0[:, lf]
            sel = mask[f].all(axis=1)
            if name == "wall":
                # all three on the lateral surface AND on the same side plane
                ii, jj = i[f], j[f]
                same = ((ii == 0).all(1) | (ii == nx).all(1) | (jj == 0).all(1) | (jj == ny).all(1))
                sel &= same
            tris.append(f[sel])
            parents.append(np.nonzero(sel)[0])
        tri = np.concatenate(tris)
        par = np.concatenate(parents)
        gN = np.nonzero(mask)[0]
        m.faces[name] = Face(name, (gN + 1).astype(np.int32), (tri + 1).astype(np.int32), par)


def face_normal_integrals(m: Mesh, face: Face, node_ids: np.ndarray, tri_sel=None):
    """sV(:,Ac) = sum_e sum_g N(a,g) w(g) n  over the face's triangles (BAFINI.f:515-529), with n the
    outward area-weighted normal of GNNB (NN.f:1974-1991).  For TRI3 (3 Gauss points, w=1/6,
    NN.f:416-422) this is (area-normal)/3 per node.  Returns val(len(node_ids), 3) for the given
    1-based global node ids; ``tri_sel`` restricts to a subset of triangles (one rank's share)."""
    tri = face.tri if tri_sel is None else face.tri[tri_sel]
    par = face.parent if tri_sel is None else face.parent[tri_sel]
    p = m.x[tri.astype(np.int64) - 1]                     # (nTri, 3, 3)
    nrm = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])  # 2*area*n
    cen_el = m.x[m.IEN[par].astype(np.int64) - 1].mean(axis=1)
    flip = np.einsum("ij,ij->i", nrm, p.mean(axis=1) - cen_el) < 0
    nrm[flip] *= -1.0
    # sum_g N(a,g)*w(g) = 1/6 for every node of TRI3; |n| from GNNB = 2*area
    contrib = nrm / 6.0
    sV = np.zeros((m.nNo, 3))
    for v in range(3):
        np.add.at(sV, tri[:, v].astype(np.int64) - 1, contrib)
    return sV[node_ids.astype(np.int64) - 1]


# ----------------------------------------------------------------------------------------------
def gen_alpha(rho_inf: float = 0.2):
    """First-order generalised-alpha constants, INITIALIZE.f:78-79,136-139."""
    am = 0.5 * (3.0 - rho_inf) / (1.0 + rho_inf)
    af = 1.0 / (1.0 + rho_inf)
    gam = 0.5 + am - af
    beta = 0.25 * (1.0 + am - af) ** 2
    return dict(am=am, af=af, gam=gam, beta=beta)


def poiseuille_state(m: Mesh, umax: float = 10.0, dpdz: float = -1.0, pert: float = 0.01,
                     seed: int = SEED, acc: float = 0.0):
    """Yg(tnNo,4) = Poiseuille (u,v,w,p) + `pert`*umax uniform noise on velocity; Ag(tnNo,4)."""
    rng = np.random.default_rng(seed)
    r2 = (m.x[:, 0] ** 2 + m.x[:, 1] ** 2) / (m.R ** 2)
    Yg = np.zeros((m.nNo, 4))
    Yg[:, 2] = umax * np.clip(1.0 - r2, 0.0, None)
    Yg[:, :3] += pert * umax * rng.uniform(-1.0, 1.0, size=(m.nNo, 3))
    Yg[:, 3] = dpdz * (m.x[:, 2] - m.L)
    Ag = np.zeros((m.nNo, 4))
    if acc:
        Ag[:, :3] = acc * rng.uniform(-1.0, 1.0, size=(m.nNo, 3))
    return Ag, Yg


# ----------------------------------------------------------------------------------------------
@dataclass
class RankMesh:
    """One rank's share, numbered like DISTRIBUTE.f:1455-1529 (first-touch local ids)."""
    rank: int
    ltg: np.ndarray           # (nNo_local,) global node id (1-based)
    IEN: np.ndarray           # (nEl_local, 4) local node ids (1-based)
    x: np.ndarray             # (nNo_local, 3)
    elems: np.ndarray         # global 0-based element ids owned by the rank (ascending)
    faces: dict = field(default_factory=dict)   # name -> dict(gN local ids, gN_global, tri_sel)

    @property
    def nNo(self):
        return self.ltg.shape[0]

    @property
    def nEl(self):
        return self.IEN.shape[0]


def partition_slabs(m: Mesh, nparts: int) -> np.ndarray:
    """Element -> rank by axial slab with (nearly) equal element counts (stand-in for ParMETIS
    PartMeshKway, SPLIT.c:114; any element-disjoint partition is valid)."""
    nz = m.dims[2]
    bounds = np.linspace(0, nz, nparts + 1).round().astype(np.int64)
    part = np.searchsorted(bounds, m.cell_k, side="right") - 1
    return np.clip(part, 0, nparts - 1).astype(np.int32)


def partition_blocks(m: Mesh, nparts: int) -> np.ndarray:
    """Element -> rank by QUADRANT of the cross-section (sign of the centroid's x and y) times
    nparts/4 axial slabs: cut nodes on the pipe axis are shared by four ranks (eight where an axial
    cut crosses the axis), every rank has three to seven neighbours -- what a ParMETIS partition of
    a real vessel looks like and the axial slabs never do (there a node is shared by two ranks at
    most).  nparts must be a multiple of 4."""
    if nparts % 4:
        raise ValueError("partition_blocks needs nparts % 4 == 0")
    c = m.x[m.IEN.astype(np.int64) - 1].mean(axis=1)
    quad = (c[:, 0] > 0).astype(np.int32) + 2 * (c[:, 1] > 0).astype(np.int32)
    nax = nparts // 4
    nz = m.dims[2]
    bounds = np.linspace(0, nz, nax + 1).round().astype(np.int64)
    ax = np.clip(np.searchsorted(bounds, m.cell_k, side="right") - 1, 0, nax - 1).astype(np.int32)
    return (ax * 4 + quad).astype(np.int32)


def first_touch_unique(ien_flat: np.ndarray):
    """Unique values of `ien_flat` in order of first appearance, and the inverse map."""
    uniq, first, inv = np.unique(ien_flat, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank_of = np.empty_like(order)
    rank_of[order] = np.arange(order.size)
    return uniq[order], rank_of[inv]


def split_mesh(m: Mesh, part: np.ndarray, nparts: int):
    out = []
    for p in range(nparts):
        elems = np.nonzero(part == p)[0]
        ien_g = m.IEN[elems].astype(np.int64)
        ltg, inv = first_touch_unique(ien_g.ravel())
        ien_l = (inv.reshape(-1, 4) + 1).astype(np.int32)
        rm = RankMesh(rank=p, ltg=ltg.astype(np.int32), IEN=ien_l,
                      x=np.ascontiguousarray(m.x[ltg - 1]), elems=elems)
        gtl = np.zeros(m.nNo + 1, dtype=np.int64)
        gtl[ltg] = np.arange(1, ltg.size + 1)
        owned = np.zeros(m.nEl, dtype=bool)
        owned[elems] = True
        for name, fa in m.faces.items():
            # PARTFACE (DISTRIBUTE.f:1679-1698): triangles whose parent tet is local; node list =
            # every global face node present on this rank.
            loc = gtl[fa.gN]
            keep = loc != 0
            rm.faces[name] = dict(gN=loc[keep].astype(np.int32), gN_global=fa.gN[keep],
                                  tri_sel=owned[fa.parent])
        out.append(rm)
    return out


def scatter_nodal(rm: RankMesh, U: np.ndarray) -> np.ndarray:
    """LOCAL(): a rank's copy of a global nodal array (DISTRIBUTE.f:221-231)."""
    return np.ascontiguousarray(U[rm.ltg.astype(np.int64) - 1])


# ----------------------------------------------------------------------------------------------
def csr_pattern(tnNo: int, IEN: np.ndarray):
    """Node-graph CSR that svFSI's LHSA builds at setup (LHSA.f:38-262): row a lists every node
    sharing an element with a (a itself included), ascending, 1-based ``rowPtr[tnNo+1]``,
    ``colPtr[nnz]``.  Host/NumPy: this is setup the Fortran driver keeps; the synthetic harness
    needs it to feed ``FSILS_LHS_CREATE``."""
    ien = IEN.astype(np.int64) - 1
    nEl = ien.shape[0]
    keys = []
    chunk = 2_000_000
    for s in range(0, nEl, chunk):
        e = ien[s:s + chunk]
        k = (e[:, :, None] * tnNo + e[:, None, :]).reshape(-1)
        keys.append(np.unique(k))
    k = np.unique(np.concatenate(keys)) if len(keys) > 1 else keys[0]
    rows = k // tnNo
    cols = k - rows * tnNo
    counts = np.bincount(rows, minlength=tnNo)
    if (counts == 0).any():
        raise RuntimeError(f"Node {int(np.argmin(counts > 0)) + 1} is isolated")
    rowPtr = np.empty(tnNo + 1, dtype=np.int32)
    rowPtr[0] = 1
    rowPtr[1:] = 1 + np.cumsum(counts)
    return rowPtr, (cols + 1).astype(np.int32)


@dataclass
class RankProblem:
    """Everything one rank hands to the C-ABI: local mesh, CSR pattern, state, face lists."""
    rm: RankMesh
    rowPtr: np.ndarray
    colPtr: np.ndarray
    Ag: np.ndarray
    Yg: np.ndarray
    faces: dict           # name -> dict(gN local ids (1-based), val (n,3) or None, bc 'Dir'|'Neu')


def build_problem(nx, ny, nz, nparts=1, R=2.0, L=30.0, umax=10.0, pert=0.01, acc=0.0,
                  seed=SEED, partition="slabs"):
    """Synthetic pipe problem: mesh, axial-slab partition, per-rank CSR and state, and the three
    faces (inlet/wall Dirichlet with zero mask, outlet Neumann with val = int N n dGamma,
    BAFINI.f:490-533)."""
    m = make_cylinder(nx, ny, nz, R=R, L=L)
    Ag, Yg = poiseuille_state(m, umax=umax, pert=pert, seed=seed, acc=acc)
    part = partition_blocks(m, nparts) if partition == "blocks" else partition_slabs(m, nparts)
    rms = split_mesh(m, part, nparts)
    out = []
    for rm in rms:
        rowPtr, colPtr = csr_pattern(rm.nNo, rm.IEN)
        faces = {}
        for name in ("inlet", "wall", "outlet"):
            fa = rm.faces[name]
            if name == "outlet":
                val = face_normal_integrals(m, m.faces[name], fa["gN_global"], fa["tri_sel"])
                faces[name] = dict(gN=fa["gN"], val=val, bc="Neu")
            else:
                faces[name] = dict(gN=fa["gN"], val=None, bc="Dir")
        out.append(RankProblem(rm, rowPtr, colPtr, scatter_nodal(rm, Ag), scatter_nodal(rm, Yg),
                               faces))
    return m, out, (Ag, Yg)


# ----------------------------------------------------------------------------------------------
def _hash_noise(gid: np.ndarray, salt: int) -> np.ndarray:
    """Deterministic uniform(-1,1) noise keyed by GLOBAL node id, so that every rank generating
    its own slab gives a shared node the same state (DISTRIBUTE.f:221-231 semantics)."""
    with np.errstate(over="ignore"):
        return _hash_noise_impl(gid, (salt * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)


def _hash_noise_impl(gid, salt64):
    v = gid.astype(np.uint64) + np.uint64(salt64)
    v ^= v >> np.uint64(30); v *= np.uint64(0xBF58476D1CE4E5B9)
    v ^= v >> np.uint64(27); v *= np.uint64(0x94D049BB133111EB)
    v ^= v >> np.uint64(31)
    return (v >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0


def build_rank_problem(nx, ny, nz, rank=0, nparts=1, R=2.0, L=30.0, umax=10.0, pert=0.01):
    """One rank's share of the (nx, ny, nz) pipe WITHOUT building the global mesh: axial slab
    `rank` of `nparts` (same split as partition_slabs), first-touch local numbering, global ids
    identical to make_cylinder's.  Used by bench.py at the 10M-tet size, one call per GPU rank."""
    bounds = np.linspace(0, nz, nparts + 1).round().astype(np.int64)
    k0, k1 = int(bounds[rank]), int(bounds[rank + 1])
    dz = L / nz
    sl = make_cylinder(nx, ny, k1 - k0, R=R, L=(k1 - k0) * dz)
    sl.x[:, 2] += k0 * dz
    idx = np.arange(sl.nNo)
    nzl = k1 - k0
    k = idx % (nzl + 1)
    ij = idx // (nzl + 1)
    gid = (ij * (nz + 1) + k + k0 + 1).astype(np.int64)          # 1-based global node ids
    # first-touch local numbering over the slab's elements (DISTRIBUTE.f:1455-1467)
    order, inv = first_touch_unique(sl.IEN.astype(np.int64).ravel() - 1)
    ien_l = (inv.reshape(-1, 4) + 1).astype(np.int32)
    ltg = gid[order].astype(np.int32)
    x = np.ascontiguousarray(sl.x[order])
    old2new = np.empty(sl.nNo, dtype=np.int64)
    old2new[order] = np.arange(1, sl.nNo + 1)
    rm = RankMesh(rank=rank, ltg=ltg, IEN=ien_l, x=x, elems=np.zeros(0, dtype=np.int64))
    rowPtr, colPtr = csr_pattern(rm.nNo, rm.IEN)
    faces = {}
    for name in ("inlet", "wall", "outlet"):
        fa = sl.faces[name]
        present = (name == "wall") or (name == "inlet" and k0 == 0) or (name == "outlet" and k1 == nz)
        if not present:
            faces[name] = dict(gN=np.zeros(0, dtype=np.int32), val=None if name != "outlet" else np.zeros((0, 3)),
                               bc="Neu" if name == "outlet" else "Dir", IEN=np.zeros((0, 3), dtype=np.int32),
                               gE=np.zeros(0, dtype=np.int32))
            continue
        gN_old = fa.gN.astype(np.int64)
        if name == "outlet":
            val = face_normal_integrals(sl, fa, fa.gN)
            # faceType IEN / gE of the rank's share (S/DISTRIBUTE.f:1679-1689): boundary triangles in
            # local node ids, parent tets as local element ids (1-based)
            faces[name] = dict(gN=old2new[gN_old - 1].astype(np.int32), val=val, bc="Neu",
                               IEN=old2new[fa.tri.astype(np.int64) - 1].astype(np.int32),
                               gE=(fa.parent + 1).astype(np.int32))
        else:
            faces[name] = dict(gN=old2new[gN_old - 1].astype(np.int32), val=None, bc="Dir")
    g = ltg.astype(np.int64)
    r2 = (x[:, 0] ** 2 + x[:, 1] ** 2) / (R ** 2)
    Yg = np.zeros((rm.nNo, 4))
    Yg[:, 2] = umax * np.clip(1.0 - r2, 0.0, None)
    for c in range(3):
        Yg[:, c] += pert * umax * _hash_noise(g, c + 1)
    Yg[:, 3] = -1.0 * (x[:, 2] - L)
    Ag = np.zeros((rm.nNo, 4))
    gnNo = (nx + 1) * (ny + 1) * (nz + 1)
    return gnNo, RankProblem(rm, rowPtr, colPtr, Ag, Yg, faces)
