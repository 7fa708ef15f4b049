"""Shared helpers of the parity tests: run the CPU oracle (oracle/) end to end on the same
synthetic problem the CUDA path gets.  Test infrastructure only."""
import os

import numpy as np

from oracle import oracle as ora
from svfsi_b200 import mesh

RHO, MU, DT = 1.06, 0.04, 5e-3
F = (0.0, 0.0, 0.0)
GA = mesh.gen_alpha(0.2)
FACE_ORDER = ("inlet", "wall", "outlet")     # faIn = 1, 2, 3


def fluid_par():
    return ora.fluid_par(RHO, MU, F, DT, GA["af"], GA["am"], GA["gam"])


def rel_err(a, b):
    """max|a-b| / max|b|  (SURVEY.md 8c parity metric)"""
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def block_class_errs(Val, Val_ref):
    """per dof-block class of the 4x4 tangent: momentum 3x3, pressure column, continuity row, dC/dP"""
    V = Val.reshape(-1, 4, 4); W = Val_ref.reshape(-1, 4, 4)
    return dict(KK=rel_err(V[:, :3, :3], W[:, :3, :3]), G=rel_err(V[:, :3, 3], W[:, :3, 3]),
                D=rel_err(V[:, 3, :3], W[:, 3, :3]), L=rel_err(V[:, 3, 3], W[:, 3, 3]))


def oracle_world(probs, gnNo, nFaces=3, with_faces=True, native=False):
    w = ora.World(gnNo, [p.rm.ltg for p in probs], [p.rowPtr for p in probs],
                  [p.colPtr for p in probs], nFaces, native=native)
    if with_faces:
        for fi, name in enumerate(FACE_ORDER, start=1):
            bc = ora.BC_TYPE_Neu if probs[0].faces[name]["bc"] == "Neu" else ora.BC_TYPE_Dir
            vals = [p.faces[name]["val"] for p in probs]
            w.bc_create(fi, [p.faces[name]["gN"] for p in probs], 3, bc,
                        None if vals[0] is None else vals)
    return w


def oracle_assemble(probs, native=False):
    par = fluid_par()
    Rs, Vs = [], []
    for p in probs:
        R, V = ora.construct_fluid(par, p.rm.IEN, p.rm.x, p.Ag, p.Yg, np.zeros((p.rm.nNo, 3)),
                                   p.rowPtr, p.colPtr, native=native)
        Rs.append(R); Vs.append(V)
    return Rs, Vs


def oracle_commu(w, probs, Rs, dof=4):
    """COMMU(R), S/ALLFUN.f:514-533: permute -> FSILS_COMMUV -> unpermute"""
    maps = [w.map(r).astype(np.int64) - 1 for r in range(len(probs))]
    tmp = []
    for R, mp in zip(Rs, maps):
        t = np.zeros_like(R); t[mp] = R; tmp.append(t)
    w.commuv(dof, tmp)
    return [t[mp].copy() for t, mp in zip(tmp, maps)]


def log_parity(name, **vals):
    """every measured parity error goes to stdout and, when SVFSI_PARITY_LOG names a file, into it
    (profiles/r02_parity.log is such a file from a B200 run)"""
    line = name + ": " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items())
    print(line, flush=True)
    path = os.environ.get("SVFSI_PARITY_LOG")
    if path:
        with open(path, "a") as fh:
            fh.write(line + "\n")


def oracle_gmres_global(nparts, relTol, sD, mItr, res_out, dims=(8, 8, 20), L=4.0,
                        ls_type=None, prec=ora.PRECOND_FSILS, perturb=None, partition="slabs", **lskw):
    """Assemble + COMMU + FSILS_SOLVE with the oracle on `nparts` simulated ranks; returns
    (ls, X_global) with X gathered by global node id."""
    m, probs, _ = mesh.build_problem(*dims, nparts=nparts, L=L, partition=partition)
    Rs, Vs = oracle_assemble(probs)
    if perturb is not None:      # rounding-level relative noise on the assembled system
        rng = np.random.default_rng(perturb)
        Vs = [v * (1.0 + 2e-16 * rng.standard_normal(v.shape)) for v in Vs]
        Rs = [r * (1.0 + 2e-16 * rng.standard_normal(r.shape)) for r in Rs]
    w = oracle_world(probs, m.nNo)
    Rc = oracle_commu(w, probs, Rs)
    ls = ora.ls_create(ls_type or ora.LS_TYPE_GMRES, relTol=relTol, absTol=1e-14, maxItr=mItr,
                       dimKry=sD, **{k: v for k, v in lskw.items() if v is not None})
    X = [r.copy() for r in Rc]
    w.solve(ls, 4, X, [v.copy() for v in Vs], prec=prec, incL=[1, 1, 1],
            res=np.array([0.0, 0.0, res_out]))
    G = np.zeros((m.nNo, 4))
    for p, x in zip(probs, X):
        G[p.rm.ltg - 1] = x
    return ls, G


def reference_reproducibility_floor(relTol, sD, mItr, res_out, ls_type=None, **kw):
    """How far the REFERENCE ALGORITHM moves when only its summation order changes: oracle on
    2 and 3 simulated MPI ranks vs 1 rank (the reference's own partition-to-partition drift,
    SURVEY.md Appendix F).  Returns (floor on ||dX||/||X||, floor on |d fNorm|/fNorm, max |d itr|)."""
    ls1, G1 = oracle_gmres_global(1, relTol, sD, mItr, res_out, ls_type=ls_type, **kw)
    fx = ff = 0.0; di = 0
    for k in (2, 3):
        lsk, Gk = oracle_gmres_global(k, relTol, sD, mItr, res_out, ls_type=ls_type, **kw)
        fx = max(fx, float(np.linalg.norm(Gk - G1) / np.linalg.norm(G1)))
        ff = max(ff, abs(lsk.RI.fNorm - ls1.RI.fNorm) / ls1.RI.fNorm)
        di = max(di, abs(lsk.RI.itr - ls1.RI.itr))
    return fx, ff, di


def rounding_floor(relTol, sD, mItr, res_out, ls_type=None, seeds=(1, 2, 3), with_range=False, **kw):
    """Same idea as reference_reproducibility_floor but partition-independent: how far the
    reference algorithm (oracle, 1 rank) moves when the assembled R / Val are perturbed by one
    ulp of relative noise -- the size of the difference between any two correct FP64 evaluations
    of the element loop.  Returns (floor on ||dX||/||X||, floor on |d fNorm|/fNorm, max |d itr|)."""
    ls1, G1 = oracle_gmres_global(1, relTol, sD, mItr, res_out, ls_type=ls_type, **kw)
    fx = ff = 0.0; di = 0
    lo = hi = ls1.RI.itr
    for sd in seeds:
        lsk, Gk = oracle_gmres_global(1, relTol, sD, mItr, res_out, ls_type=ls_type, perturb=sd, **kw)
        fx = max(fx, float(np.linalg.norm(Gk - G1) / np.linalg.norm(G1)))
        ff = max(ff, abs(lsk.RI.fNorm - ls1.RI.fNorm) / ls1.RI.fNorm)
        di = max(di, abs(lsk.RI.itr - ls1.RI.itr))
        lo, hi = min(lo, lsk.RI.itr), max(hi, lsk.RI.itr)
    if with_range:      # the spread of the reference algorithm's own iteration count under that noise
        return fx, ff, di, (lo, hi)
    return fx, ff, di


def local_face(m, rm, name):
    """One rank's share of mesh face `name` in svFSI's local numbering (faceType gN / IEN / gE,
    1-based), as PARTFACE leaves it (S/DISTRIBUTE.f:1679-1698)."""
    fa = m.faces[name]
    sel = rm.faces[name]["tri_sel"]
    gtl = np.zeros(m.nNo + 1, dtype=np.int64)
    gtl[rm.ltg] = np.arange(1, rm.ltg.size + 1)
    fIEN = gtl[fa.tri[sel].astype(np.int64)].astype(np.int32)
    gE = (np.searchsorted(rm.elems, fa.parent[sel]) + 1).astype(np.int32)
    return rm.faces[name]["gN"], fIEN, gE


class OracleCplBC:
    """SETBCCPL / CALCDERCPLBC / RCRINIT sequencing (S/SETBC.f:981-1123, S/BAFINI.f:69-104, S/MAIN.f:280)
    around the ORACLE's own RCR_Integ_X restatement (oracle.rcr_integ_x / calc_der_cplbc), so that the
    RCR parity tests do not compare the product's `svfsi_b200/cplbc.py` with itself.  `faces` only needs
    the attributes Rp, C, Rd, Pd, Xo."""

    def __init__(self, faces, dt, scheme="SI"):
        self.Rp = [f.Rp for f in faces]; self.C = [f.C for f in faces]
        self.Rd = [f.Rd for f in faces]; self.Pd = [f.Pd for f in faces]
        self.Xo0 = [f.Xo for f in faces]
        self.n, self.dt, self.schm = len(faces), float(dt), scheme
        self.xo = np.zeros(self.n); self.xn = np.zeros(self.n); self.y = np.zeros(self.n)
        self.r = np.zeros(self.n); self.Qo = np.zeros(self.n); self.Qn = np.zeros(self.n)

    def _fluxes(self, integ):
        for i in range(self.n):
            self.Qo[i] = integ(i, "o"); self.Qn[i] = integ(i, "n")

    def _calcder(self, integ, time):
        self._fluxes(integ)
        self.xn, _ = ora.rcr_integ_x(self.xo, self.Qo, self.Qn, self.Rp, self.C, self.Rd, self.Pd, self.dt, time)
        y, r = ora.calc_der_cplbc(self.xo, self.Qo, self.Qn, self.Rp, self.C, self.Rd, self.Pd, self.dt, time)
        self.y, self.r = np.array(y, dtype=np.float64), np.array(r, dtype=np.float64)

    def init(self, integ, time=0.0):
        self.xo = np.array(self.Xo0, dtype=np.float64)
        self.y[:] = 0.0
        if self.schm != "E":
            self._calcder(integ, time)

    def setbccpl(self, integ, time):
        if self.schm == "I":
            self._calcder(integ, time)
        else:
            self._fluxes(integ)
            self.xn, y = ora.rcr_integ_x(self.xo, self.Qo, self.Qn, self.Rp, self.C, self.Rd, self.Pd, self.dt, time)
            self.y = np.array(y, dtype=np.float64)
        return self.y.copy()

    def advance(self):
        self.xo = np.array(self.xn, dtype=np.float64).copy()


def oracle_rcr_time_loop(m, p, faces_rcr, nsteps=2, nnewton=3, relTol=1e-5, sD=80, umax_in=-12.0, scheme="SI"):
    """The shape of BASELINE configs[0] (04-fluid/01-pipe3D_RCR) on one rank with the ORACLE: steady
    parabolic Dirichlet inlet, no-slip wall, RCR outlet.  Per Newton iteration (S/MAIN.f:111-206):
    SETBCCPL (fluxes of Yo / Yn through the outlet -> RCR_Integ_X -> g) -> PICI -> element loop ->
    Neumann face with h = g (S/SETBC.f:267-270, BASSEMNEUBC) -> FSILS_SOLVE with res = gam*dt*r -> PICC;
    per time step cplBC%xo = cplBC%xn.  Returns per-iteration (iNorm, itr, g, Qn) and the final (An, Yn)."""
    ga = GA
    nNo = p.rm.nNo
    rng = np.random.default_rng(21)
    Ao = 0.05 * rng.standard_normal((nNo, 4)); Ao[:, 3] = 0.0
    Yo = p.Yg.copy()
    gin = p.faces["inlet"]["gN"]; gw = p.faces["wall"]["gN"]
    xin = p.rm.x[gin - 1]
    r2 = (xin[:, 0] ** 2 + xin[:, 1] ** 2) / (np.abs(p.rm.x[:, :2]).max() ** 2)
    gx = np.clip(1.0 - r2, 0.0, None)
    nV = np.tile(np.array([0.0, 0.0, -1.0]), (gin.size, 1))
    tA_in, tY_in = ora.setbcdirl(umax_in, gx, nV, 3)
    tA_w, tY_w = np.zeros((gw.size, 3)), np.zeros((gw.size, 3))
    gout, fIEN, gE = local_face(m, p.rm, "outlet")
    w = oracle_world([p], m.nNo)
    par = fluid_par()
    cpl = OracleCplBC(faces_rcr, DT, scheme)     # the ORACLE's 0-D restatement, not the product's cplbc.py
    state = dict(Yo=Yo, Yn=Yo)

    def integ(i, which):
        return ora.integ_v(p.rm.x, p.rm.IEN, fIEN, gE, state["Y" + which][:, :3])
    cpl.init(integ)
    out = []
    time = 0.0
    for ts in range(nsteps):
        time += DT
        An, Yn = ora.picp(Ao, Yo, ga["gam"])
        ora.setbcdir(An, Yn, gin, 1, tA_in, tY_in)
        ora.setbcdir(An, Yn, gw, 1, tA_w, tY_w)
        for it in range(nnewton):
            state["Yo"], state["Yn"] = Yo, Yn
            g = cpl.setbccpl(integ, time)[0]
            Ag, Yg = ora.pici(Ao, An, Yo, Yn, ga["am"], ga["af"])
            R, V = ora.construct_fluid(par, p.rm.IEN, p.rm.x, Ag, Yg, np.zeros((nNo, 3)), p.rowPtr, p.colPtr)
            hg = np.zeros(nNo); hg[gout - 1] = -g * 1.0
            ora.bassem_neu_fluid(p.rm.x, p.rm.IEN, fIEN, gE, hg, Yg, p.rowPtr, p.colPtr, R, V, RHO, 0.2,
                                 ga["af"], ga["gam"], DT)
            ls = ora.ls_create(ora.LS_TYPE_GMRES, relTol=relTol, absTol=1e-14, maxItr=10, dimKry=sD)
            w.solve(ls, 4, [R], [V], incL=[1, 1, 1], res=[0.0, 0.0, ga["gam"] * DT * cpl.r[0]])
            out.append((ls.RI.iNorm, ls.RI.itr, g, cpl.Qn[0]))
            ora.picc(An, Yn, R, ga["gam"], ga["beta"], DT)
        Ao, Yo = An, Yn
        cpl.advance()
    return out, (Ao, Yo), cpl
