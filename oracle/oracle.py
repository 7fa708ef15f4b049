"""ctypes front-end of the CPU oracle (oracle/*.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is the checker, never the product path.  How it is pinned to the reference: see ora.h
(oracle/refexec.py + tests/golden/make_ref_golden.py + tests/test_reference_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

LS_TYPE_CG, LS_TYPE_GMRES, LS_TYPE_NS, LS_TYPE_BICGS = 798, 797, 796, 795
PRECOND_FSILS, PRECOND_RCS = 701, 709
BC_TYPE_Dir, BC_TYPE_Neu = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class FluidPar(C.Structure):
    _fields_ = [("rho", C.c_double), ("mu", C.c_double), ("f", C.c_double * 3),
                ("dt", C.c_double), ("af", C.c_double), ("am", C.c_double), ("gam", C.c_double)]


class HeatPar(C.Structure):
    _fields_ = [("nu", C.c_double), ("s", C.c_double), ("rho", C.c_double),
                ("dt", C.c_double), ("af", C.c_double), ("am", C.c_double), ("gam", C.c_double)]


class SubLs(C.Structure):
    _fields_ = [("suc", C.c_int), ("mItr", C.c_int), ("sD", C.c_int), ("itr", C.c_int),
                ("absTol", C.c_double), ("relTol", C.c_double), ("iNorm", C.c_double),
                ("fNorm", C.c_double), ("dB", C.c_double), ("callD", C.c_double)]


class Ls(C.Structure):
    _fields_ = [("LS_type", C.c_int), ("Resm", C.c_int), ("Resc", C.c_int),
                ("GM", SubLs), ("CG", SubLs), ("RI", SubLs)]


def build(native: bool = False, force: bool = False) -> str:
    name = "libsvfsi_oracle_native.so" if native else "libsvfsi_oracle.so"
    path = os.path.join(_HERE, name)
    srcs = [os.path.join(_HERE, f) for f in ("ora_elem.c", "ora_fsils.c", "ora.h")]
    stale = (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return path


_libs = {}


def lib(native: bool = False):
    if native not in _libs:
        L = C.CDLL(build(native))
        L.ora_wtime.restype = C.c_double
        L.ora_dotv.restype = C.c_double
        L.ora_normv.restype = C.c_double
        L.ora_world_create.restype = C.c_void_p
        _libs[native] = L
    return _libs[native]


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _pp(arrs):
    """array of per-rank pointers"""
    return (C.c_void_p * len(arrs))(*[a.ctypes.data if a is not None else None for a in arrs])


def fluid_par(rho, mu, f, dt, af, am, gam):
    return FluidPar(rho, mu, (C.c_double * 3)(*f), dt, af, am, gam)


def heat_par(nu, s, rho, dt, af, am, gam):
    return HeatPar(nu, s, rho, dt, af, am, gam)


# ------------------------------------------------------------------------------------------
def tet4_tables():
    w = np.zeros(4); N = np.zeros((4, 4)); Nxi = np.zeros((4, 3))
    lib().ora_tet4_tables(_d(w), _d(N), _d(Nxi))
    return w, N, Nxi


def fluid_element(par, xl, al, yl, bfl):
    """lR[a][i], lK[b][a][16] for one element (S/FLUID.f:84-168)."""
    xl = np.ascontiguousarray(xl, dtype=np.float64); al = np.ascontiguousarray(al, dtype=np.float64)
    yl = np.ascontiguousarray(yl, dtype=np.float64); bfl = np.ascontiguousarray(bfl, dtype=np.float64)
    lR = np.zeros((4, 4)); lK = np.zeros((4, 4, 16)); flag = C.c_int(0)
    lib().ora_fluid_element(C.byref(par), _d(xl), _d(al), _d(yl), _d(bfl), _d(lR), _d(lK), C.byref(flag))
    return lR, lK, flag.value


def heat_element(par, xl, al, yl):
    xl = np.ascontiguousarray(xl, dtype=np.float64)
    al = np.ascontiguousarray(al, dtype=np.float64); yl = np.ascontiguousarray(yl, dtype=np.float64)
    lR = np.zeros(4); lK = np.zeros((4, 4)); flag = C.c_int(0)
    lib().ora_heat_element(C.byref(par), _d(xl), _d(al), _d(yl), _d(lR), _d(lK), C.byref(flag))
    return lR, lK, flag.value


def lhsa(tnNo, IEN):
    """(rowPtr[tnNo+1], colPtr[nnz]) 1-based, S/LHSA.f:38-262."""
    IEN = np.ascontiguousarray(IEN, dtype=np.int32)
    rp = _ip(); cp = _ip(); nnz = C.c_int(0)
    iso = lib().ora_lhsa(int(tnNo), int(IEN.shape[0]), _i(IEN), C.byref(rp), C.byref(cp), C.byref(nnz))
    rowPtr = np.ctypeslib.as_array(rp, shape=(tnNo + 1,)).copy()
    colPtr = np.ctypeslib.as_array(cp, shape=(max(nnz.value, 1),))[: nnz.value].copy()
    lib().ora_free(rp); lib().ora_free(cp)
    if iso:
        raise RuntimeError(f"Node {iso} is isolated")
    return rowPtr, colPtr


def construct_fluid(par, IEN, x, Ag, Yg, Bf, rowPtr, colPtr, faithful=False, native=False):
    IEN = np.ascontiguousarray(IEN, dtype=np.int32)
    tnNo = x.shape[0]; nnz = colPtr.shape[0]
    R = np.zeros((tnNo, 4)); Val = np.zeros((nnz, 16))
    nbad = lib(native).ora_construct_fluid(
        C.byref(par), int(IEN.shape[0]), _i(IEN), _d(np.ascontiguousarray(x)),
        _d(np.ascontiguousarray(Ag)), _d(np.ascontiguousarray(Yg)), _d(np.ascontiguousarray(Bf)),
        _i(rowPtr), _i(colPtr), _d(R), _d(Val), int(faithful))
    if nbad:
        raise RuntimeError(f"Jac < 0 @ {nbad} elements")
    return R, Val


def construct_heats(par, IEN, x, Ag, Yg, rowPtr, colPtr, native=False):
    IEN = np.ascontiguousarray(IEN, dtype=np.int32)
    tnNo = x.shape[0]; nnz = colPtr.shape[0]
    R = np.zeros(tnNo); Val = np.zeros(nnz)
    nbad = lib(native).ora_construct_heats(
        C.byref(par), int(IEN.shape[0]), _i(IEN), _d(np.ascontiguousarray(x)),
        _d(np.ascontiguousarray(Ag)), _d(np.ascontiguousarray(Yg)), _i(rowPtr), _i(colPtr), _d(R), _d(Val))
    if nbad:
        raise RuntimeError(f"Jac < 0 @ {nbad} elements")
    return R, Val


# ------------------------------------------------------------------------------------------
class World:
    """nTasks simulated FSILS ranks (FSILS_LHS_CREATE on each, L/LHS.f:51-293)."""

    def __init__(self, gnNo, ltg_list, rowPtr_list, colPtr_list, nFaces, native=False):
        self.L = lib(native)
        self.n = len(ltg_list)
        self.ltg = [np.ascontiguousarray(a, dtype=np.int32) for a in ltg_list]
        self.rowPtr = [np.ascontiguousarray(a, dtype=np.int32) for a in rowPtr_list]
        self.colPtr = [np.ascontiguousarray(a, dtype=np.int32) for a in colPtr_list]
        self.nNo = np.array([a.size for a in self.ltg], dtype=np.int32)
        self.nnz = np.array([a.size for a in self.colPtr], dtype=np.int32)
        self.nFaces = nFaces
        self.h = C.c_void_p(self.L.ora_world_create(
            self.n, int(gnNo), _i(self.nNo), _i(self.nnz), _pp(self.ltg), _pp(self.rowPtr),
            _pp(self.colPtr), int(nFaces)))

    def __del__(self):
        try:
            if self.h:
                self.L.ora_world_free(self.h)
                self.h = None
        except Exception:
            pass

    def info(self, r):
        a = C.c_int(); b = C.c_int(); c = C.c_int()
        self.L.ora_world_info(self.h, r, C.byref(a), C.byref(b), C.byref(c))
        return dict(mynNo=a.value, shnNo=b.value, nReq=c.value)

    def map(self, r):
        m = np.zeros(self.nNo[r], dtype=np.int32)
        self.L.ora_world_map(self.h, r, _i(m))
        return m

    def rowptr(self, r):
        rp = np.zeros((self.nNo[r], 2), dtype=np.int32); cp = np.zeros(self.nnz[r], dtype=np.int32)
        dp = np.zeros(self.nNo[r], dtype=np.int32)
        self.L.ora_world_rowptr(self.h, r, _i(rp), _i(cp), _i(dp))
        return rp, cp, dp

    def cs(self, r):
        out = []
        for i in range(self.info(r)["nReq"]):
            iP = C.c_int(); n = C.c_int()
            self.L.ora_world_cs(self.h, r, i, C.byref(iP), C.byref(n), None)
            ptr = np.zeros(n.value, dtype=np.int32)
            self.L.ora_world_cs(self.h, r, i, C.byref(iP), C.byref(n), _i(ptr))
            out.append((iP.value, ptr))
        return out

    def bc_create(self, faIn, gNodes_list, dof, BC_type, val_list=None):
        gN = [np.ascontiguousarray(g, dtype=np.int32) for g in gNodes_list]
        nNo = np.array([g.size for g in gN], dtype=np.int32)
        if val_list is None:
            vp = None
        else:
            vals = [np.ascontiguousarray(v, dtype=np.float64) for v in val_list]
            vp = _pp(vals)
        self.L.ora_bc_create(self.h, int(faIn), _i(nNo), int(dof), int(BC_type), _pp(gN), vp)

    def commuv(self, dof, R_list):
        self.L.ora_commuv(self.h, int(dof), _pp(R_list))

    def sparmul_vv(self, dof, K_list, U_list):
        KU = [np.zeros_like(u) for u in U_list]
        self.L.ora_sparmul_vv(self.h, int(dof), _pp(K_list), _pp(U_list), _pp(KU))
        return KU

    def sparmul_ss(self, K_list, U_list):
        KU = [np.zeros_like(u) for u in U_list]
        self.L.ora_sparmul_ss(self.h, _pp(K_list), _pp(U_list), _pp(KU))
        return KU

    def sparmul_vs(self, dof, K_list, U_list):
        KU = [np.zeros(u.shape[0]) for u in U_list]
        self.L.ora_sparmul_vs(self.h, int(dof), _pp(K_list), _pp(U_list), _pp(KU))
        return KU

    def sparmul_sv(self, dof, K_list, U_list):
        KU = [np.zeros((u.shape[0], dof)) for u in U_list]
        self.L.ora_sparmul_sv(self.h, int(dof), _pp(K_list), _pp(U_list), _pp(KU))
        return KU

    def dotv(self, dof, U_list, V_list):
        return self.L.ora_dotv(self.h, int(dof), _pp(U_list), _pp(V_list))

    def solve(self, ls: Ls, dof, Ri_list, Val_list, prec=PRECOND_FSILS, incL=None, res=None):
        """FSILS_SOLVE (L/SOLVE.f:51-143): Ri_list/Val_list are modified in place."""
        incp = _i(np.ascontiguousarray(incL, dtype=np.int32)) if incL is not None else None
        resp = _d(np.ascontiguousarray(res, dtype=np.float64)) if res is not None else None
        self.L.ora_fsils_solve(self.h, C.byref(ls), int(dof), _pp(Ri_list), _pp(Val_list),
                               int(prec), incp, resp)
        return ls


def ls_create(LS_type, relTol=None, absTol=None, maxItr=None, dimKry=None,
              relTolIn=None, absTolIn=None, maxItrIn=None) -> Ls:
    """FSILS_LS_CREATE with the optional overrides of L/LS.f:97-116."""
    ls = Ls()
    lib().ora_ls_create(C.byref(ls), int(LS_type))
    if relTol is not None: ls.RI.relTol = relTol
    if absTol is not None: ls.RI.absTol = absTol
    if maxItr is not None: ls.RI.mItr = maxItr
    if dimKry is not None:
        ls.RI.sD = dimKry; ls.GM.sD = dimKry
    if relTolIn is not None:
        ls.GM.relTol, ls.CG.relTol = relTolIn
    if absTolIn is not None:
        ls.GM.absTol, ls.CG.absTol = absTolIn
    if maxItrIn is not None:
        ls.GM.mItr, ls.CG.mItr = maxItrIn
    return ls


def ge(A, B):
    A = np.asfortranarray(A, dtype=np.float64); B = np.array(B, dtype=np.float64)
    ok = lib().ora_ge(A.shape[0], B.size, _d(A), _d(B))
    return bool(ok), B


# ---------------------------------------------------------------------------------------------
# Generalised-alpha predictor / initiator / corrector and the strong Dirichlet overwrite, for one
# equation occupying dofs s..e of (tnNo, tDof) row-major arrays (NumPy restatement; arrays are the
# transposes of the Fortran (tDof, tnNo) ones).
def picp(Ao, Yo, gam, Do=None):
    """S/PIC.f:82-126 (no IB / prestress / dFlag branches): An = Ao*(gam-1)/gam; Yn = Yo; Dn = Do"""
    coef = (gam - 1.0) / gam
    An = Ao * coef
    Yn = Yo.copy()
    return (An, Yn) if Do is None else (An, Yn, Do.copy())


def setbcdir(lA, lY, gN, s, tmpA, tmpY):
    """S/SETBC.f:118-123 (std/ustd Dirichlet, no eDrn, not impD): lA(s:e,Ac) = tmpA(:,a);
    lY(s:e,Ac) = tmpY(:,a).  gN 1-based svFSI local ids; s 1-based first dof."""
    tmpA = np.atleast_2d(np.asarray(tmpA, dtype=np.float64).T).T
    tmpY = np.atleast_2d(np.asarray(tmpY, dtype=np.float64).T).T
    lDof = tmpA.shape[1]
    idx = np.asarray(gN, dtype=np.int64) - 1
    if lA.ndim == 1:
        lA[idx] = tmpA[:, 0]; lY[idx] = tmpY[:, 0]
    else:
        lA[idx, s - 1:s - 1 + lDof] = tmpA
        lY[idx, s - 1:s - 1 + lDof] = tmpY


def setbcdirl(g, gx, nV=None, lDof=3, dirA=0.0):
    """S/SETBC.f:202-232 SETBCDIRL for a steady (std) profile: lA = dirA*gx*nV, lY = g*gx*nV when
    lDof == nsd, else the scalar profile repeated over the dofs."""
    gx = np.asarray(gx, dtype=np.float64)
    if nV is not None and lDof == nV.shape[1]:
        return dirA * gx[:, None] * nV, g * gx[:, None] * nV
    return np.repeat((dirA * gx)[:, None], lDof, 1), np.repeat((g * gx)[:, None], lDof, 1)


def pici(Ao, An, Yo, Yn, am, af, Do=None, Dn=None):
    """S/PIC.f:141-152"""
    c1, c2, c3, c4 = 1.0 - am, am, 1.0 - af, af
    Ag = Ao * c1 + An * c2
    Yg = Yo * c3 + Yn * c4
    if Do is None:
        return Ag, Yg
    return Ag, Yg, Do * c3 + Dn * c4


def picc(An, Yn, R, gam, beta, dt, Dn=None):
    """S/PIC.f:203-207 (the non-sstEq branch), in place"""
    c1, c2 = gam * dt, beta * dt * dt
    An -= R
    Yn -= R * c1
    if Dn is not None:
        Dn -= R * c2


# ---------------------------------------------------------------------------------------------
# Face integrals (TRI3 faces of a TET4 mesh), NumPy restatement vectorised over face elements.
_TRI3_W = 1.0 / 6.0                                             # S/NN.f:417


def tri3_N(g):
    """S/NN.f:416-422 (Gauss points) and :1100-1103 (shape functions)"""
    s, t = 2.0 / 3.0, 1.0 / 6.0
    x1 = s if g == 1 else t
    x2 = s if g == 2 else t
    return np.array([x1, x2, 1.0 - x1 - x2])


def gnnb_tri3(x, fIEN, IEN, gE):
    """GNNB, S/NN.f:1856-1996 for TRI3-in-TET4: n = CROSS(xXi) with xXi(:,i) = sum_a Nx(i,a) lX(:,ptr(a)),
    Nx = [[1,0,-1],[0,1,-1]]; flipped so that n.(x_ptr(1) - x_offface) >= 0.  Ids 1-based.  Returns the
    area-weighted normals (nEl, 3)."""
    f = np.asarray(fIEN, dtype=np.int64) - 1
    par = np.asarray(IEN, dtype=np.int64)[np.asarray(gE, dtype=np.int64) - 1] - 1     # (nEl, 4)
    onface = (par[:, :, None] == f[:, None, :]).any(axis=2)
    if not (onface.sum(axis=1) == 3).all():
        raise ValueError("could not find matching face nodes")
    opp = par[~onface]                                               # one per element
    x0, x1, x2 = x[f[:, 0]], x[f[:, 1]], x[f[:, 2]]
    a = (0.0 + 1.0 * x0 + 0.0 * x1) + (-1.0) * x2
    b = (0.0 + 0.0 * x0 + 1.0 * x1) + (-1.0) * x2
    n = np.cross(a, b)
    v = x0 - x[opp]
    flip = np.einsum("ij,ij->i", n, v) < 0.0
    n[flip] = -n[flip]
    return n


def bassem_neu_fluid(x, IEN, fIEN, gE, hg, Yg, rowPtr, colPtr, R, Val, rho, bfStab, af, gam, dt):
    """BASSEMNEUBC (S/EQASSEM.f:90-192) + BFLUID (S/FLUID.f:1279-1336, nsd = 3, no mesh motion) +
    DOASSEM (S/LHSA.f:266-298), in place on R (tnNo, 4) and Val (nnz, 16).  hg: (tnNo,) nodal
    Neumann values; Yg: (tnNo, 4).  rowPtr / colPtr 1-based CSR of svFSI."""
    f = np.asarray(fIEN, dtype=np.int64) - 1
    nEl = f.shape[0]
    n = gnnb_tri3(x, fIEN, IEN, gE)
    Jac = np.sqrt((n * n).sum(axis=1))
    nV = n / Jac[:, None]
    w = _TRI3_W * Jac
    wl = w * af * gam * dt
    hl = hg[f]                       # (nEl, 3)
    yl = Yg[f][:, :, :3]             # (nEl, 3 nodes, 3)
    lR = np.zeros((nEl, 3, 3)); lK = np.zeros((nEl, 3, 3))
    for g in range(3):
        N = tri3_N(g)
        h = np.zeros(nEl); u = np.zeros((nEl, 3))
        for a in range(3):
            h = h + N[a] * hl[:, a]
            u = u + N[a] * yl[:, a, :]
        udn = np.zeros(nEl)
        for i in range(3):
            udn = udn + u[:, i] * nV[:, i]
        udn = 0.5 * bfStab * rho * (udn - np.abs(udn))
        hc = h[:, None] * nV + udn[:, None] * u
        for a in range(3):
            for i in range(3):
                lR[:, a, i] = lR[:, a, i] - w * N[a] * hc[:, i]
            for b in range(3):
                lK[:, a, b] = lK[:, a, b] - wl * N[a] * N[b] * udn
    rp = np.asarray(rowPtr, dtype=np.int64) - 1
    cp = np.asarray(colPtr, dtype=np.int64) - 1
    V = Val.reshape(-1, 16)
    for e in range(nEl):             # element order = accumulation order of the reference
        for a in range(3):
            row = f[e, a]
            R[row, :3] += lR[e, a]
            for b in range(3):
                seg = cp[rp[row]:rp[row + 1]]
                p = rp[row] + int(np.searchsorted(seg, f[e, b]))
                V[p, 0] += lK[e, a, b]; V[p, 5] += lK[e, a, b]; V[p, 10] += lK[e, a, b]
    return lR, lK


def integ_v(x, IEN, fIEN, gE, S):
    """IntegV, S/ALLFUN.f:199-262: flux of the nodal vector S (tnNo, 3) through the face (one rank)"""
    f = np.asarray(fIEN, dtype=np.int64) - 1
    n = gnnb_tri3(x, fIEN, IEN, gE)
    tot = 0.0
    acc = np.zeros(f.shape[0])
    for g in range(3):
        N = tri3_N(g)
        sHat = np.zeros(f.shape[0])
        for a in range(3):
            for i in range(3):
                sHat = sHat + N[a] * S[f[:, a], i] * n[:, i]
        acc = acc + _TRI3_W * sHat
    for v in acc:
        tot = tot + v
    return tot


# ------------------------------------------------------------------------------------------
# RCR (Windkessel) 0-D coupling, host side of the reference: RCR_Integ_X and CALCDERCPLBC
def rcr_integ_x(xo, Qo, Qn, Rp, C, Rd, Pd, dt, time, nTS=100):
    """RCR_Integ_X, S/SETBC.f:1292-1372, one face at a time in scalar arithmetic (loop order of the
    Fortran): returns (xn[nFa], y[nFa])"""
    nX = len(xo)
    X = [float(v) for v in xo]
    tt = max(time - dt, 0.0)
    dtt = dt / float(nTS)
    Qrk = [[0.0] * 4 for _ in range(nX)]
    for n in range(1, nTS + 1):
        for i in range(1, 5):
            r = float(i - 1) / 3.0
            r = (float(n - 1) + r) / float(nTS)
            for k in range(nX):
                Qrk[k][i - 1] = Qo[k] + (Qn[k] - Qo[k]) * r
        for k in range(nX):
            f1 = (Qrk[k][0] - (X[k] - Pd[k]) / Rd[k]) / C[k]
            Xrk = X[k] + dtt * f1 / 3.0
            f2 = (Qrk[k][1] - (Xrk - Pd[k]) / Rd[k]) / C[k]
            Xrk = X[k] - dtt * f1 / 3.0 + dtt * f2
            f3 = (Qrk[k][2] - (Xrk - Pd[k]) / Rd[k]) / C[k]
            Xrk = X[k] + dtt * f1 - dtt * f2 + dtt * f3
            f4 = (Qrk[k][3] - (Xrk - Pd[k]) / Rd[k]) / C[k]
            X[k] = X[k] + (dtt / 8.0) * (f1 + 3.0 * (f2 + f3) + f4)
        tt = tt + dtt
    return np.array(X), np.array([X[k] + Qn[k] * Rp[k] for k in range(nX)])


def calc_der_cplbc(xo, Qo, Qn, Rp, C, Rd, Pd, dt, time):
    """CALCDERCPLBC, S/SETBC.f:1037-1123 (all faces of the Neumann group): returns (y, r) with
    r_i = (y_i(Qn_i + diff) - y_i(Qn)) / diff, diff = max(absTol, relTol * rms(Qo))"""
    nX = len(xo)
    _, y0 = rcr_integ_x(xo, Qo, Qn, Rp, C, Rd, Pd, dt, time)
    diff = 0.0
    for k in range(nX):
        diff = diff + Qo[k] * Qo[k]
    diff = np.sqrt(diff / float(nX))
    diff = 1e-8 if diff * 1e-5 < 1e-8 else diff * 1e-5
    r = np.zeros(nX)
    for k in range(nX):
        Q = [float(v) for v in Qn]
        Q[k] = Qn[k] + diff
        _, y = rcr_integ_x(xo, Qo, Q, Rp, C, Rd, Pd, dt, time)
        r[k] = (y[k] - y0[k]) / diff
    return y0, r
