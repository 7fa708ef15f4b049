#!/usr/bin/env python
"""Golden vectors from the REFERENCE'S OWN SOURCE TEXT (this container only: needs /root/reference).

The reference (Fortran + MPI) cannot be compiled in this image.  `oracle/refexec.py` translates its procedures to
Python/NumPy on the fly, statement by statement, and this script runs the hot path through them on small cases:

  element loop   SELECTELE, GETGIP, GETGNN, INITFSMSH, GETTHOODFS, CONSTRUCT_FLUID, GNN, GNNxx, FLUID3D_M, FLUID3D_C,
                 GETVISCOSITY, DOMAIN, ISZERO, DOASSEM                      (svFSI/NN.f, FS.f, FLUID.f, ALLFUN.f, UTIL.f, LHSA.f)
  linear solver  FSILS_LHS_CREATE, FSILS_BC_CREATE, FSILS_SOLVE, PRECONDDIAG, PRECONDRCS, GMRESV / GMRESS / GMRES,
                 CGRADV / CGRADS / SCHUR, BICGSV / BICGSS, NSSOLVER, DEPART, GE, ADDBCMUL, FSILS_SPARMUL*, FSILS_DOT*,
                 FSILS_NORM*, OMP* , FSILS_BCAST*                            (svFSILS/*.f, one task)

Nothing is copied: the sources are read where they lie.  Output: tests/golden/ref_*.npz (inputs + the reference's
results), which `tests/test_reference_golden.py` compares the oracle with (CPU) and `-m gpu` tests the CUDA path with.

    python tests/golden/make_ref_golden.py            # regenerate all
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refexec as rx  # noqa: E402

REF = os.environ.get("SVFSI_REFERENCE", "/root/reference")
S = os.path.join(REF, "Code", "Source", "svFSI")
LS = os.path.join(REF, "Code", "Source", "svFSILS")


# ------------------------------------------------------------------------------------------------ svFSI side
def svfsi_gen():
    lib = rx.Library()
    for h in ("FSILS_TYPEDEF.h", "FSILS_STRUCT.h"):          # eqType carries an FSILS_lsType
        lib.add_include(os.path.join(LS, h))
    for f in ("CONSTS.f", "TYPEMOD.f", "UTIL.f", "MOD.f", "ALLFUN.f", "NN.f", "FS.f", "FLUID.f", "HEATS.f", "LHSA.f",
              "EQASSEM.f", "PIC.f", "SETBC.f", "BAFINI.f", "OUTPUT.f"):
        lib.add_file(os.path.join(S, f))
    for f in sorted(os.listdir(LS)):                         # FSILSINI / LSSOLVE call into svFSILS on COMMOD's lhs
        if f.endswith(".f"):
            lib.add_file(os.path.join(LS, f))

    def dgesv(n, nrhs, a, lda, ipiv, b, ldb, info):
        """LAPACK DGESV (the reference links LAPACK): A <- LU, B <- solution, INFO handed back"""
        from scipy.linalg import lapack
        lu, piv, xsol, inf = lapack.dgesv(a[:n, :n], b[:n, :nrhs])
        a[:n, :n] = lu
        b[:n, :nrhs] = xsol
        ipiv[:n] = piv + 1
        return (int(inf),)

    gen = rx.CodeGen(lib, externals={"dgesv": dgesv, "fsils_cput": lambda: 0.0, "cput": lambda: 1.5,
                                     "str": lambda *a: str(a[0])})      # (number -> text, only ever used in messages / file names)
    gen.ext_outs["dgesv"] = [7]

    class Cm:                        # svFSI's communicator object (CMMOD, type-bound procedures), one task
        @staticmethod
        def seq():
            return True

        @staticmethod
        def reduce(v, *a):
            return v

        @staticmethod
        def mas():
            return True

        @staticmethod
        def tf():
            return 1

        @staticmethod
        def bcast(*a):
            return None
    gen.M.cm = Cm()
    gen.M.ikind, gen.M.rkind = 4, 8     # kind numbers (only ever passed as KIND= arguments)
    for n in ("mpint", "mpreal", "mplog", "mpchar", "mpi_sum", "mpi_max", "mpi_min", "stdout"):
        setattr(gen.M, n, 0)
    gen.M.mpsts = 6
    return gen


def element_loop(gen, x, IEN, rowPtr, colPtr, Ag, Yg, rho, mu, f, dt, af, am, gam, physics="fluid", heat=None, Bf=None):
    """CONSTRUCT_FLUID / CONSTRUCT_HEATS of the reference on one mesh; returns (R[nNo, dof], Val[nnz, dof*dof])"""
    M, rt = gen.M, gen.rt
    nNo, nEl = x.shape[0], IEN.shape[0]
    dof = 4 if physics == "fluid" else 1
    M.nsd, M.nsymd, M.tdof, M.dof, M.tnno = 3, 6, Ag.shape[1], dof, nNo
    M.x = np.asfortranarray(x.T.copy())
    M.bf = np.zeros((3, nNo), order="F") if Bf is None else np.asfortranarray(Bf.T.copy())
    M.ceq, M.cdmn, M.dt, M.mvmsh, M.nmsh = 1, 1, float(dt), False, 1
    eq = rt.new("eqtype")
    eq.af, eq.am, eq.gam = float(af), float(am), float(gam)
    eq.phys = M.phys_fluid if physics == "fluid" else M.phys_heats
    eq.ndmn = 1
    dmn = rt.new("dmntype")
    dmn.phys, dmn.id = eq.phys, -1
    dmn.prop = np.zeros(int(M.maxnprop), order="F")
    if physics == "fluid":
        dmn.prop[M.fluid_density - 1] = rho
        dmn.prop[M.f_x - 1], dmn.prop[M.f_y - 1], dmn.prop[M.f_z - 1] = f
        dmn.visc.visctype = M.visctype_const
        dmn.visc.mu_i = float(mu)
    else:
        dmn.prop[M.conductivity - 1] = heat["nu"]
        dmn.prop[M.source_term - 1] = heat["s"]
        dmn.prop[M.solid_density - 1] = heat["rho"]
    eq.dmn = rx.FList([dmn])
    eq.s, eq.e, eq.dof = 1, dof, dof
    M.eq = rx.FList([eq])
    # COMMOD's assembled system (S/LHSA.f DOASSEM works on these)
    M.rowptr = np.asfortranarray(rowPtr.astype(np.int64))
    M.colptr = np.asfortranarray(colPtr.astype(np.int64))
    M.r = np.zeros((dof, nNo), order="F")
    M.val = np.zeros((dof * dof, colPtr.size), order="F")
    # the mesh: SELECTELE fills the element tables, INITFSMSH the function space
    lM = rt.new("mshtype")
    lM.lshl, lM.lfib = False, False
    lM.enon, lM.nel, lM.gnel, lM.nno, lM.nfs = 4, nEl, nEl, nNo, 1
    lM.ien = np.asfortranarray(IEN.T.astype(np.int64))
    gen.get("selectele")(lM)
    gen.get("initfsmsh")(lM)
    M.msh = rx.FList([lM])
    Agf, Ygf = np.asfortranarray(Ag.T.copy()), np.asfortranarray(Yg.T.copy())
    gen.get("construct_fluid" if physics == "fluid" else "construct_heats")(lM, Agf, Ygf)
    tables = dict(w=np.array(lM.w), xi=np.array(lM.xi), N=np.array(lM.n), Nx=np.array(lM.nx))
    return M.r.T.copy(), M.val.T.copy(), tables       # (copies: the face terms are added to COMMOD's R / Val later)


def face_neumann(gen, gN, fIEN, gE, hg, Yg, bfStab):
    """SELECTELEB + IntegV + BASSEMNEUBC (-> GNNB, BFLUID, DOASSEM) on one face of msh(1); works on the R / Val that
    the element loop left in COMMOD.  Returns (flux of Yg(1:3), R, Val, face tables)"""
    M, rt = gen.M, gen.rt
    lM = M.msh[0]
    lFa = rt.new("facetype")
    lFa.im, lFa.enon, lFa.nel, lFa.nno = 1, 3, int(fIEN.shape[0]), int(gN.size)
    lFa.ien = np.asfortranarray(fIEN.T.astype(np.int64))
    lFa.ge = gE.astype(np.int64)
    lFa.gn = gN.astype(np.int64)
    gen.get("selecteleb")(lM, lFa)
    M.eq[0].dmn[0].prop[M.backflow_stab - 1] = bfStab
    M.ibflag = False
    Ygf = np.asfortranarray(Yg.T.copy())
    flux = gen.get("integv")(lFa, np.asfortranarray(Yg[:, :3].T.copy()))
    gen.get("bassemneubc")(lFa, np.asarray(hg, dtype=np.float64), Ygf)
    tab = dict(w=np.array(lFa.w), N=np.array(lFa.n), Nx=np.array(lFa.nx))
    return float(flux), M.r.T.copy(), M.val.T.copy(), tab


def pic_cycle(gen, Ao, Yo, Do, Rinc, gam, beta, am, af, dt, iNorm, tol, absTol, minItr, maxItr, itr0):
    """PICP, PICI, PICC (S/PIC.f) for one equation of 4 dofs; Rinc = what the linear solver left in R.
    Returns the arrays after each routine and PICC's bookkeeping."""
    M, rt = gen.M, gen.rt
    nNo = Ao.shape[0]
    F = lambda a: np.asfortranarray(a.T.copy())
    M.psteq, M.ibflag, M.eccpld, M.dflag, M.ssteq, M.cmminit, M.neq = False, False, False, False, False, False, 1
    M.ao, M.yo, M.do = F(Ao), F(Yo), F(Do)
    M.an, M.yn, M.dn = np.zeros_like(M.ao), np.zeros_like(M.ao), np.zeros_like(M.ao)
    eq = M.eq[0]
    eq.s, eq.e, eq.dof = 1, 4, 4
    eq.gam, eq.beta, eq.am, eq.af = float(gam), float(beta), float(am), float(af)
    eq.itr, eq.ok, eq.coupled = int(itr0), False, False
    eq.tol, eq.abstol, eq.minitr, eq.maxitr, eq.inorm, eq.pnorm = float(tol), float(absTol), int(minItr), int(maxItr), 0.0, 0.0
    M.dt, M.ceq, M.tnno, M.tdof, M.nsd = float(dt), 1, nNo, 4, 3
    out = {}
    gen.get("picp")()
    out["picp_An"], out["picp_Yn"], out["picp_Dn"] = (a.T.copy() for a in (M.an, M.yn, M.dn))
    Ag, Yg, Dg = (np.full((4, nNo), np.nan, order="F") for _ in range(3))
    gen.get("pici")(Ag, Yg, Dg)
    out["pici_Ag"], out["pici_Yg"], out["pici_Dg"] = (a.T.copy() for a in (Ag, Yg, Dg))
    M.r = F(Rinc)
    eq.fsils.ri.inorm = float(iNorm)
    gen.get("picc")()
    out["picc_An"], out["picc_Yn"], out["picc_Dn"] = (a.T.copy() for a in (M.an, M.yn, M.dn))
    out["picc_book"] = np.array([eq.itr, float(eq.ok), eq.inorm, eq.pnorm, M.ceq], dtype=np.float64)
    return out


# ------------------------------------------------------------------------------------------------ FSILS side
def fsils_gen():
    lib = rx.Library()
    for h in ("FSILS_TYPEDEF.h", "FSILS_STRUCT.h"):
        lib.add_include(os.path.join(LS, h))
    for f in sorted(os.listdir(LS)):
        if f.endswith(".f"):
            lib.add_file(os.path.join(LS, f))

    def cput():
        return 0.0

    gen = rx.CodeGen(lib, externals={"fsils_cput": cput, "cpu_time": lambda t: (0.0,)})
    gen.ext_outs["cpu_time"] = [0]
    # mpif.h is not part of the reference tree: the handles are never used with one task, only their names appear
    for n in ("mpint", "mpreal", "mplog", "mpchar", "mpi_sum", "mpi_max", "mpi_min", "mpi_lor", "mpi_comm_world",
              "mpi_integer", "mpi_double_precision", "mpi_logical", "mpi_character", "mpi_in_place", "stdout"):
        setattr(gen.M, n, 0)
    gen.M.mpsts = gen.M.mpi_status_size = 6
    gen.M.lsip, gen.M.lsrp = 4, 8       # kind numbers (only ever passed as KIND= arguments)
    return gen


class MpiEmu:
    """The MPI calls of svFSILS for `n` tasks that run as THREADS of this process (threading.current_thread().rank):
    ALLREDUCE (rank-ordered sum / max), ALLGATHER(V), SEND / RECV, ISEND / IRECV / WAIT.  Scalars that MPI would
    store through a reference are handed back (refexec's convention); arrays are filled in place."""

    def __init__(self, n):
        import collections
        import queue
        import threading
        self.n = n
        self.bar = threading.Barrier(n)
        self.slots = [None] * n
        self.mail = collections.defaultdict(queue.Queue)
        self.reqs = {}
        self.lock = threading.Lock()
        self.nreq = 0
        self.threading = threading

    def rank(self):
        return self.threading.current_thread().rank

    @staticmethod
    def _flat(a, count):
        return a.reshape(-1, order="F")[:int(count)] if isinstance(a, np.ndarray) else a

    def _exchange(self, val):
        r = self.rank()
        self.slots[r] = np.array(val, copy=True) if isinstance(val, np.ndarray) else val
        self.bar.wait()
        allv = list(self.slots)
        self.bar.wait()
        return allv

    def allreduce(self, send, recv, count, dtype, op, comm, ierr):
        allv = self._exchange(self._flat(send, count))
        res = allv[0]
        for v in allv[1:]:                       # rank order
            res = np.maximum(res, v) if op == "max" else res + v
        if isinstance(recv, np.ndarray) and recv.ndim > 0:
            self._flat(recv, count)[...] = res
            return recv, 0
        if isinstance(res, np.ndarray):
            res = res.reshape(-1)[0]
        return (int(res) if isinstance(send, (int, np.integer)) and not isinstance(send, bool) else res), 0

    def allgatherv(self, send, scount, stype, recv, rcounts, displs, rtype, comm, ierr):
        allv = self._exchange(self._flat(send, scount))
        flat = recv.reshape(-1, order="F")
        assert np.shares_memory(flat, recv)
        for r, v in enumerate(allv):
            flat[int(displs[r]):int(displs[r]) + int(rcounts[r])] = v[:int(rcounts[r])]
        return (0,)

    def allgather(self, send, scount, stype, recv, rcount, rtype, comm, ierr):
        allv = self._exchange(send)
        for r, v in enumerate(allv):
            recv[r] = v
        return (0,)

    def send(self, buf, count, dtype, dest, tag, comm, *rest):
        self.mail[(self.rank(), int(dest), int(tag))].put(np.array(self._flat(buf, count), copy=True))

    def recv(self, buf, count, dtype, src, tag, comm, stat, ierr):
        self._flat(buf, count)[...] = self.mail[(int(src), self.rank(), int(tag))].get(timeout=60)
        return (0,)

    def isend(self, buf, count, dtype, dest, tag, comm, req, ierr):
        self.send(buf, count, dtype, dest, tag, comm)
        return -1, 0

    def irecv(self, buf, count, dtype, src, tag, comm, req, ierr):
        with self.lock:
            self.nreq += 1
            rid = self.nreq
            self.reqs[rid] = (buf, int(count), int(src), int(tag), self.rank())
        return rid, 0

    def wait(self, req, stat, ierr):
        if int(req) > 0:
            buf, count, src, tag, me = self.reqs.pop(int(req))
            flat = buf.reshape(-1, order="F")
            assert np.shares_memory(flat, buf)
            flat[:count] = self.mail[(src, me, tag)].get(timeout=60)
        return (0,)

    def install(self, gen):
        M = gen.M
        M.mpi_sum, M.mpi_max = "sum", "max"
        ext = {"mpi_allreduce": (self.allreduce, [1, 6]), "mpi_allgatherv": (self.allgatherv, [8]),
               "mpi_allgather": (self.allgather, [7]), "mpi_send": (self.send, []), "mpi_recv": (self.recv, [7]),
               "mpi_isend": (self.isend, [6, 7]), "mpi_irecv": (self.irecv, [6, 7]), "mpi_wait": (self.wait, [2])}
        for k, (fn, outs) in ext.items():
            gen.externals[k] = fn
            gen.ext_outs[k] = outs

    def run(self, fn):
        """fn(rank) on every task; returns the list of results (exceptions are re-raised)"""
        res, errs = [None] * self.n, []

        def body(r):
            try:
                res[r] = fn(r)
            except BaseException as ex:      # noqa: BLE001
                errs.append((r, ex))
                self.bar.abort()
        ths = []
        for r in range(self.n):
            t = self.threading.Thread(target=body, args=(r,))
            t.rank = r
            ths.append(t)
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0][1]
        return res


def fsils_multitask(probs, gnNo, Rs, Vs, cases, face_order):
    """FSILS on len(probs) emulated MPI tasks: FSILS_LHS_CREATE (reordering + communication lists), FSILS_BC_CREATE
    (shared faces), FSILS_COMMUV on R, FSILS_SOLVE.  Returns per case and task the solution in svFSI's local order."""
    n = len(probs)
    fg = fsils_gen()
    emu = MpiEmu(n)
    emu.install(fg)
    FM, rt = fg.M, fg.rt
    # translate everything once, single-threaded (the translator itself is not re-entrant)
    for name in ("fsils_lhs_create", "fsils_bc_create", "fsils_commuv", "fsils_commus", "fsils_ls_create", "fsils_solve"):
        fg.get(name)

    def make_lhs(r):
        p = probs[r]
        commu = rt.new("fsils_commutype")
        commu.foc, commu.masf, commu.master, commu.task, commu.tf, commu.ntasks, commu.comm = True, r == 0, 0, r, r + 1, n, 0
        lhs = rt.new("fsils_lhstype")
        fg.get("fsils_lhs_create")(lhs, commu, int(gnNo), int(p.rm.nNo), int(p.colPtr.size), p.rm.ltg.astype(np.int64),
                                   p.rowPtr.astype(np.int64), p.colPtr.astype(np.int64), len(face_order))
        for fi, fname in enumerate(face_order, start=1):
            fa = p.faces[fname]
            bc = FM.bc_type_neu if fa["bc"] == "Neu" else FM.bc_type_dir
            v = None if fa["val"] is None else np.asfortranarray(np.asarray(fa["val"], dtype=np.float64).reshape(-1, 3).T.copy())
            fg.get("fsils_bc_create")(lhs, fi, int(fa["gN"].size), 3, int(bc), fa["gN"].astype(np.int64), v)
        return lhs

    def task(r):
        out = {}
        lhs = make_lhs(r)
        mp = np.array(lhs.map) - 1
        out["map"], out["mynNo"], out["shnNo"], out["nReq"] = np.array(lhs.map), lhs.mynno, lhs.shnno, lhs.nreq
        out["cS_iP"] = np.array([c.ip for c in lhs.cs], dtype=np.int64)
        out["cS_ptr"] = np.concatenate([np.array(c.ptr) for c in lhs.cs]) if lhs.cs else np.zeros(0, dtype=np.int64)
        out["cS_n"] = np.array([c.n for c in lhs.cs], dtype=np.int64)
        # COMMU(R) as S/ALLFUN.f:514-533 does it: permute, FSILS_COMMUV, unpermute
        tmp = np.zeros((4, probs[r].rm.nNo), order="F")
        tmp[:, mp] = Rs[r].T
        fg.get("fsils_commuv")(lhs, 4, tmp)
        Rc = tmp[:, mp].T.copy()
        out["Rc"] = Rc
        for name, lst, prec, kw, res_out in cases:
            lhs = make_lhs(r)
            X, _, cnt = fsils_solve(fg, lhs, lst, 4, Rc, Vs[r], prec, incL=[1, 1, 1], res=[0.0, 0.0, res_out], **kw)
            out[name + "_X"] = X
            out[name + "_cnt"] = np.array([cnt[k] for k in sorted(cnt)], dtype=np.float64)
            out["cnt_keys"] = np.array(sorted(cnt))
        return out
    MT_GEN[0] = fg
    return emu.run(task), FM


def svfsi_commu_multitask(probs, gnNo, Rs):
    """COMMU(R) exactly as svFSI does it (S/ALLFUN.f:514-533: U = MKC(U); FSILS_COMMUV; MKCI(U)) on len(probs) emulated
    MPI tasks, each with its own COMMOD (one translator per task: COMMOD is process-global state in the reference)"""
    n = len(probs)
    emu = MpiEmu(n)
    gens = []
    for r in range(n):
        g = svfsi_gen()
        emu.install(g)

        class Cm:
            @staticmethod
            def seq():
                return False
        g.M.cm = Cm()
        for name in ("fsils_lhs_create", "commuv"):
            g.get(name)
        gens.append(g)

    def task(r):
        g, p = gens[r], probs[r]
        M, rt = g.M, g.rt
        commu = rt.new("fsils_commutype")
        commu.foc, commu.masf, commu.master, commu.task, commu.tf, commu.ntasks, commu.comm = True, r == 0, 0, r, r + 1, n, 0
        M.lhs = rt.new("fsils_lhstype")
        g.get("fsils_lhs_create")(M.lhs, commu, int(gnNo), int(p.rm.nNo), int(p.colPtr.size), p.rm.ltg.astype(np.int64),
                                  p.rowPtr.astype(np.int64), p.colPtr.astype(np.int64), 0)
        U = np.asfortranarray(Rs[r].T.copy())
        g.get("commuv")(U)
        return U.T.copy()
    return emu.run(task)


MT_GEN = [None]


def fsils_lhs(gen, gnNo, rowPtr, colPtr, faces):
    """FSILS_COMMU (one task, built by hand: FSILS_COMMU_CREATE only wraps MPI calls), FSILS_LHS_CREATE,
    FSILS_BC_CREATE.  faces: list of (gN 1-based, dof, bc_type, val[nNo_face, dof] | None)"""
    rt, M = gen.rt, gen.M
    commu = rt.new("fsils_commutype")
    commu.foc, commu.masf, commu.master, commu.task, commu.tf, commu.ntasks, commu.comm = True, True, 0, 0, 1, 1, 0
    lhs = rt.new("fsils_lhstype")
    nNo = rowPtr.size - 1
    gN = np.arange(1, nNo + 1, dtype=np.int64)
    gen.get("fsils_lhs_create")(lhs, commu, int(gnNo), nNo, int(colPtr.size), gN, rowPtr.astype(np.int64),
                                colPtr.astype(np.int64), len(faces))
    for fi, (g, dof, bc, val) in enumerate(faces, start=1):
        v = None if val is None else np.asfortranarray(np.asarray(val, dtype=np.float64).T.copy())
        gen.get("fsils_bc_create")(lhs, fi, int(g.size), int(dof), int(bc), g.astype(np.int64), v)
    return lhs


def fsils_solve(gen, lhs, ls_type, dof, R, Val, prec, incL=None, res=None, **lskw):
    """FSILS_LS_CREATE + FSILS_SOLVE; returns (X[nNo, dof], scaled Val[nnz, dof*dof], counters)"""
    rt, M = gen.rt, gen.M
    ls = rt.new("fsils_lstype")
    kw = {k.lower(): v for k, v in lskw.items() if v is not None}
    for k in ("reltolin", "abstolin"):
        if k in kw:
            kw[k] = np.array(kw[k], dtype=np.float64)
    if "maxitrin" in kw:
        kw["maxitrin"] = np.array(kw["maxitrin"], dtype=np.int64)
    gen.get("fsils_ls_create")(ls, int(ls_type), **kw)
    Ri = np.asfortranarray(R.reshape(R.shape[0], -1).T.copy())
    V = np.asfortranarray(Val.reshape(Val.shape[0], -1).T.copy())
    gen.get("fsils_solve")(lhs, ls, int(dof), Ri, V, int(prec),
                           None if incL is None else np.asarray(incL, dtype=np.int64),
                           None if res is None else np.asarray(res, dtype=np.float64))
    cnt = {}
    for sub in ("ri", "gm", "cg"):
        o = getattr(ls, sub)
        for f in ("itr", "suc", "inorm", "fnorm", "db"):
            v = getattr(o, f)
            cnt[f"{sub}_{f}"] = float(v) if not isinstance(v, (bool, np.bool_)) else bool(v)
    return Ri.T.copy(), V.T.copy(), cnt


def main():
    import common as cm
    import unstructured as un
    from svfsi_b200 import mesh
    t0 = time.time()
    gen = svfsi_gen()
    # ---- case A: 2x2x3 Kuhn-lattice pipe (the mesh family of the parity tests)
    m, probs, _ = mesh.build_problem(2, 2, 3, nparts=1, L=1.5)
    p = probs[0]
    R, V, tab = element_loop(gen, p.rm.x, p.rm.IEN, p.rowPtr, p.colPtr, p.Ag, p.Yg, cm.RHO, cm.MU, cm.F, cm.DT,
                             cm.GA["af"], cm.GA["am"], cm.GA["gam"])
    print(f"lattice: nEl={p.rm.IEN.shape[0]} |R|={np.abs(R).max():.3e} |Val|={np.abs(V).max():.3e} {time.time()-t0:.1f}s")
    np.savez_compressed(os.path.join(HERE, "ref_fluid_lattice.npz"), x=p.rm.x, IEN=p.rm.IEN, rowPtr=p.rowPtr,
                        colPtr=p.colPtr, Ag=p.Ag, Yg=p.Yg, rho=cm.RHO, mu=cm.MU, f=np.array(cm.F), dt=cm.DT,
                        af=cm.GA["af"], am=cm.GA["am"], gam=cm.GA["gam"], R=R, Val=V, **{"tab_" + k: v for k, v in tab.items()})
    # ---- FSILS on that system: the three faces of the parity tests, every solver / preconditioner
    fg = fsils_gen()
    FM = fg.M
    faces = []
    for name in cm.FACE_ORDER:
        fa = p.faces[name]
        faces.append((fa["gN"], 3, FM.bc_type_neu if fa["bc"] == "Neu" else FM.bc_type_dir, fa["val"]))
    cases = [
        ("gmres_diag", FM.ls_type_gmres, FM.precond_fsils, dict(relTol=1e-6, absTol=1e-14, maxItr=10, dimKry=30), 0.0),
        ("gmres_diag_res", FM.ls_type_gmres, FM.precond_fsils, dict(relTol=1e-6, absTol=1e-14, maxItr=10, dimKry=30), 0.7),
        ("gmres_restart", FM.ls_type_gmres, FM.precond_fsils, dict(relTol=1e-8, absTol=1e-14, maxItr=20, dimKry=8), 0.0),
        # PRECONDRCS with a COUPLED face is undefined in the reference: the valM update is commented out
        # (L/PRECOND.f:352-362), so ADDBCMUL reads the never-initialised valM of FSILS_BC_CREATE (NaN here: refexec
        # poisons uninitialised REALs).  The RCS cases therefore run without resistance.
        ("gmres_rcs", FM.ls_type_gmres, FM.precond_rcs, dict(relTol=1e-6, absTol=1e-14, maxItr=10, dimKry=30), 0.0),
        ("ns_default", FM.ls_type_ns, FM.precond_fsils, dict(), 0.0),
        ("ns_tight_res", FM.ls_type_ns, FM.precond_fsils, dict(relTol=1e-4, absTol=1e-14, maxItr=10, dimKry=40,
                                                                relTolIn=(1e-3, 1e-2), maxItrIn=(3, 100)), 3.0),
        ("bicgs_diag", FM.ls_type_bicgs, FM.precond_fsils, dict(relTol=1e-6, absTol=1e-14, maxItr=300), 0.0),
        ("bicgs_rcs_res", FM.ls_type_bicgs, FM.precond_rcs, dict(relTol=1e-5, absTol=1e-14, maxItr=300), 0.7),
    ]
    sol = {}
    for name, lst, prec, kw, res_out in cases:
        lhs = fsils_lhs(fg, p.rm.nNo, p.rowPtr, p.colPtr, faces)
        t1 = time.time()
        X, Vs, cnt = fsils_solve(fg, lhs, lst, 4, R, V, prec, incL=[1, 1, 1], res=[0.0, 0.0, res_out], **kw)
        print(f"  FSILS {name}: RI itr={cnt['ri_itr']:.0f} suc={cnt['ri_suc']} iNorm={cnt['ri_inorm']:.6e} "
              f"fNorm={cnt['ri_fnorm']:.3e} GM={cnt['gm_itr']:.0f} CG={cnt['cg_itr']:.0f} {time.time()-t1:.1f}s")
        sol[name + "_X"] = X
        sol[name + "_Vscaled"] = Vs
        sol[name + "_cnt"] = np.array([cnt[k] for k in sorted(cnt)], dtype=np.float64)
        sol[name + "_kw"] = np.array(repr(dict(ls_type=int(lst), prec=int(prec), res_out=res_out, **kw)))
    sol["cnt_keys"] = np.array(sorted(cnt))
    # FSILS_LS_CREATE defaults (L/LS.f:69-95) of every solver type: relTol, absTol, mItr, sD of RI / GM / CG
    for tname in ("ns", "gmres", "cg", "bicgs"):
        ls = fg.rt.new("fsils_lstype")
        for sub in ("ri", "gm", "cg"):          # poison what FSILS_LS_CREATE does not set, so that it shows
            o = getattr(ls, sub)
            o.reltol = o.abstol = float("nan"); o.mitr = o.sd = -1
        fg.get("fsils_ls_create")(ls, int(getattr(FM, "ls_type_" + tname)))
        sol["lsdef_" + tname] = np.array([[getattr(ls, sub).reltol, getattr(ls, sub).abstol, getattr(ls, sub).mitr,
                                           getattr(ls, sub).sd] for sub in ("ri", "gm", "cg")], dtype=np.float64)
        sol["lstype_" + tname] = int(getattr(FM, "ls_type_" + tname))
    np.savez_compressed(os.path.join(HERE, "ref_fsils_lattice.npz"), **sol)

    # ---- LHSA (S/LHSA.f:40-264): the block-CSR pattern of the mesh the element loop just ran on
    M = gen.M
    M.shleq, M.neq = False, 1
    M.eq[0].nbc = 0
    (nnz,) = gen.get("lhsa")(0)
    print(f"  LHSA: nnz={nnz} (harness pattern: {p.colPtr.size})")
    np.savez_compressed(os.path.join(HERE, "ref_lhsa_lattice.npz"), IEN=p.rm.IEN, nNo=p.rm.nNo, nnz=nnz,
                        rowPtr=np.array(M.rowptr), colPtr=np.array(M.colptr))

    # ---- face terms on the lattice: outlet with the flow reversed (backflow stabilisation active), inlet as it is
    fc = {}
    for fname, flow, h in (("outlet", -1.0, 3.5), ("inlet", 1.0, -2.0), ("outlet", 1.0, 0.7)):
        gN, fIEN, gE = cm.local_face(m, p.rm, fname)
        Yg2 = p.Yg.copy(); Yg2[:, :3] *= flow
        R0, V0, _ = element_loop(gen, p.rm.x, p.rm.IEN, p.rowPtr, p.colPtr, p.Ag, Yg2, cm.RHO, cm.MU, cm.F, cm.DT,
                                 cm.GA["af"], cm.GA["am"], cm.GA["gam"])
        hg = np.zeros(p.rm.nNo); hg[gN - 1] = -h
        flux, R1, V1, ftab = face_neumann(gen, gN, fIEN, gE, hg, Yg2, 0.2)
        key = f"{fname}_{'rev' if flow < 0 else 'fwd'}"
        print(f"  face {key}: nEl={fIEN.shape[0]} flux={flux:.6e} |dR|={np.abs(R1 - R0).max():.3e} |dVal|={np.abs(V1 - V0).max():.3e}")
        fc.update({f"{key}_gN": gN, f"{key}_fIEN": fIEN, f"{key}_gE": gE, f"{key}_hg": hg, f"{key}_Yg": Yg2,
                   f"{key}_R0": R0, f"{key}_V0": V0, f"{key}_R1": R1, f"{key}_V1": V1, f"{key}_flux": flux})
    # resistance outlet as svFSI applies it: SETBCNEUL (S/SETBC.f:251-311) -> h = r * Integ(lFa, Yn, 1, 3), hg = -h gx,
    # BASSEMNEUBC -- on top of a fresh element loop
    gN, fIEN, gE = cm.local_face(m, p.rm, "outlet")
    R0, V0, _ = element_loop(gen, p.rm.x, p.rm.IEN, p.rowPtr, p.colPtr, p.Ag, p.Yg, cm.RHO, cm.MU, cm.F, cm.DT,
                             cm.GA["af"], cm.GA["am"], cm.GA["gam"])
    M, rt = gen.M, gen.rt
    lFa = rt.new("facetype")
    lFa.im, lFa.enon, lFa.nel, lFa.nno = 1, 3, int(fIEN.shape[0]), int(gN.size)
    lFa.ien, lFa.ge, lFa.gn = np.asfortranarray(fIEN.T.astype(np.int64)), gE.astype(np.int64), gN.astype(np.int64)
    gen.get("selecteleb")(M.msh[0], lFa)
    lBc = rt.new("bctype")
    lBc.btype = (1 << M.btype_neu) | (1 << M.btype_res)
    lBc.r, lBc.g, lBc.flwp = 55.0, 0.0, False
    lBc.gx = np.ones(gN.size)
    M.eq[0].dmn[0].prop[M.backflow_stab - 1] = 0.2
    M.ibflag = False
    Ynf = np.asfortranarray(p.Yg.T.copy())
    M.yn = Ynf                                     # the flux is taken from Yn (here: the same state)
    gen.get("setbcneul")(lBc, lFa, Ynf, np.zeros_like(Ynf))
    fc.update(res_r=55.0, res_gN=gN, res_fIEN=fIEN, res_gE=gE, res_Yg=p.Yg, res_R0=R0, res_V0=V0, res_R1=M.r.T.copy(),
              res_V1=M.val.T.copy())
    print(f"  SETBCNEUL resistance: |dR|={np.abs(M.r.T - R0).max():.3e}")
    fc.update({"tab_" + k: v for k, v in ftab.items()})
    fc.update(bfStab=0.2, rho=cm.RHO, af=cm.GA["af"], gam=cm.GA["gam"], dt=cm.DT)
    np.savez_compressed(os.path.join(HERE, "ref_face_lattice.npz"), **fc)

    # ---- FSILSINI (S/BAFINI.f:468-583): what svFSI hands to FSILS_BC_CREATE for a Dirichlet face (zero mask) and for a
    # resistance face (val = int N n dGamma), on COMMOD's lhs
    M, rt = gen.M, gen.rt
    commu = rt.new("fsils_commutype")
    commu.foc, commu.masf, commu.master, commu.task, commu.tf, commu.ntasks, commu.comm = True, True, 0, 0, 1, 1, 0
    M.lhs = rt.new("fsils_lhstype")
    gen.get("fsils_lhs_create")(M.lhs, commu, int(p.rm.nNo), int(p.rm.nNo), int(p.colPtr.size),
                                np.arange(1, p.rm.nNo + 1, dtype=np.int64), p.rowPtr.astype(np.int64),
                                p.colPtr.astype(np.int64), 3)
    ini, lsPtr = {}, 0
    for fname in cm.FACE_ORDER:
        gN, fIEN, gE = cm.local_face(m, p.rm, fname)
        lFa = rt.new("facetype")
        lFa.im, lFa.enon, lFa.nel, lFa.nno = 1, 3, int(fIEN.shape[0]), int(gN.size)
        lFa.ien, lFa.ge, lFa.gn = np.asfortranarray(fIEN.T.astype(np.int64)), gE.astype(np.int64), gN.astype(np.int64)
        gen.get("selecteleb")(M.msh[0], lFa)
        lBc = rt.new("bctype")
        lBc.weakdir = False
        lBc.edrn = np.zeros(int(M.maxnsd), dtype=np.int64)
        lBc.btype = ((1 << M.btype_neu) | (1 << M.btype_res)) if fname == "outlet" else ((1 << M.btype_dir) | (1 << M.btype_std))
        (lsPtr,) = gen.get("fsilsini")(lBc, lFa, lsPtr)
        f = M.lhs.face[lsPtr - 1]
        ini.update({f"{fname}_lsPtr": lsPtr, f"{fname}_bGrp": f.bgrp, f"{fname}_glob": np.array(f.glob),
                    f"{fname}_val": np.array(f.val).T.copy(), f"{fname}_gN": gN})
        print(f"  FSILSINI {fname}: lsPtr={lsPtr} bGrp={f.bgrp} nNo={f.nno} |val|={np.abs(f.val).max():.4e}")
    np.savez_compressed(os.path.join(HERE, "ref_fsilsini_lattice.npz"), **ini)

    # ---- the Newton / time loop of S/MAIN.f:111-206 composed from the reference's own routines: two time steps of three
    # Newton iterations on the pipe (steady parabolic Dirichlet inlet along the normal, no-slip wall, resistance outlet
    # coupled through res = gam dt r).  The loop below restates MAIN.f's ORDER of calls and its three bookkeeping lines
    # (R = 0 / Val = 0 of LSALLOC, incL / res of :186-192, Ao = An / Yo = Yn of :277-279); everything else is executed
    # from source: PICP, SETBCDIR(L), PICI, CONSTRUCT_FLUID, SETBCNEUL, FSILS_SOLVE (LSSOLVE's call), PICC.
    ga = cm.GA
    nN = p.rm.nNo
    rngl = np.random.default_rng(21)
    Ao = 0.05 * rngl.standard_normal((nN, 4)); Ao[:, 3] = 0.0
    Yo = p.Yg.copy()
    resist = 60.0
    gin, gw = p.faces["inlet"]["gN"], p.faces["wall"]["gN"]
    xin = p.rm.x[gin - 1]
    gx_in = np.clip(1.0 - (xin[:, 0] ** 2 + xin[:, 1] ** 2) / (np.abs(p.rm.x[:, :2]).max() ** 2), 0.0, None)
    gout, fIENo, gEo = cm.local_face(m, p.rm, "outlet")
    # state of COMMOD for the loop (the element-loop call below also (re)creates eq, msh, x, rowPtr, ...)
    element_loop(gen, p.rm.x, p.rm.IEN, p.rowPtr, p.colPtr, p.Ag, p.Yg, cm.RHO, cm.MU, cm.F, cm.DT, ga["af"], ga["am"], ga["gam"])
    eq = M.eq[0]
    faces_l, bcs_l = [], []
    for k, (fname, g, gxv) in enumerate((("inlet", -12.0, gx_in), ("wall", 0.0, np.ones(gw.size))), start=1):
        gNf = p.faces[fname]["gN"]
        fa = rt.new("facetype")
        fa.nno, fa.gn = int(gNf.size), gNf.astype(np.int64)
        fa.nv = np.asfortranarray(np.tile(np.array([[0.0], [0.0], [-1.0]]), (1, gNf.size)))
        bc = rt.new("bctype")
        bc.btype = (1 << M.btype_dir) | (1 << M.btype_std)
        bc.weakdir, bc.ifa, bc.im, bc.g, bc.gx = False, k, 1, g, np.array(gxv, dtype=np.float64)
        bc.edrn = np.zeros(int(M.maxnsd), dtype=np.int64)
        faces_l.append(fa); bcs_l.append(bc)
    lFo = rt.new("facetype")
    lFo.im, lFo.enon, lFo.nel, lFo.nno = 1, 3, int(fIENo.shape[0]), int(gout.size)
    lFo.ien, lFo.ge, lFo.gn = np.asfortranarray(fIENo.T.astype(np.int64)), gEo.astype(np.int64), gout.astype(np.int64)
    gen.get("selecteleb")(M.msh[0], lFo)
    bco = rt.new("bctype")
    bco.btype = (1 << M.btype_neu) | (1 << M.btype_res)
    bco.r, bco.g, bco.flwp, bco.gx, bco.ifa, bco.im = resist, 0.0, False, np.ones(gout.size), 3, 1
    faces_l.append(lFo); bcs_l.append(bco)
    M.msh[0].fa = rx.FList(faces_l)
    eq.bc, eq.nbc = rx.FList(bcs_l), 3
    eq.s, eq.e, eq.dof, eq.phys, eq.coupled = 1, 4, 4, M.phys_fluid, False
    eq.gam, eq.beta, eq.am, eq.af = float(ga["gam"]), float(ga["beta"]), float(ga["am"]), float(ga["af"])
    eq.tol, eq.abstol, eq.minitr, eq.maxitr = 1e-30, 1e-30, 1, 3
    eq.dmn[0].prop[M.backflow_stab - 1] = 0.2
    M.psteq, M.ibflag, M.eccpld, M.dflag, M.ssteq, M.cmminit, M.neq, M.ceq = False, False, False, False, False, False, 1, 1
    # COMMOD's lhs with its three faces (FSILS_LHS_CREATE + FSILSINI), the equation's linear solver (FSILS_LS_CREATE)
    commu = rt.new("fsils_commutype")
    commu.foc, commu.masf, commu.master, commu.task, commu.tf, commu.ntasks, commu.comm = True, True, 0, 0, 1, 1, 0
    M.lhs = rt.new("fsils_lhstype")
    gen.get("fsils_lhs_create")(M.lhs, commu, nN, nN, int(p.colPtr.size), np.arange(1, nN + 1, dtype=np.int64),
                                p.rowPtr.astype(np.int64), p.colPtr.astype(np.int64), 3)
    lsPtr = 0
    for bc, fa in zip(bcs_l[:2], faces_l[:2]):
        fa.im, fa.enon, fa.nel = 1, 3, 0
    for bc, fa in zip(bcs_l, faces_l):
        (lsPtr,) = gen.get("fsilsini")(bc, fa, lsPtr)
    gen.get("fsils_ls_create")(eq.fsils, int(M.ls_type_gmres), 1e-5, 1e-14, 10, 80)
    F = lambda a: np.asfortranarray(a.T.copy())
    M.ao, M.yo, M.do = F(Ao), F(Yo), np.zeros((4, nN), order="F")
    M.an, M.yn, M.dn = np.zeros_like(M.ao), np.zeros_like(M.ao), np.zeros_like(M.ao)
    M.nfacesls = 3
    lp = dict(Ao=Ao, Yo=Yo, resist=resist, gx_in=gx_in, inlet_g=-12.0, relTol=1e-5, absTol=1e-14, maxItr=10, dimKry=80)
    norms, fluxes = [], []
    for ts in range(2):
        eq.itr, eq.ok, eq.inorm = 0, False, 0.0                                   # S/MAIN.f:99-100
        gen.get("picp")()
        gen.get("setbcdir")(M.an, M.yn, M.dn)
        for it in range(3):
            Ag, Yg, Dg = (np.full((4, nN), np.nan, order="F") for _ in range(3))
            gen.get("pici")(Ag, Yg, Dg)
            M.r = np.zeros((4, nN), order="F")                                    # LSALLOC, S/LS.f:44-51
            M.val = np.zeros((16, p.colPtr.size), order="F")
            gen.get("construct_fluid")(M.msh[0], Ag, Yg)
            fluxes.append(float(gen.get("integv")(lFo, np.asfortranarray(M.yn[:3, :]))))
            gen.get("setbcneul")(bco, lFo, Yg, Dg)
            incL = np.zeros(3, dtype=np.int64); resv = np.zeros(3)
            for bc in bcs_l:                                                      # S/MAIN.f:186-192
                if bc.lsptr != 0:
                    resv[bc.lsptr - 1] = eq.gam * M.dt * bc.r
                    incL[bc.lsptr - 1] = 1
            gen.get("fsils_solve")(M.lhs, eq.fsils, 4, M.r, M.val, int(M.precond_fsils), incL, resv)   # LSSOLVE, S/LS.f:92
            norms.append((eq.fsils.ri.inorm, eq.fsils.ri.itr))
            gen.get("picc")()
        M.ao, M.yo = M.an.copy(order="F"), M.yn.copy(order="F")                   # S/MAIN.f:277-279
    lp.update(norms=np.array(norms), fluxes=np.array(fluxes), An=M.an.T.copy(), Yn=M.yn.T.copy())
    print(f"  time loop: iNorm/itr per Newton iteration = {[(float(f'{a:.4e}'), int(b)) for a, b in norms]}")
    np.savez_compressed(os.path.join(HERE, "ref_timeloop_lattice.npz"), **lp)
    eq.nbc = 0

    # ---- generalised-alpha predictor / initiator / corrector (S/PIC.f) on random states
    rng = np.random.default_rng(17)
    nN = p.rm.nNo
    Ao, Yo, Do, Rinc = (rng.standard_normal((nN, 4)) for _ in range(4))
    pc = dict(Ao=Ao, Yo=Yo, Do=Do, Rinc=Rinc, gam=cm.GA["gam"], beta=cm.GA["beta"], am=cm.GA["am"], af=cm.GA["af"], dt=cm.DT)
    for tag, iNorm, tol, itr0 in (("first", 12.5, 1e-3, 0), ("conv", 1e-9, 1e-3, 2)):
        o = pic_cycle(gen, Ao, Yo, Do, Rinc, cm.GA["gam"], cm.GA["beta"], cm.GA["am"], cm.GA["af"], cm.DT, iNorm, tol,
                      1e-8, 1, 10, itr0)
        print(f"  PIC {tag}: book(itr, ok, iNorm, pNorm, cEq)={o['picc_book']}")
        pc.update({f"{tag}_{k}": v for k, v in o.items()})
        pc[f"{tag}_in"] = np.array([iNorm, tol, 1e-8, 1, 10, itr0])
    # strong Dirichlet: SETBCDIR + SETBCDIRL (S/SETBC.f:39-228), a steady parabolic inlet along the face normals and a
    # no-slip wall, on the predictor's An / Yn
    M, rt = gen.M, gen.rt
    faces, bcs = [], []
    for k, (fname, g) in enumerate((("inlet", -7.5), ("wall", 0.0)), start=1):
        gN = p.faces[fname]["gN"]
        fa = rt.new("facetype")
        fa.nno, fa.gn = int(gN.size), gN.astype(np.int64)
        nV = rng.standard_normal((3, gN.size)); nV /= np.linalg.norm(nV, axis=0)
        fa.nv = np.asfortranarray(nV)
        bc = rt.new("bctype")
        bc.btype = (1 << M.btype_dir) | (1 << M.btype_std)
        bc.weakdir, bc.ifa, bc.im, bc.g = False, k, 1, g
        bc.edrn = np.zeros(int(M.maxnsd), dtype=np.int64)
        bc.gx = rng.uniform(0.0, 1.0, gN.size)
        faces.append(fa); bcs.append(bc)
        pc.update({f"dir{k}_gN": gN, f"dir{k}_nV": nV.T.copy(), f"dir{k}_gx": np.array(bc.gx), f"dir{k}_g": g})
    M.msh[0].fa = rx.FList(faces)
    M.eq[0].bc, M.eq[0].nbc, M.neq = rx.FList(bcs), 2, 1
    M.eq[0].s, M.eq[0].e, M.eq[0].dof, M.eq[0].phys = 1, 4, 4, M.phys_fluid
    lA, lY, lD = (np.asfortranarray(rng.standard_normal((4, nN))) for _ in range(3))
    pc.update(dir_A0=lA.T.copy(), dir_Y0=lY.T.copy())
    gen.get("setbcdir")(lA, lY, lD)
    pc.update(dir_A1=lA.T.copy(), dir_Y1=lY.T.copy())
    M.eq[0].nbc = 0
    print(f"  SETBCDIR: {int((pc['dir_A1'] != pc['dir_A0']).any(axis=1).sum())} nodes overwritten")
    np.savez_compressed(os.path.join(HERE, "ref_pic.npz"), **pc)

    # ---- restart record: WRITERESTART (S/OUTPUT.f:132-232), fluid case (no displacement, no IB, no CEP)
    nX = 2
    M.stamp = np.array([1, 1, 1, nN, nX, 4, 0], dtype=np.int64)        # INITIALIZE.f:152
    M.cts, M.time, M.stfilename, M.stfilerepl, M.cepeq = 37, 0.185, "stFile", True, False
    M.recln = 4 * (1 + 7) + 8 * (2 + 1 + nX + 2 * 4 * nN)              # INITIALIZE.f:163
    cplr = rt.new("cplbctype"); cplr.xn = np.array([333.25, -12.5]); M.cplbc = cplr
    M.eq[0].inorm = 6.963e3
    rs_Yn, rs_An = rngl.standard_normal((nN, 4)), rngl.standard_normal((nN, 4))
    M.yn, M.an = np.asfortranarray(rs_Yn.T.copy()), np.asfortranarray(rs_An.T.copy())
    gen.rt.files.clear()
    gen.get("writerestart")(np.array([0.25, 0.0, 0.0]))
    rec = gen.rt.files[27][1]
    print(f"  WRITERESTART: record of {len(rec)} bytes (recLn = {M.recln})")
    np.savez_compressed(os.path.join(HERE, "ref_restart.npz"), record=np.frombuffer(rec, dtype=np.uint8), recLn=M.recln,
                        stamp=np.array(M.stamp), cTS=M.cts, time=M.time, timeP=1.5 - 0.25, iNorm=np.array([6.963e3]),
                        xn=np.array(cplr.xn), Yn=rs_Yn, An=rs_An)

    # ---- RCR (Windkessel) outlets: RCR_Integ_X (S/SETBC.f:1292-1372), two faces, three consecutive time steps
    cpl = rt.new("cplbctype")
    par = [(121.0, 1.5e-4, 1212.0, 0.0, 300.0), (80.0, 3.0e-4, 900.0, 50.0, 120.0)]      # Rp, C, Rd, Pd, X0
    cpl.nfa = len(par)
    cpl.fa = rx.FList(rt.new("cplfacetype") for _ in par)
    for fa, (Rp, Cc, Rd, Pd, X0) in zip(cpl.fa, par):
        fa.rcr.rp, fa.rcr.c, fa.rcr.rd, fa.rcr.pd = Rp, Cc, Rd, Pd
    cpl.xo = np.array([q[4] for q in par]); cpl.xn = np.zeros(len(par)); cpl.xp = np.zeros(len(par) + 1)
    M.cplbc, M.dt = cpl, 2.5e-3
    rc = dict(par=np.array(par), dt=M.dt)
    Q = np.array([[1.0, 0.4], [3.5, 1.1], [2.0, -0.3], [0.5, 0.2]])
    for k in range(3):
        M.time = (k + 1) * M.dt
        for i, fa in enumerate(cpl.fa):
            fa.qo, fa.qn = float(Q[k, i]), float(Q[k + 1, i])
        (istat,) = gen.get("rcr_integ_x")(0)
        rc[f"s{k}_xo"], rc[f"s{k}_xn"], rc[f"s{k}_y"] = np.array(cpl.xo), np.array(cpl.xn), np.array([fa.y for fa in cpl.fa])
        rc[f"s{k}_Q"], rc[f"s{k}_time"] = Q[k:k + 2].copy(), M.time
        cpl.xo = np.array(cpl.xn)
        print(f"  RCR step {k}: istat={istat} X={rc[f's{k}_xn']} y={rc[f's{k}_y']}")
    np.savez_compressed(os.path.join(HERE, "ref_rcr.npz"), **rc)

    # ---- case B: irregular mesh (Delaunay box, shuffled elements), body force and nodal body force Bf
    x, IEN = un.delaunay_box(n=70, seed=7)
    rowPtr, colPtr, Ag, Yg = un.problem(x, IEN)
    rng = np.random.default_rng(5)
    Bf = 0.3 * rng.standard_normal((x.shape[0], 3))
    fB = (0.1, -0.2, 0.3)
    R, V, _ = element_loop(gen, x, IEN, rowPtr, colPtr, Ag, Yg, cm.RHO, cm.MU, fB, cm.DT, cm.GA["af"], cm.GA["am"],
                           cm.GA["gam"], Bf=Bf)
    print(f"delaunay: nEl={IEN.shape[0]} nNo={x.shape[0]} |R|={np.abs(R).max():.3e} |Val|={np.abs(V).max():.3e}")
    np.savez_compressed(os.path.join(HERE, "ref_fluid_delaunay.npz"), x=x, IEN=IEN, rowPtr=rowPtr, colPtr=colPtr, Ag=Ag,
                        Yg=Yg, Bf=Bf, rho=cm.RHO, mu=cm.MU, f=np.array(fB), dt=cm.DT, af=cm.GA["af"], am=cm.GA["am"],
                        gam=cm.GA["gam"], R=R, Val=V)

    # ---- case C: heat equation (CONSTRUCT_HEATS / HEATS3D) on the lattice + CGRADS / GMRESS / BICGSS
    heat = dict(nu=0.7, s=0.3, rho=1.3)
    Ah = np.ascontiguousarray(p.Yg[:, :1] * 0.1)
    Yh = np.ascontiguousarray(p.Yg[:, 2:3])
    Rh, Vh, _ = element_loop(gen, p.rm.x, p.rm.IEN, p.rowPtr, p.colPtr, Ah, Yh, 0.0, 0.0, (0, 0, 0), cm.DT, cm.GA["af"],
                             cm.GA["am"], cm.GA["gam"], physics="heat", heat=heat)
    print(f"heat: |R|={np.abs(Rh).max():.3e} |Val|={np.abs(Vh).max():.3e}")
    hs = dict(x=p.rm.x, IEN=p.rm.IEN, rowPtr=p.rowPtr, colPtr=p.colPtr, Ag=Ah[:, 0], Yg=Yh[:, 0], dt=cm.DT, af=cm.GA["af"],
              am=cm.GA["am"], gam=cm.GA["gam"], R=Rh[:, 0], Val=Vh[:, 0], **heat)
    hfaces = [(p.faces[n]["gN"], 1, FM.bc_type_dir, None) for n in ("inlet", "outlet")]
    for name, lst, prec, kw in (("cg", FM.ls_type_cg, FM.precond_fsils, dict(relTol=1e-8, absTol=1e-14, maxItr=200)),
                                ("gmres", FM.ls_type_gmres, FM.precond_fsils, dict(relTol=1e-8, absTol=1e-14, maxItr=5, dimKry=20)),
                                ("bicgs", FM.ls_type_bicgs, FM.precond_rcs, dict(relTol=1e-8, absTol=1e-14, maxItr=200))):
        lhs = fsils_lhs(fg, p.rm.nNo, p.rowPtr, p.colPtr, hfaces)
        X, Vs, cnt = fsils_solve(fg, lhs, lst, 1, Rh, Vh, prec, incL=[1, 1], res=[0.0, 0.0], **kw)
        print(f"  FSILS heat {name}: itr={cnt['ri_itr']:.0f} suc={cnt['ri_suc']} iNorm={cnt['ri_inorm']:.6e} fNorm={cnt['ri_fnorm']:.3e}")
        hs[name + "_X"] = X[:, 0]
        hs[name + "_cnt"] = np.array([cnt[k] for k in sorted(cnt)], dtype=np.float64)
        hs[name + "_kw"] = np.array(repr(dict(ls_type=int(lst), prec=int(prec), **kw)))
    # CGRADV (L/CGRAD.f:52-123): a symmetric positive definite system with 3 dofs per node = the heat matrix x I3
    V3 = np.zeros((Vh.shape[0], 9)); V3[:, 0] = V3[:, 4] = V3[:, 8] = Vh[:, 0]
    R3 = np.random.default_rng(23).standard_normal((p.rm.nNo, 3))
    lhs = fsils_lhs(fg, p.rm.nNo, p.rowPtr, p.colPtr, [(p.faces[n]["gN"], 3, FM.bc_type_dir, None) for n in ("inlet", "outlet")])
    kw = dict(relTol=1e-8, absTol=1e-14, maxItr=200)
    X, Vs, cnt = fsils_solve(fg, lhs, FM.ls_type_cg, 3, R3, V3, FM.precond_fsils, incL=[1, 1], res=[0.0, 0.0], **kw)
    print(f"  FSILS CGRADV dof=3: itr={cnt['ri_itr']:.0f} suc={cnt['ri_suc']} iNorm={cnt['ri_inorm']:.6e} fNorm={cnt['ri_fnorm']:.3e}")
    hs.update(cgv_R=R3, cgv_Val=V3, cgv_X=X, cgv_cnt=np.array([cnt[k] for k in sorted(cnt)], dtype=np.float64),
              cgv_kw=np.array(repr(dict(ls_type=int(FM.ls_type_cg), prec=int(FM.precond_fsils), **kw))))
    hs["cnt_keys"] = np.array(sorted(cnt))
    np.savez_compressed(os.path.join(HERE, "ref_heat_lattice.npz"), **hs)
    # ---- FSILS on 2, 3 and 4 MPI tasks (emulated): axial slabs, and quadrant blocks whose axis nodes belong to all four tasks
    for nparts, dims, part in ((2, (2, 2, 4), "slabs"), (3, (2, 2, 4), "slabs"), (4, (4, 4, 2), "blocks")):
        mm, pp, _ = mesh.build_problem(*dims, nparts=nparts, L=2.0, partition=part)
        Rs, Vs = cm.oracle_assemble(pp)          # per-task element loop (pinned above to the reference, bit for bit)
        mcases = [("gmres_res", FM.ls_type_gmres, FM.precond_fsils, dict(relTol=1e-6, absTol=1e-14, maxItr=10, dimKry=30), 0.7),
                  ("ns", FM.ls_type_ns, FM.precond_fsils, dict(relTol=1e-3, absTol=1e-14, maxItr=10, dimKry=30), 0.0),
                  ("bicgs_rcs", FM.ls_type_bicgs, FM.precond_rcs, dict(relTol=1e-6, absTol=1e-14, maxItr=300), 0.0)]
        res, _ = fsils_multitask(pp, mm.nNo, Rs, Vs, mcases, cm.FACE_ORDER)
        mt = dict(nparts=nparts, dims=np.array(dims), L=2.0, partition=np.array(part))
        Rc_allfun = svfsi_commu_multitask(pp, mm.nNo, Rs)       # svFSI's own COMMU wrapper around FSILS_COMMUV
        for r in range(nparts):
            assert np.array_equal(Rc_allfun[r], res[r]["Rc"]), "COMMUV (ALLFUN) differs from permute + FSILS_COMMUV"
            mt[f"t{r}_Rc_allfun"] = Rc_allfun[r]
        for r, o in enumerate(res):
            mt.update({f"t{r}_{k}": v for k, v in o.items()})
        for name, lst, prec, kw, res_out in mcases:
            mt[name + "_kw"] = np.array(repr(dict(ls_type=int(lst), prec=int(prec), res_out=res_out, **kw)))
            c = dict(zip(res[0]["cnt_keys"], res[0][name + "_cnt"]))
            print(f"  FSILS {nparts} tasks {name}: itr={c['ri_itr']:.0f} GM={c['gm_itr']:.0f} CG={c['cg_itr']:.0f} iNorm={c['ri_inorm']:.6e} fNorm={c['ri_fnorm']:.3e}")
        np.savez_compressed(os.path.join(HERE, f"ref_fsils_{nparts}tasks.npz"), **mt)
    # ---- manifest: every procedure of the reference that was translated from its source for the vectors above
    lines = set()
    for g in (gen, fg, MT_GEN[0]):
        for name in g.done:
            u = g.lib.units[name]
            lines.add(f"{os.path.relpath(u.file, os.path.join(REF, 'Code', 'Source'))} : {name.upper()}"
                      + (f" (+ internal {', '.join(i.name.upper() for i in u.internal)})" if u.internal else ""))
    with open(os.path.join(HERE, "ref_routines.txt"), "w") as fh:
        fh.write("# procedures of /root/reference/Code/Source that oracle/refexec.py translated from their source text for the\n"
                 "# golden vectors (a procedure is translated when a translated caller references it; the branches taken by the\n"
                 "# cases of tests/golden/make_ref_golden.py were executed).  file : procedure\n")
        fh.write("\n".join(sorted(lines)) + "\n")
    print(f"done in {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
