"""ctypes binding of ``lib/libsvfsi_b200.so`` (the C-ABI in ``include/svfsi_b200.h``) plus a thin
host-side mirror of the reference's call sites, with the reference's names and argument meaning:

=====================  ===========================================  ==========================
here                   reference                                    file:line (Code/Source/)
=====================  ===========================================  ==========================
``FSILS_LHS_CREATE``   ``FSILS_LHS_CREATE(lhs, commu, gnNo, ...)``  svFSILS/LHS.f:51
``FSILS_BC_CREATE``    ``FSILS_BC_CREATE(lhs, faIn, nNo, dof, ..)`` svFSILS/BC.f:50
``FSILS_LS_CREATE``    ``FSILS_LS_CREATE(ls, LS_type, ...)``        svFSILS/LS.f:50
``FSILS_SOLVE``        ``FSILS_SOLVE(lhs, ls, dof, Ri, Val, ...)``  svFSILS/SOLVE.f:51
``FSILS_COMMUV``       ``FSILS_COMMUV(lhs, dof, R)``                svFSILS/INCOMMU.f:56
``CONSTRUCT_FLUID``    ``CONSTRUCT_FLUID(lM, Ag, Yg)``              svFSI/FLUID.f:40
``CONSTRUCT_HEATS``    ``CONSTRUCT_HEATS(lM, Ag, Yg)``              svFSI/HEATS.f:39
=====================  ===========================================  ==========================

All arrays are numpy, C-contiguous with the Fortran memory layout of SURVEY.md Appendix B
(``R(dof,tnNo)`` is ``R[tnNo][dof]`` etc.), ids are 1-based int32.  There is no CPU fallback: if the
CUDA library is missing or no GPU is present every compute call raises ``SvfsiError``.
PyTorch is not needed by this module (it is used by bench.py / tests only for process-group
plumbing).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsvfsi_b200.so")

LS_TYPE_CG, LS_TYPE_GMRES, LS_TYPE_NS, LS_TYPE_BICGS = 798, 797, 796, 795
PRECOND_FSILS, PRECOND_RCS = 701, 709
BC_TYPE_Dir, BC_TYPE_Neu = 0, 1
ASM_ATOMIC, ASM_COLORED, ASM_GATHER = 0, 1, 2
NTIMERS = 16
PROF_SLOTS = ["asm", "spmv", "halo", "dot", "axpy", "precond", "small", "allreduce", "solve"]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


# error codes of include/svfsi_b200.h
(OK, ERR_CUDA, ERR_STATE, ERR_ARG, ERR_JAC, ERR_COMM, ERR_UNSUPPORTED) = range(7)


class SvfsiError(RuntimeError):
    """A non-zero return code of the C-ABI (the reference would PRINT + STOP); `.code` holds it."""

    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


class SubLs(C.Structure):
    _fields_ = [("suc", C.c_int32), ("mItr", C.c_int32), ("sD", C.c_int32), ("itr", C.c_int32),
                ("absTol", C.c_double), ("relTol", C.c_double), ("iNorm", C.c_double),
                ("fNorm", C.c_double), ("dB", C.c_double), ("callD", C.c_double)]


class Ls(C.Structure):
    _fields_ = [("LS_type", C.c_int32), ("Resm", C.c_int32), ("Resc", C.c_int32),
                ("reserved", C.c_int32), ("GM", SubLs), ("CG", SubLs), ("RI", SubLs)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _ip, C.c_int32, _ip)

# every symbol include/svfsi_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "gpu_nccl_unique_id_", "gpu_init_", "gpu_set_host_allgather_", "gpu_finalize_",
    "gpu_last_error_", "gpu_lhs_create_", "gpu_lhs_free_", "gpu_lhs_info_", "gpu_lhs_cs_",
    "svfsi_lhs_plan_", "gpu_bc_create_", "gpu_bc_free_", "gpu_mesh_create_", "gpu_mesh_ncolors_",
    "gpu_construct_fluid_", "gpu_construct_heats_", "gpu_state_upload_",
    "gpu_construct_fluid_dev_", "gpu_construct_heats_dev_", "gpu_get_r_", "gpu_set_r_",
    "gpu_get_val_", "gpu_set_val_", "gpu_commu_", "gpu_commu_dev_", "gpu_solve_",
    "gpu_solve_dev_", "gpu_ls_create_", "gpu_sparmul_", "gpu_dot_", "gpu_time_kernel_",
    "gpu_prof_enable_", "gpu_prof_reset_", "gpu_prof_get_", "gpu_launch_count_",
    "gpu_get_stream_", "gpu_sync_", "gpu_comm_mode_", "gpu_spmv_variant_",
    "gpu_pic_init_", "gpu_pic_free_", "gpu_picp_", "gpu_setbcdir_", "gpu_pici_", "gpu_picc_",
    "gpu_pic_advance_", "gpu_pic_get_",
    "gpu_face_create_", "gpu_face_free_", "gpu_bassem_neu_fluid_", "gpu_face_integ_v_",
    "gpu_prof_spmv_", "gpu_set_comm_timeout_", "gpu_set_spmv_small_",
]


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "clean"],
                              stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"],
                          stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SvfsiError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                             "g.build()'` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        for name in EXPORTS:
            getattr(_lib, name).restype = C.c_int32
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _ci(v):
    return C.byref(C.c_int32(int(v)))


def _cd(v):
    return C.byref(C.c_double(float(v)))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _check(rc):
    if rc != 0:
        buf = C.create_string_buffer(512)
        lib().gpu_last_error_(buf, _ci(512))
        raise SvfsiError(f"svfsi_b200 error {rc}: {buf.value.decode(errors='replace')}", rc)


# --------------------------------------------------------------------------- life cycle
_keep = {}


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().gpu_nccl_unique_id_(buf))
    return buf.raw


def init(device: int = 0, rank: int = 0, nranks: int = 1, uid: bytes | None = None):
    """FSILS_COMMU_CREATE analogue (svFSILS/COMMU.f:50-86): rank, size, communicator."""
    ub = C.create_string_buffer(uid, 128) if uid is not None else None
    _check(lib().gpu_init_(_ci(device), _ci(rank), _ci(nranks), ub))


def set_host_allgather(fn):
    """fn(send: np.int32[n]) -> np.int32[nranks*n]; lent to the library for setup collectives."""
    def tramp(_ctx, send, n, recv):
        try:
            s = np.ctypeslib.as_array(send, shape=(n,)).copy() if n > 0 else np.zeros(0, np.int32)
            out = np.ascontiguousarray(fn(s), dtype=np.int32)
            if out.size:
                C.memmove(recv, out.ctypes.data, out.nbytes)
            return 0
        except Exception as ex:  # pragma: no cover
            print("host allgather failed:", ex)
            return 1
    cb = ALLGATHER_FN(tramp)
    _keep["allgather"] = cb
    _check(lib().gpu_set_host_allgather_(cb, None))


def init_distributed(device=None):
    """One process per GPU under torchrun: NCCL unique id broadcast through torch.distributed."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    obj = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)

    def ag(send):
        t = torch.from_numpy(send.copy())
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return torch.cat(outs).cpu().numpy()
    init(device, rank, world, obj[0])
    set_host_allgather(ag)


def finalize():
    _check(lib().gpu_finalize_())


def sync():
    _check(lib().gpu_sync_())


# --------------------------------------------------------------------------- FSILS mirror
def FSILS_LHS_CREATE(gnNo, nNo, nnz, gNodes, rowPtr, colPtr, nFaces):
    gNodes = _i32(gNodes); rowPtr = _i32(rowPtr); colPtr = _i32(colPtr)
    assert gNodes.size == nNo and rowPtr.size == nNo + 1 and colPtr.size == nnz
    _check(lib().gpu_lhs_create_(_ci(gnNo), _ci(nNo), _ci(nnz), _i(gNodes), _i(rowPtr), _i(colPtr),
                                 _ci(nFaces)))
    _keep["nNo"], _keep["nnz"] = int(nNo), int(nnz)


def FSILS_LHS_FREE():
    _check(lib().gpu_lhs_free_())


def lhs_info():
    nNo = _keep["nNo"]
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    m = np.zeros(nNo, dtype=np.int32)
    _check(lib().gpu_lhs_info_(C.byref(a), C.byref(b), C.byref(c), _i(m)))
    cs = []
    for i in range(1, c.value + 1):
        iP, n = C.c_int32(), C.c_int32()
        _check(lib().gpu_lhs_cs_(_ci(i), C.byref(iP), C.byref(n), None))
        ptr = np.zeros(n.value, dtype=np.int32)
        _check(lib().gpu_lhs_cs_(_ci(i), C.byref(iP), C.byref(n), _i(ptr)))
        cs.append((iP.value, ptr))
    return dict(mynNo=a.value, shnNo=b.value, nReq=c.value, map=m, cS=cs)


def lhs_plan(rank, nranks, gnNo, ltg_list):
    """Host-only node reordering / halo schedule of FSILS_LHS_CREATE for `rank` (0-based) given
    every rank's ltg list -- no CUDA call, usable on a CPU-only box."""
    maxnNo = max(len(l) for l in ltg_list)
    aN = np.zeros((nranks, maxnNo), dtype=np.int32)
    for r, l in enumerate(ltg_list):
        aN[r, :len(l)] = l
    nNo = len(ltg_list[rank])
    m = np.zeros(nNo, dtype=np.int32)
    my, sh, nr = C.c_int32(), C.c_int32(), C.c_int32()
    iP = np.zeros(nranks, dtype=np.int32); n = np.zeros(nranks, dtype=np.int32)
    cap = nranks * nNo
    ptr = np.zeros(max(cap, 1), dtype=np.int32)
    _check(lib().svfsi_lhs_plan_(_ci(rank), _ci(nranks), _ci(gnNo), _ci(nNo), _ci(maxnNo), _i(aN),
                                 _i(m), C.byref(my), C.byref(sh), C.byref(nr), _i(iP), _i(n),
                                 _i(ptr), _ci(cap)))
    cs, off = [], 0
    for k in range(nr.value):
        cs.append((int(iP[k]), ptr[off:off + n[k]].copy()))
        off += n[k]
    return dict(mynNo=my.value, shnNo=sh.value, nReq=nr.value, map=m, cS=cs)


def FSILS_BC_CREATE(faIn, nNo, dof, BC_type, gNodes, Val=None):
    gNodes = _i32(gNodes)
    v = _f64(Val) if Val is not None else None
    _check(lib().gpu_bc_create_(_ci(faIn), _ci(nNo), _ci(dof), _ci(BC_type), _i(gNodes), _d(v)))


def FSILS_BC_FREE(faIn):
    _check(lib().gpu_bc_free_(_ci(faIn)))


def FSILS_LS_CREATE(LS_type, relTol=None, absTol=None, maxItr=None, dimKry=None,
                    relTolIn=None, absTolIn=None, maxItrIn=None) -> Ls:
    """Defaults of svFSILS/LS.f:69-95 with the optional overrides of :97-116."""
    ls = Ls()
    _check(lib().gpu_ls_create_(C.byref(ls), _ci(LS_type)))
    if relTol is not None: ls.RI.relTol = relTol
    if absTol is not None: ls.RI.absTol = absTol
    if maxItr is not None: ls.RI.mItr = maxItr
    if dimKry is not None:
        ls.RI.sD = dimKry; ls.GM.sD = dimKry
    if relTolIn is not None: ls.GM.relTol, ls.CG.relTol = relTolIn
    if absTolIn is not None: ls.GM.absTol, ls.CG.absTol = absTolIn
    if maxItrIn is not None: ls.GM.mItr, ls.CG.mItr = maxItrIn
    return ls


def FSILS_SOLVE(ls: Ls, dof, Ri, Val=None, prec=PRECOND_FSILS, incL=None, res=None):
    """Ri(dof,nNo) in: RHS, out: solution (in place).  Val=None uses the device-resident matrix
    left by CONSTRUCT_*; a host Val is uploaded (and, like the reference, considered consumed)."""
    assert Ri.dtype == np.float64 and Ri.flags.c_contiguous
    v = _f64(Val) if Val is not None else None
    il = _i32(incL) if incL is not None else None
    rs = _f64(res) if res is not None else None
    _check(lib().gpu_solve_(C.byref(ls), _ci(dof), _d(Ri), _d(v), _ci(prec), _i(il), _d(rs)))
    return ls


def solve_dev(ls: Ls, dof, prec=PRECOND_FSILS, incL=None, res=None):
    il = _i32(incL) if incL is not None else None
    rs = _f64(res) if res is not None else None
    _check(lib().gpu_solve_dev_(C.byref(ls), _ci(dof), _ci(prec), _i(il), _d(rs)))
    return ls


def FSILS_COMMUV(dof, R):
    assert R.dtype == np.float64 and R.flags.c_contiguous
    _check(lib().gpu_commu_(_ci(dof), _d(R)))


def commu_dev(dof):
    _check(lib().gpu_commu_dev_(_ci(dof)))


def FSILS_SPARMUL(kind, dof, K, U):
    """kind: 'VV','VS','SV','SS' (svFSILS/SPARMUL.f:51-297), halo sum included."""
    k = {"VV": 0, "VS": 1, "SV": 2, "SS": 3}[kind]
    K = _f64(K); U = _f64(U)
    nNo = _keep["nNo"]
    br = dof if k in (0, 2) else 1
    if k == 3:
        br = 1
    KU = np.zeros((nNo, br)) if br > 1 else np.zeros(nNo)
    _check(lib().gpu_sparmul_(_ci(k), _ci(dof), _d(K), _d(U), _d(KU)))
    return KU


def FSILS_DOTV(dof, U, V):
    out = C.c_double()
    U = _f64(U); V = _f64(V)
    _check(lib().gpu_dot_(_ci(dof), _d(U), _d(V), C.byref(out)))
    return out.value


# --------------------------------------------------------------------------- element loop mirror
def mesh_create(IEN, x):
    IEN = _i32(IEN); x = _f64(x)
    _check(lib().gpu_mesh_create_(_ci(IEN.shape[0]), _ci(IEN.shape[1]), _i(IEN), _d(x)))
    _keep["nEl"] = IEN.shape[0]


def mesh_ncolors():
    n = C.c_int32()
    _check(lib().gpu_mesh_ncolors_(C.byref(n)))
    return n.value


def CONSTRUCT_FLUID(Ag, Yg, Bf, rho, mu, f, dt, af, am, gam, variant=ASM_ATOMIC):
    Ag = _f64(Ag); Yg = _f64(Yg); Bf = _f64(Bf) if Bf is not None else None
    f = _f64(f)
    _check(lib().gpu_construct_fluid_(_d(Ag), _d(Yg), _d(Bf), _cd(rho), _cd(mu), _d(f), _cd(dt),
                                      _cd(af), _cd(am), _cd(gam), _ci(variant)))


def CONSTRUCT_HEATS(Ag, Yg, nu, s, rho, dt, af, am, gam, variant=ASM_ATOMIC):
    Ag = _f64(Ag); Yg = _f64(Yg)
    _check(lib().gpu_construct_heats_(_d(Ag), _d(Yg), _cd(nu), _cd(s), _cd(rho), _cd(dt), _cd(af),
                                      _cd(am), _cd(gam), _ci(variant)))


def state_upload(tDof, Ag, Yg, Bf=None):
    Ag = _f64(Ag); Yg = _f64(Yg); Bf = _f64(Bf) if Bf is not None else None
    _check(lib().gpu_state_upload_(_ci(tDof), _d(Ag), _d(Yg), _d(Bf)))


def construct_fluid_dev(rho, mu, f, dt, af, am, gam, variant=ASM_ATOMIC):
    f = _f64(f)
    _check(lib().gpu_construct_fluid_dev_(_cd(rho), _cd(mu), _d(f), _cd(dt), _cd(af), _cd(am),
                                          _cd(gam), _ci(variant)))


def construct_heats_dev(nu, s, rho, dt, af, am, gam, variant=ASM_ATOMIC):
    _check(lib().gpu_construct_heats_dev_(_cd(nu), _cd(s), _cd(rho), _cd(dt), _cd(af), _cd(am),
                                          _cd(gam), _ci(variant)))


def get_R(dof):
    nNo = _keep["nNo"]
    R = np.zeros((nNo, dof)) if dof > 1 else np.zeros(nNo)
    _check(lib().gpu_get_r_(_ci(dof), _d(R)))
    return R


def set_R(dof, R):
    _check(lib().gpu_set_r_(_ci(dof), _d(_f64(R))))


def get_Val(dof):
    nnz = _keep["nnz"]
    V = np.zeros((nnz, dof * dof)) if dof > 1 else np.zeros(nnz)
    _check(lib().gpu_get_val_(_ci(dof), _d(V)))
    return V


def set_Val(dof, Val):
    _check(lib().gpu_set_val_(_ci(dof), _d(_f64(Val))))


# --------------------------------------------------------------------------- measurement
def time_kernel(what, dof=4, k=1, reps=10, variant=0):
    ms = C.c_double()
    _check(lib().gpu_time_kernel_(_ci(what), _ci(dof), _ci(k), _ci(reps), _ci(variant),
                                  C.byref(ms)))
    return ms.value


def set_spmv_small(mode):
    """kernel family of the small-block SpMV shapes: -1 per-shape default, 0 lane-per-block, 1..4 (SVFSI_SPMV_SMALL)"""
    _check(lib().gpu_set_spmv_small_(_ci(mode)))


def prof_enable(on=True):
    _check(lib().gpu_prof_enable_(_ci(1 if on else 0)))


def prof_reset():
    _check(lib().gpu_prof_reset_())


def prof_get():
    ms = np.zeros(NTIMERS); n = np.zeros(NTIMERS, dtype=np.int64)
    _check(lib().gpu_prof_get_(_d(ms), n.ctypes.data_as(C.POINTER(C.c_int64))))
    return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(PROF_SLOTS)}


def prof_spmv():
    b = C.c_double(); n = C.c_int64()
    _check(lib().gpu_prof_spmv_(C.byref(b), C.byref(n)))
    return b.value, n.value


# ---------------------------------------------------------------------------------------------
# generalised-alpha time integration on the device: PICP / SETBCDIR / PICI / PICC (S/PIC.f, S/SETBC.f)
class EqState:
    """the scalars of eqType that PICC's convergence decision reads (S/PIC.f:262-275, S/MOD.f)"""

    def __init__(self, tol=1e-6, absTol=1e-12, minItr=1, maxItr=10):
        self.tol, self.absTol, self.minItr, self.maxItr = tol, absTol, minItr, maxItr
        self.itr, self.iNorm, self.pNorm, self.ok = 0, 0.0, 0.0, False


def _iszero(x):
    """ISZERO with one argument, S/UTIL.f:879-903"""
    eps = np.finfo(np.float64).eps
    a = abs(x)
    return a / max(a, eps) < 10.0 * eps


def pic_init(tDof, Ao, Yo, Do=None):
    _check(lib().gpu_pic_init_(_ci(tDof), _d(_f64(Ao)), _d(_f64(Yo)),
                               _d(_f64(Do)) if Do is not None else None))


def PICP(gam):
    _check(lib().gpu_picp_(_cd(gam)))


def SETBCDIR(gN, s, tmpA, tmpY):
    """tmpA / tmpY = (nNo_face, lDof) arrays as SETBCDIRL fills them (S/SETBC.f:202-232); s = first
    dof (1-based) of the equation"""
    gN = _i32(gN); tmpA = _f64(tmpA); tmpY = _f64(tmpY)
    lDof = 1 if tmpA.ndim == 1 else tmpA.shape[1]
    _check(lib().gpu_setbcdir_(_ci(gN.size), _i(gN), _ci(s), _ci(lDof), _d(tmpA), _d(tmpY)))


def PICI(eq: EqState, am, af):
    eq.itr += 1                       # S/PIC.f:139
    _check(lib().gpu_pici_(_cd(am), _cd(af)))


def PICC(eq: EqState, ls: "Ls", gam, beta, dt):
    """corrector + the convergence decision of S/PIC.f:262-275 (single, uncoupled equation)"""
    _check(lib().gpu_picc_(_cd(gam), _cd(beta), _cd(dt)))
    return picc_decision(eq, ls.RI)


def picc_decision(eq: EqState, RI):
    """S/PIC.f:262-275: the scalar part of PICC (RI = the FSILS_subLsType of the solve that just ran)"""
    eps = float(np.finfo(np.float64).eps)
    if _iszero(RI.iNorm):
        RI.iNorm = eps
    if _iszero(eq.iNorm):
        eq.iNorm = RI.iNorm
    if eq.itr == 1:
        eq.pNorm = RI.iNorm / eq.iNorm
    r1 = RI.iNorm / eq.iNorm
    l1 = eq.iNorm <= eq.absTol
    l2 = eq.itr >= eq.maxItr
    l3 = r1 <= eq.tol
    l4 = r1 <= eq.tol * eq.pNorm
    l5 = eq.itr >= eq.minItr
    if l1 or l2 or ((l3 or l4) and l5):
        eq.ok = True
    return eq.ok


def pici(am, af):
    """the vector part of PICI alone (S/PIC.f:141-152), no eqType bookkeeping"""
    _check(lib().gpu_pici_(_cd(am), _cd(af)))


def picc(gam, beta, dt):
    """the vector part of PICC alone (S/PIC.f:203-207)"""
    _check(lib().gpu_picc_(_cd(gam), _cd(beta), _cd(dt)))


def pic_advance(eq: EqState | None = None):
    _check(lib().gpu_pic_advance_())
    if eq is not None:
        eq.itr, eq.ok = 0, False      # S/MAIN.f:99-100


def pic_get(which, tDof, nNo, withD=False):
    shp = (nNo, tDof) if tDof > 1 else (nNo,)
    A = np.zeros(shp); Y = np.zeros(shp); D = np.zeros(shp) if withD else None
    _check(lib().gpu_pic_get_(_ci(which), _d(A), _d(Y), _d(D) if withD else None))
    return (A, Y, D) if withD else (A, Y)


# ---------------------------------------------------------------------------------------------
# face integrals on the device: BASSEMNEUBC/BFLUID (S/EQASSEM.f:90-192, S/FLUID.f:1279-1336), IntegV
def face_create(iFa, gN, IEN, gE):
    """gN (nNo,), IEN (nEl, 3), gE (nEl,): 1-based ids in svFSI's local numbering (faceType)"""
    gN = _i32(gN); IEN = _i32(IEN); gE = _i32(gE)
    _check(lib().gpu_face_create_(_ci(iFa), _ci(gN.size), _i(gN), _ci(gE.size), _ci(3), _i(IEN), _i(gE)))


def face_free(iFa):
    _check(lib().gpu_face_free_(_ci(iFa)))


def BASSEMNEUBC_FLUID(iFa, hgN, rho, bfStab, af, gam, dt):
    hgN = _f64(hgN)
    _check(lib().gpu_bassem_neu_fluid_(_ci(iFa), _d(hgN), _cd(rho), _cd(bfStab), _cd(af), _cd(gam), _cd(dt)))


def IntegV(iFa, which=0, s=1):
    out = C.c_double()
    _check(lib().gpu_face_integ_v_(_ci(iFa), _ci(which), _ci(s), C.byref(out)))
    return out.value


COMM_MODES = {0: "single rank", 1: "nccl", 2: "peer-memory kernels (CUDA IPC over NVLink)",
              3: "peer-memory, halo send fused into the SpMV kernel"}


def comm_mode():
    m = C.c_int32()
    _check(lib().gpu_comm_mode_(C.byref(m)))
    return m.value


def spmv_variant():
    v = C.c_int32()
    _check(lib().gpu_spmv_variant_(C.byref(v)))
    return v.value


def set_comm_timeout(seconds):
    """bound of every in-kernel wait on a peer's flag; afterwards calls return ERR_COMM"""
    _check(lib().gpu_set_comm_timeout_(_cd(seconds)))


def launch_count():
    n = C.c_int64()
    _check(lib().gpu_launch_count_(C.byref(n)))
    return n.value


def stream_ptr():
    p = C.c_void_p()
    _check(lib().gpu_get_stream_(C.byref(p)))
    return p.value
