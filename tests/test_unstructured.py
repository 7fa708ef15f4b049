"""Irregular meshes (tests/unstructured.py): CPU checks of the oracle on them, and -m gpu parity of
the CUDA assembly / SpMV / Jacobi-GMRES path against the oracle on the same meshes -- arbitrary
valence and element numbering, and a row longer than the row-owner assembly kernel's 64-block limit."""
import numpy as np
import pytest

import common as cm
import unstructured as un
from oracle import oracle as ora
from svfsi_b200 import api

TOL_ASM = 1e-12
MESHES = {"delaunay": un.delaunay_box, "fan": un.fan}


def _oracle(x, IEN, rowPtr, colPtr, Ag, Yg):
    return ora.construct_fluid(cm.fluid_par(), IEN, x, Ag, Yg, np.zeros((x.shape[0], 3)), rowPtr,
                               colPtr)


@pytest.mark.parametrize("name", list(MESHES))
def test_mesh_is_valid_and_irregular(name):
    x, IEN = MESHES[name]()
    t = IEN.astype(np.int64) - 1
    det = np.linalg.det(x[t[:, :3]] - x[t[:, 3:4]])
    assert (det > 0).all() and np.unique(t).size == x.shape[0]
    rowPtr, colPtr = un.problem(x, IEN)[:2]
    rowlen = np.diff(rowPtr)
    assert rowlen.max() != rowlen.min()
    if name == "fan":
        assert rowlen.max() > 64 and np.bincount(t.ravel()).max() > 64


@pytest.mark.parametrize("name", list(MESHES))
def test_oracle_hydrostatic_state_on_irregular_mesh(name):
    """u = 0, p linear, body force f = grad p / rho: the momentum residual vanishes node by node
    (interior AND boundary: the pressure term is integrated by parts but there is no traction face,
    so only interior nodes are checked) and the continuity residual is zero."""
    x, IEN = MESHES[name]()
    rowPtr, colPtr, _, _ = un.problem(x, IEN)
    nNo = x.shape[0]
    Yg = np.zeros((nNo, 4)); Yg[:, 3] = 7.0 - 3.0 * x[:, 2]
    Ag = np.zeros((nNo, 4))
    par = ora.fluid_par(cm.RHO, cm.MU, (0.0, 0.0, -3.0 / cm.RHO), cm.DT, cm.GA["af"], cm.GA["am"],
                        cm.GA["gam"])
    R, V = ora.construct_fluid(par, IEN, x, Ag, Yg, np.zeros((nNo, 3)), rowPtr, colPtr)
    assert np.abs(R[:, 3]).max() <= 1e-13
    # interior nodes = nodes none of whose element faces is on the boundary
    t = IEN.astype(np.int64) - 1
    faces = np.sort(np.concatenate([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]]), axis=1)
    uf, cnt = np.unique(faces, axis=0, return_counts=True)
    bnd = np.zeros(nNo, dtype=bool); bnd[uf[cnt == 1].ravel()] = True
    if (~bnd).any():
        assert np.abs(R[~bnd, :3]).max() <= 1e-12 * max(1.0, np.abs(R[:, :3]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MESHES))
def test_gpu_assembly_and_solve_on_irregular_mesh(gpu_lib, name):
    x, IEN = MESHES[name]()
    rowPtr, colPtr, Ag, Yg = un.problem(x, IEN)
    nNo = x.shape[0]
    ltg = np.arange(1, nNo + 1, dtype=np.int32)
    Rr, Vr = _oracle(x, IEN, rowPtr, colPtr, Ag, Yg)
    api.FSILS_LHS_CREATE(nNo, nNo, colPtr.size, ltg, rowPtr, colPtr, 0)
    try:
        api.mesh_create(IEN, x)
        got = {}
        for variant in (api.ASM_ATOMIC, api.ASM_COLORED, api.ASM_GATHER):
            api.CONSTRUCT_FLUID(Ag, Yg, None, cm.RHO, cm.MU, cm.F, cm.DT, cm.GA["af"], cm.GA["am"],
                                cm.GA["gam"], variant)
            R, V = api.get_R(4), api.get_Val(4)
            got[variant] = (R, V)
            assert cm.rel_err(R[:, :3], Rr[:, :3]) <= TOL_ASM, (name, variant)
            assert cm.rel_err(R[:, 3], Rr[:, 3]) <= TOL_ASM, (name, variant)
            errs = cm.block_class_errs(V, Vr)
            assert max(errs.values()) <= TOL_ASM, (name, variant, errs)
        # every gather-kernel variant (row-owner ones fall back to block-owner on the fan)
        for tune in (0, 8, 40, 104, 808, 128 + 40, 1024, 128 + 1024, 128 + 11264, 16384, 128 + 24576, 128 + 57344, 131072, 128 + 139264, 262144, 128 + 270336, 128 + 790528):
            api.time_kernel(5, 4, 7, 1, tune)
            assert cm.rel_err(api.get_R(4), Rr) <= TOL_ASM, (name, tune)
            assert max(cm.block_class_errs(api.get_Val(4), Vr).values()) <= TOL_ASM, (name, tune)
        # FSILS_SPARMULVV on the assembled matrix
        rng = np.random.default_rng(2)
        U = rng.standard_normal((nNo, 4))
        w = ora.World(nNo, [ltg], [rowPtr], [colPtr], 0)
        KU_o = w.sparmul_vv(4, [Vr], [U])[0]
        KU = api.FSILS_SPARMUL("VV", 4, V, U)
        assert cm.rel_err(KU, KU_o) <= 1e-13
        # the small block shapes, every kernel family (rows of up to 91 blocks: several trips per row, trips
        # that start in the middle of a block)
        for mode in range(0, 14):
            api.set_spmv_small(mode)
            for kind, dof in (("VV", 3), ("VS", 3), ("SV", 3), ("SS", 1)):
                br = dof if kind in ("VV", "SV") else 1
                bc = dof if kind in ("VV", "VS") else 1
                Ks = rng.standard_normal((colPtr.size, br * bc)); Us = rng.standard_normal((nNo, bc))
                if kind == "VV":
                    ref = w.sparmul_vv(dof, [Ks], [Us])[0]
                elif kind == "VS":
                    ref = w.sparmul_vs(dof, [Ks], [Us])[0]
                elif kind == "SV":
                    ref = w.sparmul_sv(dof, [Ks], [Us.reshape(-1)])[0]
                else:
                    ref = w.sparmul_ss([Ks.reshape(-1)], [Us.reshape(-1)])[0]
                got = api.FSILS_SPARMUL(kind, dof, Ks, Us)
                assert cm.rel_err(got.reshape(ref.shape), ref) <= 1e-13, (name, mode, kind, dof)
        api.set_spmv_small(-1)
        # Jacobi + GMRES Newton step (no Dirichlet face: the mass term keeps the matrix regular)
        api.CONSTRUCT_FLUID(Ag, Yg, None, cm.RHO, cm.MU, cm.F, cm.DT, cm.GA["af"], cm.GA["am"],
                            cm.GA["gam"], api.ASM_GATHER)
        ls = api.FSILS_LS_CREATE(api.LS_TYPE_GMRES, relTol=1e-6, absTol=1e-14, maxItr=10, dimKry=60)
        api.solve_dev(ls, 4)
        X = api.get_R(4)
        ls_o = ora.ls_create(ora.LS_TYPE_GMRES, relTol=1e-6, absTol=1e-14, maxItr=10, dimKry=60)
        Xo = Rr.copy()
        w.solve(ls_o, 4, [Xo], [Vr.copy()])
        assert abs(ls.RI.itr - ls_o.RI.itr) <= 1
        assert np.linalg.norm(X - Xo) / np.linalg.norm(Xo) <= 1e-8
    finally:
        api.FSILS_LHS_FREE()


@pytest.mark.gpu
def test_gpu_degenerate_element_is_reported(gpu_lib):
    """ISZERO(Jac) -> err "Jac < 0 @ element" (S/FLUID.f:115): the C-ABI returns SVFSI_ERR_JAC."""
    x, IEN = un.delaunay_box(n=60, seed=4)
    x = x.copy()
    t0 = IEN[0].astype(np.int64) - 1
    x[t0[3]] = x[t0[0]]          # element 1 collapses: Jac == 0 exactly (ISZERO is |Jac| < ~5e-31)
    rowPtr, colPtr, Ag, Yg = un.problem(x, IEN)
    nNo = x.shape[0]
    api.FSILS_LHS_CREATE(nNo, nNo, colPtr.size, np.arange(1, nNo + 1, dtype=np.int32), rowPtr, colPtr, 0)
    try:
        api.mesh_create(IEN, x)
        for variant in (api.ASM_ATOMIC, api.ASM_GATHER):
            with pytest.raises(api.SvfsiError) as ei:
                api.CONSTRUCT_FLUID(Ag, Yg, None, cm.RHO, cm.MU, cm.F, cm.DT, cm.GA["af"],
                                    cm.GA["am"], cm.GA["gam"], variant)
            assert ei.value.code == api.ERR_JAC
    finally:
        api.FSILS_LHS_FREE()
