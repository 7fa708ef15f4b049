"""-m gpu: BASELINE.json configs[1] ("synthetic 1M-tet cylinder, unsteady Navier-Stokes, FSILS GMRES +
diagonal precond, 1 GPU") against the oracle AT THAT SIZE: 40 x 40 x 104 Kuhn lattice = 998 400 tets,
176 505 nodes, 2.57M blocks.  The small-mesh parity tests cannot reach the code paths the benchmark
runs -- block descriptors in length-sorted chunks of 512, 148-SM grids with several waves, multi-tile
multi-dots (k > 8), partial-sum trees over all 592 CTAs -- this one does, with the north_star
tolerances: assembled R / Val <= 1e-12 per block class, FSILS_SPARMULVV <= 1e-14, GMRES iteration count
+-1, iNorm 1e-10, Newton step <= 1e-8.  The oracle needs ~15 s for it (element loop 6 s, GMRES 5 s)."""
import numpy as np
import pytest

import common as cm
from oracle import oracle as ora
from svfsi_b200 import api, mesh

pytestmark = pytest.mark.gpu

DIMS = (40, 40, 104)
L_PIPE = 30.0 * 104 / 408          # the bench's axial spacing
LS = dict(relTol=1e-3, absTol=1e-12, maxItr=10, dimKry=50)     # bench.py's solver settings


@pytest.fixture(scope="module")
def c2(gpu_lib):
    m, probs, _ = mesh.build_problem(*DIMS, nparts=1, L=L_PIPE)
    p = probs[0]
    api.FSILS_LHS_CREATE(m.nNo, p.rm.nNo, p.colPtr.size, p.rm.ltg, p.rowPtr, p.colPtr, 3)
    for fi, name in enumerate(cm.FACE_ORDER, start=1):
        fa = p.faces[name]
        api.FSILS_BC_CREATE(fi, fa["gN"].size, 3,
                            api.BC_TYPE_Neu if fa["bc"] == "Neu" else api.BC_TYPE_Dir, fa["gN"], fa["val"])
    api.mesh_create(p.rm.IEN, p.rm.x)
    Rs, Vs = cm.oracle_assemble([p], native=False)
    yield m, p, Rs[0], Vs[0]
    api.FSILS_LHS_FREE()


@pytest.mark.parametrize("variant", [api.ASM_GATHER, api.ASM_ATOMIC, api.ASM_COLORED])
def test_c2_assembly(c2, variant):
    m, p, Ro, Vo = c2
    assert m.nEl == 998400
    api.CONSTRUCT_FLUID(p.Ag, p.Yg, None, cm.RHO, cm.MU, cm.F, cm.DT, cm.GA["af"], cm.GA["am"],
                        cm.GA["gam"], variant)
    R, V = api.get_R(4), api.get_Val(4)
    eRm, eRc = cm.rel_err(R[:, :3], Ro[:, :3]), cm.rel_err(R[:, 3], Ro[:, 3])
    errs = cm.block_class_errs(V, Vo)
    cm.log_parity(f"C2 1M-tet assembly variant={variant}", R_mom=eRm, R_cont=eRc, **errs)
    assert eRm <= 1e-12 and eRc <= 1e-12
    assert max(errs.values()) <= 1e-12, errs
    if variant != api.ASM_ATOMIC:            # deterministic variants: bitwise repeatable at this size too
        api.CONSTRUCT_FLUID(p.Ag, p.Yg, None, cm.RHO, cm.MU, cm.F, cm.DT, cm.GA["af"], cm.GA["am"],
                            cm.GA["gam"], variant)
        assert np.array_equal(api.get_R(4), R) and np.array_equal(api.get_Val(4), V)


def test_c2_sparmulvv(c2):
    m, p, Ro, Vo = c2
    rng = np.random.default_rng(11)
    K = rng.standard_normal((p.colPtr.size, 16)); U = rng.standard_normal((p.rm.nNo, 4))
    w = cm.oracle_world([p], m.nNo, with_faces=False)
    ref = w.sparmul_vv(4, [K], [U])[0]
    got = api.FSILS_SPARMUL("VV", 4, K, U)
    e = cm.rel_err(got, ref)
    cm.log_parity("C2 1M-tet SPARMULVV dof=4", err=e)
    assert e <= 1e-14
    # the assembled matrix too (its entries span ten orders of magnitude, unlike the random one)
    ref = w.sparmul_vv(4, [Vo], [U])[0]
    got = api.FSILS_SPARMUL("VV", 4, Vo, U)
    e = cm.rel_err(got, ref)
    cm.log_parity("C2 1M-tet SPARMULVV dof=4 (assembled Val)", err=e)
    assert e <= 1e-14


@pytest.mark.parametrize("res_out", [0.0, 0.7])
def test_c2_gmres_newton_step(c2, res_out):
    """assembly -> FSILS_SOLVE(GMRES sD=50, relTol=1e-3: the benchmark's settings) at 1M tets"""
    m, p, Ro, Vo = c2
    w = cm.oracle_world([p], m.nNo)
    ls_o = ora.ls_create(ora.LS_TYPE_GMRES, **LS)
    Xo = Ro.copy()
    w.solve(ls_o, 4, [Xo], [Vo.copy()], incL=[1, 1, 1], res=[0.0, 0.0, res_out])
    api.CONSTRUCT_FLUID(p.Ag, p.Yg, None, cm.RHO, cm.MU, cm.F, cm.DT, cm.GA["af"], cm.GA["am"],
                        cm.GA["gam"], api.ASM_GATHER)
    ls = api.FSILS_LS_CREATE(api.LS_TYPE_GMRES, **LS)
    api.solve_dev(ls, 4, incL=[1, 1, 1], res=[0.0, 0.0, res_out])
    X = api.get_R(4)
    num = float(np.linalg.norm(X - Xo) / np.linalg.norm(Xo))
    cm.log_parity(f"C2 1M-tet GMRES(sD=50, relTol=1e-3) res={res_out}", itr_gpu=int(ls.RI.itr),
                  itr_oracle=int(ls_o.RI.itr), iNorm_rel=abs(ls.RI.iNorm - ls_o.RI.iNorm) / ls_o.RI.iNorm,
                  fNorm_rel=abs(ls.RI.fNorm - ls_o.RI.fNorm) / ls_o.RI.fNorm, step=num)
    assert abs(ls.RI.itr - ls_o.RI.itr) <= 1, (ls.RI.itr, ls_o.RI.itr)
    assert bool(ls.RI.suc) == bool(ls_o.RI.suc)
    assert abs(ls.RI.iNorm - ls_o.RI.iNorm) <= 1e-10 * ls_o.RI.iNorm
    if ls.RI.itr == ls_o.RI.itr:
        assert num <= 1e-8, num


def test_c2_nssolver_newton_step(c2):
    """FSILS_NSSOLVER with the FSILS defaults (svFSI's default for fluid, L/LS.f:70-78) at 1M tets"""
    m, p, Ro, Vo = c2
    w = cm.oracle_world([p], m.nNo)
    ls_o = ora.ls_create(ora.LS_TYPE_NS)
    Xo = Ro.copy()
    w.solve(ls_o, 4, [Xo], [Vo.copy()], incL=[1, 1, 1], res=[0.0, 0.0, 0.0])
    api.CONSTRUCT_FLUID(p.Ag, p.Yg, None, cm.RHO, cm.MU, cm.F, cm.DT, cm.GA["af"], cm.GA["am"],
                        cm.GA["gam"], api.ASM_GATHER)
    ls = api.FSILS_LS_CREATE(api.LS_TYPE_NS)
    api.solve_dev(ls, 4, incL=[1, 1, 1], res=[0.0, 0.0, 0.0])
    X = api.get_R(4)
    num = float(np.linalg.norm(X - Xo) / np.linalg.norm(Xo))
    cm.log_parity("C2 1M-tet NSSOLVER (FSILS defaults)", RI=f"{ls.RI.itr}/{ls_o.RI.itr}",
                  GM=f"{ls.GM.itr}/{ls_o.GM.itr}", CG=f"{ls.CG.itr}/{ls_o.CG.itr}", step=num)
    assert ls.RI.itr == ls_o.RI.itr and bool(ls.RI.suc) == bool(ls_o.RI.suc)
    assert abs(ls.GM.itr - ls_o.GM.itr) <= 1 and abs(ls.CG.itr - ls_o.CG.itr) <= 1
    assert abs(ls.RI.iNorm - ls_o.RI.iNorm) <= 1e-10 * ls_o.RI.iNorm
    if (ls.GM.itr, ls.CG.itr) == (ls_o.GM.itr, ls_o.CG.itr):
        assert num <= 1e-8, num
