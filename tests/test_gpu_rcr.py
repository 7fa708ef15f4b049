"""-m gpu: the device-resident time loop with an RCR (Windkessel) outlet -- the shape of BASELINE
configs[0], 04-fluid/01-pipe3D_RCR -- against the oracle's loop (`cm.oracle_rcr_time_loop`).  The 0-D
model stays on the host (`svfsi_b200/cplbc.py`, as in the reference); the device supplies the outlet flux
(`gpu_face_integ_v_`) and consumes the face pressure (`gpu_bassem_neu_fluid_`) and the resistance
(`res` of `gpu_solve_dev_` -> ADDBCMUL).

First run on a B200 in round 2 (gpurun_out/r02s1_pytest_rcr.log: passed); the oracle loop it is compared
with integrates the 0-D model with the oracle's own restatement (tests/common.py::OracleCplBC)."""

import numpy as np
import pytest

import common as cm
from svfsi_b200 import api, cplbc, mesh

pytestmark = pytest.mark.gpu


def test_device_resident_rcr_time_loop(gpu_lib):
    m, probs, _ = mesh.build_problem(8, 8, 20, nparts=1, L=4.0)
    p = probs[0]
    ga = cm.GA
    nNo = p.rm.nNo
    rcr = [cplbc.RCR(121.0, 1.5e-4, 1212.0)]
    nsteps, nnewton = 2, 3
    o_hist, (oA, oY), o_cpl = cm.oracle_rcr_time_loop(m, p, rcr, nsteps, nnewton)

    api.FSILS_LHS_CREATE(m.nNo, nNo, p.colPtr.size, p.rm.ltg, p.rowPtr, p.colPtr, 3)
    try:
        for fi, name in enumerate(cm.FACE_ORDER, start=1):
            fa = p.faces[name]
            api.FSILS_BC_CREATE(fi, fa["gN"].size, 3,
                                api.BC_TYPE_Neu if fa["bc"] == "Neu" else api.BC_TYPE_Dir, fa["gN"], fa["val"])
        api.mesh_create(p.rm.IEN, p.rm.x)
        # same initial state and Dirichlet data as the oracle loop
        rng = np.random.default_rng(21)
        Ao = 0.05 * rng.standard_normal((nNo, 4)); Ao[:, 3] = 0.0
        Yo = p.Yg.copy()
        gin = p.faces["inlet"]["gN"]; gw = p.faces["wall"]["gN"]
        xin = p.rm.x[gin - 1]
        r2 = (xin[:, 0] ** 2 + xin[:, 1] ** 2) / (np.abs(p.rm.x[:, :2]).max() ** 2)
        gx = np.clip(1.0 - r2, 0.0, None)
        nV = np.tile(np.array([0.0, 0.0, -1.0]), (gin.size, 1))
        from oracle import oracle as ora
        tA_in, tY_in = ora.setbcdirl(-12.0, gx, nV, 3)
        tA_w, tY_w = np.zeros((gw.size, 3)), np.zeros((gw.size, 3))
        gout, fIEN, gE = cm.local_face(m, p.rm, "outlet")

        eq = api.EqState(tol=1e-30, maxItr=nnewton)
        api.pic_init(4, Ao, Yo)
        api.face_create(3, gout, fIEN, gE)
        cpl = cplbc.CplBC(rcr, cm.DT, "SI")
        # flux of Yo: right after pic_init the integrator's Yn holds Yo; afterwards Yo of a step is the
        # final Yn of the previous one, so the host keeps the last flux
        q = dict(o=api.IntegV(3, which=1, s=1))

        def integ(i, which):
            return q["o"] if which == "o" else api.IntegV(3, which=1, s=1)
        cpl.init(integ)
        g_hist = []
        time = 0.0
        for ts in range(nsteps):
            time += cm.DT
            api.PICP(ga["gam"])
            api.SETBCDIR(gin, 1, tA_in, tY_in)
            api.SETBCDIR(gw, 1, tA_w, tY_w)
            while True:
                g = cpl.setbccpl(integ, time)[0]
                api.PICI(eq, ga["am"], ga["af"])
                api.construct_fluid_dev(cm.RHO, cm.MU, cm.F, cm.DT, ga["af"], ga["am"], ga["gam"],
                                        api.ASM_GATHER)
                api.BASSEMNEUBC_FLUID(3, np.full(gout.size, -g * 1.0), cm.RHO, 0.2, ga["af"], ga["gam"], cm.DT)
                api.commu_dev(4)
                ls = api.FSILS_LS_CREATE(api.LS_TYPE_GMRES, relTol=1e-5, absTol=1e-14, maxItr=10, dimKry=80)
                api.solve_dev(ls, 4, incL=[1, 1, 1], res=[0.0, 0.0, ga["gam"] * cm.DT * cpl.r[0]])
                g_hist.append((ls.RI.iNorm, ls.RI.itr, g, cpl.Qn[0]))
                if api.PICC(eq, ls, ga["gam"], ga["beta"], cm.DT):
                    break
            q["o"] = api.IntegV(3, which=1, s=1)
            api.pic_advance(eq)
            cpl.advance()
        gA, gY = api.pic_get(0, 4, nNo)
        api.face_free(3)
    finally:
        api.FSILS_LHS_FREE()

    assert len(g_hist) == len(o_hist)
    assert abs(cpl.r[0] - o_cpl.r[0]) <= 1e-9 * o_cpl.r[0]
    for (gi, gitr, gg, gq), (oi, oitr, og, oq) in zip(g_hist, o_hist):
        assert abs(gitr - oitr) <= 1
        assert abs(gi - oi) <= 1e-7 * max(oi, 1e-12 * o_hist[0][0]) + 1e-9 * o_hist[0][0]
        assert abs(gq - oq) <= 1e-8 * abs(oq) and abs(gg - og) <= 1e-8 * abs(og)
    assert np.linalg.norm(gY - oY) / np.linalg.norm(oY) <= 1e-8
