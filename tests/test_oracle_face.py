"""CPU tests of the oracle's face integrals (oracle/oracle.py: gnnb_tri3, bassem_neu_fluid, integ_v --
S/NN.f:1856-1996, S/EQASSEM.f:90-192, S/FLUID.f:1279-1336, S/ALLFUN.f:199-262) through identities
that do not depend on the restatement: divergence theorem, disk area, traction resultants."""
import numpy as np
import pytest

import common as cm
from oracle import oracle as ora
from svfsi_b200 import mesh


@pytest.fixture(scope="module")
def pipe():
    m, probs, _ = mesh.build_problem(6, 6, 10, nparts=1, L=3.0)
    return m, probs[0]


def test_normals_are_outward_and_close_the_surface(pipe):
    m, p = pipe
    tot = np.zeros(3); area = {}
    for name in ("inlet", "wall", "outlet"):
        gN, fIEN, gE = cm.local_face(m, p.rm, name)
        n = ora.gnnb_tri3(p.rm.x, fIEN, p.rm.IEN, gE)
        tot += n.sum(axis=0) / 2.0               # |n| = 2 * area
        area[name] = np.linalg.norm(n, axis=1).sum() / 2.0
        c = p.rm.x[fIEN.astype(np.int64) - 1].mean(axis=1)
        if name == "outlet":
            assert (n[:, 2] > 0).all()
        elif name == "inlet":
            assert (n[:, 2] < 0).all()
        else:
            assert (np.einsum("ij,ij->i", n[:, :2], c[:, :2]) > 0).all()
    assert np.abs(tot).max() < 1e-12             # closed surface: sum of area vectors = 0
    assert abs(area["inlet"] - area["outlet"]) < 1e-12


def test_integ_v_divergence_theorem_and_linear_field(pipe):
    m, p = pipe
    x = p.rm.x
    # constant field: net flux through the closed boundary is zero; linear field u = (x, 2y, -0.5z):
    # div u = 2.5 -> total flux = 2.5 * volume (exact for P1 fields on a polyhedral domain)
    const = np.tile(np.array([0.3, -1.1, 0.7]), (p.rm.nNo, 1))
    lin = np.stack([x[:, 0], 2.0 * x[:, 1], -0.5 * x[:, 2]], axis=1)
    fc = fl = 0.0
    for name in ("inlet", "wall", "outlet"):
        gN, fIEN, gE = cm.local_face(m, p.rm, name)
        fc += ora.integ_v(x, p.rm.IEN, fIEN, gE, const)
        fl += ora.integ_v(x, p.rm.IEN, fIEN, gE, lin)
    X = x[p.rm.IEN.astype(np.int64) - 1]
    vol = np.abs(np.linalg.det(X[:, :3] - X[:, 3:4])).sum() / 6.0
    assert abs(fc) < 1e-11
    assert abs(fl - 2.5 * vol) < 1e-10 * vol


def test_bassem_neu_traction_resultant_and_backflow_tangent(pipe):
    m, p = pipe
    nNo = p.rm.nNo
    gN, fIEN, gE = cm.local_face(m, p.rm, "outlet")
    n = ora.gnnb_tri3(p.rm.x, fIEN, p.rm.IEN, gE)
    hg = np.zeros(nNo); hg[gN - 1] = -7.5                  # hg = -h*gx (S/SETBC.f:303-306)
    Y = np.zeros((nNo, 4)); Y[:, 2] = 3.0                  # pure outflow: no backflow term
    nnz = p.colPtr.size
    R = np.zeros((nNo, 4)); V = np.zeros((nnz, 16))
    ora.bassem_neu_fluid(p.rm.x, p.rm.IEN, fIEN, gE, hg, Y, p.rowPtr, p.colPtr, R, V, 1.06, 0.2,
                         cm.GA["af"], cm.GA["gam"], cm.DT)
    # sum_a lR(:,a) = -h * integral of n  (N sums to one); no tangent without backflow
    assert np.allclose(R[:, :3].sum(axis=0), 7.5 * n.sum(axis=0) / 2.0, rtol=1e-12, atol=1e-12)
    assert np.abs(R[:, 3]).max() == 0.0 and np.abs(V).max() == 0.0
    # reversed flow: udn = beta*rho*(u.n) < 0 -> residual gains udn*u, tangent -wl*N_a*N_b*udn > 0 on
    # the velocity diagonal only, symmetric in (a,b), total = -T1*udn*area
    Y[:, 2] = -3.0
    R2 = np.zeros((nNo, 4)); V2 = np.zeros((nnz, 16))
    ora.bassem_neu_fluid(p.rm.x, p.rm.IEN, fIEN, gE, hg, Y, p.rowPtr, p.colPtr, R2, V2, 1.06, 0.2,
                         cm.GA["af"], cm.GA["gam"], cm.DT)
    area = np.linalg.norm(n, axis=1).sum() / 2.0
    udn = 0.5 * 0.2 * 1.06 * (-3.0 - 3.0)
    T1 = cm.GA["af"] * cm.GA["gam"] * cm.DT
    B = V2.reshape(-1, 4, 4)
    assert np.allclose(B[:, 0, 0].sum(), -T1 * udn * area, rtol=1e-12)
    assert np.array_equal(B[:, 0, 0], B[:, 1, 1]) and np.array_equal(B[:, 0, 0], B[:, 2, 2])
    off = B.copy(); off[:, [0, 1, 2], [0, 1, 2]] = 0.0
    assert np.abs(off).max() == 0.0
    assert np.allclose((R2 - R)[:, 2].sum(), -udn * (-3.0) * area, rtol=1e-12)
