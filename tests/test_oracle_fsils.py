"""CPU tests of the oracle's FSILS restatement (oracle/ora_fsils.c): independent algebraic checks
(SciPy BSR SpMV, true residuals of the solves, partition independence, dense GE vs LAPACK) and the
committed 2-rank golden fixture.  No GPU."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import common as cm
from oracle import oracle as ora
from svfsi_b200 import mesh

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pipe():
    m, probs, _ = mesh.build_problem(6, 6, 10, nparts=1, L=3.0)
    Rs, Vs = cm.oracle_assemble(probs)
    return m, probs[0], Rs[0], Vs[0]


def bsr(p, V, dof):
    return sp.bsr_matrix((V.reshape(-1, dof, dof), p.colPtr - 1, p.rowPtr - 1),
                         shape=(p.rm.nNo * dof, p.rm.nNo * dof)).tocsr()


def test_sparmul_shapes_vs_scipy(pipe):
    m, p, R, V = pipe
    rng = np.random.default_rng(0)
    w = cm.oracle_world([p], m.nNo, with_faces=False)
    nnz, n = p.colPtr.size, p.rm.nNo
    K = rng.standard_normal((nnz, 16)); U = rng.standard_normal((n, 4))
    ref = (bsr(p, K, 4) @ U.reshape(-1)).reshape(n, 4)
    assert cm.rel_err(w.sparmul_vv(4, [K], [U])[0], ref) < 1e-14
    K3 = rng.standard_normal((nnz, 3)); U3 = rng.standard_normal((n, 3)); s = rng.standard_normal(n)
    A = sp.csr_matrix
    rows = np.repeat(np.arange(n), np.diff(p.rowPtr))
    D = sp.csr_matrix((K3.reshape(-1), (np.repeat(rows, 3), (np.repeat((p.colPtr - 1) * 3, 3) + np.tile(np.arange(3), nnz)))),
                      shape=(n, 3 * n))
    assert cm.rel_err(w.sparmul_vs(3, [K3], [U3])[0], D @ U3.reshape(-1)) < 1e-14
    G = sp.csr_matrix((K3.reshape(-1), (np.repeat(rows * 3, 3) + np.tile(np.arange(3), nnz), np.repeat(p.colPtr - 1, 3))),
                      shape=(3 * n, n))
    assert cm.rel_err(w.sparmul_sv(3, [K3], [s])[0].reshape(-1), G @ s) < 1e-14
    k1 = rng.standard_normal(nnz)
    S = sp.csr_matrix((k1, p.colPtr - 1, p.rowPtr - 1), shape=(n, n))
    assert cm.rel_err(w.sparmul_ss([k1], [s])[0], S @ s) < 1e-14


def _free_mask(p, dof=4):
    free = np.ones((p.rm.nNo, dof), bool)
    for name in ("inlet", "wall"):
        free[p.faces[name]["gN"] - 1, :3] = False
    return free


@pytest.mark.parametrize("ls_type,kw", [
    (ora.LS_TYPE_GMRES, dict(relTol=1e-6, maxItr=10, dimKry=80)),
    (ora.LS_TYPE_NS, dict(relTol=1e-5, maxItr=15, dimKry=60, relTolIn=(1e-3, 1e-2))),
])
def test_fluid_solves_reduce_the_true_residual(pipe, ls_type, kw):
    """x returned by FSILS_SOLVE solves K x = R on the free dofs (Dirichlet dofs get x = 0):
    independent check of PRECONDDIAG + Krylov + un-scaling through the UNscaled matrix."""
    m, p, R, V = pipe
    w = cm.oracle_world([p], m.nNo)
    ls = ora.ls_create(ls_type, absTol=1e-14, **kw)
    X = R.copy()
    w.solve(ls, 4, [X], [V.copy()], incL=[1, 1, 1], res=[0.0, 0.0, 0.0])
    assert ls.RI.suc
    free = _free_mask(p)
    assert np.abs(X[~free]).max() == 0.0
    K = bsr(p, V, 4)
    r = (R.reshape(-1) - K @ X.reshape(-1)).reshape(-1, 4)
    # Jacobi-scaled residual norm, the quantity FSILS controls
    d = np.abs(K.diagonal()).reshape(-1, 4); d[d == 0] = 1.0
    W = 1 / np.sqrt(d)
    rel = np.linalg.norm((W * r)[free]) / np.linalg.norm((W * R)[free])
    assert rel < 20 * kw["relTol"], rel


def test_coupled_resistance_face_adds_rank_one_term(pipe):
    """with res != 0 the operator is K + res * v v^T (ADDBCMUL): check through the true residual"""
    m, p, R, V = pipe
    w = cm.oracle_world([p], m.nNo)
    res = 4.0
    ls = ora.ls_create(ora.LS_TYPE_GMRES, relTol=1e-7, absTol=1e-14, maxItr=10, dimKry=100)
    X = R.copy()
    w.solve(ls, 4, [X], [V.copy()], incL=[1, 1, 1], res=[0.0, 0.0, res])
    free = _free_mask(p)
    v = np.zeros((p.rm.nNo, 4)); v[p.faces["outlet"]["gN"] - 1, :3] = p.faces["outlet"]["val"]
    v[~free] = 0.0     # valM = val * W with W = 0 on Dirichlet dofs
    K = bsr(p, V, 4)
    r = (R.reshape(-1) - K @ X.reshape(-1) - res * v.reshape(-1) * (v.reshape(-1) @ X.reshape(-1))).reshape(-1, 4)
    d = np.abs(K.diagonal()).reshape(-1, 4); d[d == 0] = 1.0
    W = 1 / np.sqrt(d)
    rel = np.linalg.norm((W * r)[free]) / np.linalg.norm((W * R)[free])
    assert rel < 1e-5, rel


def test_heat_cg_solves_spd_system():
    m, probs, _ = mesh.build_problem(6, 6, 8, nparts=1, L=2.0)
    p = probs[0]
    rng = np.random.default_rng(1)
    par = ora.heat_par(1.0, 0.5, 1.0, 1e-2, cm.GA["af"], cm.GA["am"], cm.GA["gam"])
    R, V = ora.construct_heats(par, p.rm.IEN, p.rm.x, rng.standard_normal(p.rm.nNo),
                               rng.standard_normal(p.rm.nNo), p.rowPtr, p.colPtr)
    w = ora.World(m.nNo, [p.rm.ltg], [p.rowPtr], [p.colPtr], 1)
    w.bc_create(1, [p.faces["inlet"]["gN"]], 1, ora.BC_TYPE_Dir, None)
    ls = ora.ls_create(ora.LS_TYPE_CG, relTol=1e-10, absTol=1e-14, maxItr=500)
    X = R.copy()
    w.solve(ls, 1, [X], [V.copy()], incL=[1], res=None)
    assert ls.RI.suc
    A = sp.csr_matrix((V, p.colPtr - 1, p.rowPtr - 1), shape=(p.rm.nNo,) * 2)
    free = np.ones(p.rm.nNo, bool); free[p.faces["inlet"]["gN"] - 1] = False
    xs = np.zeros(p.rm.nNo)
    import scipy.sparse.linalg as spl
    xs[free] = spl.spsolve(A[free][:, free].tocsc(), R[free])
    assert np.linalg.norm(X - xs) / np.linalg.norm(xs) < 1e-8


@pytest.mark.parametrize("nparts", [2, 3, 4])
def test_partition_independence(nparts):
    """k simulated ranks == 1 rank: LHS reorder, halo sums, owned-node dots, GMRES / NS / assembly"""
    dims, L = (4, 4, 12), 3.0
    m1, p1, _ = mesh.build_problem(*dims, nparts=1, L=L)
    mk, pk, _ = mesh.build_problem(*dims, nparts=nparts, L=L)
    R1, V1 = cm.oracle_assemble(p1)
    Rk, Vk = cm.oracle_assemble(pk)
    wk = cm.oracle_world(pk, mk.nNo)
    # every rank's reordering is a permutation with [shared-lower | interior | shared-higher]
    owned = np.zeros(mk.nNo, int)
    for r, p in enumerate(pk):
        info = wk.info(r); mp = wk.map(r)
        assert sorted(mp) == list(range(1, p.rm.nNo + 1))
        owned[p.rm.ltg[np.argsort(mp)][: info["mynNo"]] - 1] += 1
    assert (owned == 1).all()          # every global node is owned by exactly one rank
    # assembled + halo-summed residual equals the 1-rank residual
    Rc = cm.oracle_commu(wk, pk, Rk)
    Rg = np.zeros((m1.nNo, 4)); Rg[p1[0].rm.ltg - 1] = R1[0]
    for p, r in zip(pk, Rc):
        assert cm.rel_err(r, Rg[p.rm.ltg - 1]) < 1e-13
    for lst, kw in ((ora.LS_TYPE_GMRES, dict(relTol=1e-4, sD=60, mItr=10)),
                    (ora.LS_TYPE_NS, dict(relTol=1e-3, sD=60, mItr=10))):
        l1, G1 = cm.oracle_gmres_global(1, res_out=2.0, dims=dims, L=L, ls_type=lst, **kw)
        lk, Gk = cm.oracle_gmres_global(nparts, res_out=2.0, dims=dims, L=L, ls_type=lst, **kw)
        assert abs(l1.RI.itr - lk.RI.itr) <= 1
        assert np.linalg.norm(G1 - Gk) / np.linalg.norm(G1) < 1e-6


def test_ge_matches_lapack():
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 7, 20):
        M = rng.standard_normal((n, n)); A = M @ M.T + n * np.eye(n)
        b = rng.standard_normal(n)
        ok, x = ora.ge(A, b)
        assert ok and np.allclose(x, np.linalg.solve(A, b), rtol=1e-10)
    ok, x = ora.ge(np.zeros((2, 2)), np.ones(2))
    assert not ok and (x == 0).all()


def test_ls_defaults():
    ns = ora.ls_create(ora.LS_TYPE_NS)
    assert (ns.RI.relTol, ns.GM.relTol, ns.CG.relTol) == (0.4, 1e-2, 0.2)
    assert (ns.RI.mItr, ns.GM.mItr, ns.CG.mItr, ns.GM.sD) == (10, 2, 500, 100)
    gm = ora.ls_create(ora.LS_TYPE_GMRES)
    assert (gm.RI.relTol, gm.RI.mItr, gm.RI.sD, gm.RI.absTol) == (0.1, 4, 250, 1e-10)
    cg = ora.ls_create(ora.LS_TYPE_CG)
    assert (cg.RI.relTol, cg.RI.mItr) == (1e-2, 1000)


def test_golden_two_rank_pipe():
    """regression pin (fixture generated by tests/golden/make_golden.py from the oracle)"""
    g = np.load(os.path.join(GOLD, "pipe_2rank.npz"))
    m, probs, _ = mesh.build_problem(4, 4, 6, nparts=2, L=3.0)
    Rs, Vs = cm.oracle_assemble(probs)
    assert np.array_equal(Rs[0], g["R0"]) and np.array_equal(Vs[1], g["V1"])
    w = cm.oracle_world(probs, m.nNo)
    Rc = cm.oracle_commu(w, probs, Rs)
    assert np.array_equal(Rc[0], g["Rc0"]) and np.array_equal(Rc[1], g["Rc1"])
    for tag, lst in (("gmres", ora.LS_TYPE_GMRES), ("ns", ora.LS_TYPE_NS)):
        ls = ora.ls_create(lst, relTol=1e-8, absTol=1e-14, maxItr=20 if lst == ora.LS_TYPE_NS else 6,
                           dimKry=60)
        X = [r.copy() for r in Rc]
        w.solve(ls, 4, X, [v.copy() for v in Vs], incL=[1, 1, 1], res=[0.0, 0.0, 3.0])
        st = g[f"{tag}_stats"]
        assert (ls.RI.itr, ls.RI.suc) == (int(st[0]), int(st[1]))
        assert np.allclose(X[0], g[f"{tag}_X0"], rtol=0, atol=1e-9 * np.abs(g[f"{tag}_X0"]).max())


# ---------------------------------------------------------------------------------------------
# BICGSV/BICGSS (L/BICGS.f) and PRECONDRCS (L/PRECOND.f:150-368)
def _rcs_reference(K, dir_mask, maxiter=10, tol=2.0):
    """independent NumPy/SciPy restatement of PRECONDRCS on the scalar CSR expansion of the matrix:
    returns (scaled matrix, W1, W2).  dir_mask[i] = True where dof i is a Dirichlet dof."""
    n = K.shape[0]
    keep = (~dir_mask).astype(float)
    A = sp.diags(keep) @ K @ sp.diags(keep)
    A = A.tolil()
    for i in np.nonzero(dir_mask)[0]:
        A[i, i] = 1.0
    A = A.tocsr()
    W1 = np.ones(n); W2 = np.ones(n)
    it, flag = 0, True
    while flag:
        it += 1
        if it >= maxiter:
            flag = False
        B = abs(A)
        wr = np.asarray(B.max(axis=1).todense()).reshape(-1)
        wc = np.asarray(B.max(axis=0).todense()).reshape(-1)
        if np.abs(1 - wr).max() < tol and np.abs(1 - wc).max() < tol:
            flag = False
        wr = 1 / np.sqrt(wr); wc = 1 / np.sqrt(wc)
        A = (sp.diags(wr) @ A @ sp.diags(wc)).tocsr()
        W1 *= wr; W2 *= wc
    return A, W1, W2


@pytest.mark.parametrize("ls_type,prec,kw", [
    (ora.LS_TYPE_BICGS, ora.PRECOND_FSILS, dict(relTol=1e-8, maxItr=400)),
    (ora.LS_TYPE_BICGS, ora.PRECOND_RCS, dict(relTol=1e-8, maxItr=400)),
    (ora.LS_TYPE_GMRES, ora.PRECOND_RCS, dict(relTol=1e-7, maxItr=10, dimKry=80)),
])
def test_bicgs_and_rcs_solve_the_unscaled_system(pipe, ls_type, prec, kw):
    """x returned by FSILS_SOLVE solves K x = R on the free dofs whatever the preconditioner:
    checked through the UNscaled matrix against a SciPy direct solve."""
    import scipy.sparse.linalg as spl
    m, p, R, V = pipe
    w = cm.oracle_world([p], m.nNo)
    ls = ora.ls_create(ls_type, absTol=1e-14, **kw)
    X = R.copy()
    Vs = V.copy()
    w.solve(ls, 4, [X], [Vs], prec=prec, incL=[1, 1, 1], res=[0.0, 0.0, 0.0])
    assert ls.RI.suc and ls.RI.itr > 3
    free = _free_mask(p).reshape(-1)
    assert np.abs(X.reshape(-1)[~free]).max() == 0.0
    K = bsr(p, V, 4)
    xs = np.zeros(free.size)
    xs[free] = spl.spsolve(K[free][:, free].tocsc(), R.reshape(-1)[free])
    assert np.linalg.norm(X.reshape(-1) - xs) / np.linalg.norm(xs) < 2e-5
    if prec == ora.PRECOND_RCS:
        # Val left behind = the equilibrated matrix: identity on the Dirichlet dofs, every row and
        # column max within (1/3, 3) unless the 10-sweep cap hit; equals the SciPy restatement
        A, W1, W2 = _rcs_reference(K, ~free)
        S = bsr(p, Vs, 4)
        # (the reference rewrites EVERY diagonal entry as Wr*(Val-1)+1, L/PRECOND.f:203-237: on free
        #  dofs that is (Val-1)+1, an absolute 1e-16 perturbation of diagonals that may be ~1e-4)
        assert abs(S - A).max() < 1e-10 * abs(A).max()


def test_bicgs_scalar_heat_matches_direct_solve():
    m, probs, _ = mesh.build_problem(6, 6, 8, nparts=1, L=2.0)
    p = probs[0]
    rng = np.random.default_rng(5)
    par = ora.heat_par(1.0, 0.5, 1.0, 1e-2, cm.GA["af"], cm.GA["am"], cm.GA["gam"])
    R, V = ora.construct_heats(par, p.rm.IEN, p.rm.x, rng.standard_normal(p.rm.nNo),
                               rng.standard_normal(p.rm.nNo), p.rowPtr, p.colPtr)
    import scipy.sparse.linalg as spl
    A = sp.csr_matrix((V, p.colPtr - 1, p.rowPtr - 1), shape=(p.rm.nNo,) * 2)
    free = np.ones(p.rm.nNo, bool); free[p.faces["inlet"]["gN"] - 1] = False
    xs = np.zeros(p.rm.nNo)
    xs[free] = spl.spsolve(A[free][:, free].tocsc(), R[free])
    for prec in (ora.PRECOND_FSILS, ora.PRECOND_RCS):
        w = ora.World(m.nNo, [p.rm.ltg], [p.rowPtr], [p.colPtr], 1)
        w.bc_create(1, [p.faces["inlet"]["gN"]], 1, ora.BC_TYPE_Dir, None)
        ls = ora.ls_create(ora.LS_TYPE_BICGS, relTol=1e-10, absTol=1e-14, maxItr=500)
        X = R.copy()
        w.solve(ls, 1, [X], [V.copy()], prec=prec, incL=[1], res=None)
        assert ls.RI.suc
        assert np.linalg.norm(X - xs) / np.linalg.norm(xs) < 1e-7


@pytest.mark.parametrize("nparts", [2, 3])
def test_bicgs_rcs_partition_independence(nparts):
    """k simulated ranks == 1 rank for BICGS + RCS.  Shared rows see the SUM of the per-rank row
    maxima (FSILS_COMMUV on Wr/Wc, L/PRECOND.f:320-321), so the scaling differs between
    partitions but the un-scaled solution does not."""
    dims, L = (4, 4, 12), 3.0
    out = []
    for k in (1, nparts):
        m, probs, _ = mesh.build_problem(*dims, nparts=k, L=L)
        Rs, Vs = cm.oracle_assemble(probs)
        w = cm.oracle_world(probs, m.nNo)
        Rc = cm.oracle_commu(w, probs, Rs)
        ls = ora.ls_create(ora.LS_TYPE_BICGS, relTol=1e-9, absTol=1e-14, maxItr=400)
        X = [r.copy() for r in Rc]
        w.solve(ls, 4, X, [v.copy() for v in Vs], prec=ora.PRECOND_RCS, incL=[1, 1, 1],
                res=np.array([0.0, 0.0, 0.0]))
        assert ls.RI.suc
        G = np.zeros((m.nNo, 4))
        for p, x in zip(probs, X):
            G[p.rm.ltg - 1] = x
        out.append(G)
    assert np.linalg.norm(out[0] - out[1]) / np.linalg.norm(out[0]) < 1e-5
