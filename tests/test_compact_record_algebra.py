"""CPU check of the ALGEBRA behind the CUDA element kernels, independent of any GPU:
for linear tetrahedra the shape-function gradients, the metric tensor and the velocity gradient are
element constants, so the four Gauss points of FLUID3D_M / FLUID3D_C (S/FLUID.f:192-560, :813-1084)
reduce to an 80-double compact record per element (svfsi_b200/csrc/asm_kernels.cu, `fluid_elem_compute` +
`fluid_elem_store`) from which each 4x4 tangent block (a,b) is expanded at gather time (`tangent_pair`;
the lane-constant-coefficient form of the lean / wide / quad kernels with sum_g N_a(g) = 1).

This file restates those two steps in NumPy, statement by statement, and compares them with the oracle's
Gauss-point loop (oracle/ora_elem.c, the restatement of the Fortran) on random elements: the
reorganisation is exact up to rounding (<= 1e-13 here; the CUDA kernels measure <= 4e-15 on the pipe)."""
import numpy as np
import pytest

import common as cm
from oracle import oracle as ora

S5 = (5.0 + 3.0 * np.sqrt(5.0)) / 20.0
T5 = (5.0 - np.sqrt(5.0)) / 20.0
EPS = np.finfo(float).eps


def gauss_N():
    """N[g][a], S/NN.f:268-275 (xi) and :654-658 (N4 = 1 - xi1 - xi2 - xi3)"""
    N = np.full((4, 4), T5)
    for g in range(3):
        N[g, g] = S5
    for g in range(4):
        N[g, 3] = 1.0 - N[g, 0] - N[g, 1] - N[g, 2]
    return N


def compact_record(rho, mu, f, dt, af, am, gam, xl, al, yl, bfl):
    """fluid_elem_compute + fluid_elem_store of asm_kernels.cu"""
    X = (xl[:3] - xl[3]).T                       # X[r][c] = xl[c][r] - xl[3][r]
    Jac = np.linalg.det(X)
    XI = np.linalg.inv(X)                        # xiX = adj / Jac, S/NN.f:1534-1546
    ks = XI.T @ XI
    Nx = np.vstack([XI, -XI.sum(axis=0)])        # Nx[a][c]
    T1 = af * gam * dt
    amd = am / T1
    w = Jac / 24.0
    wl, wr = w * T1, w * rho
    ald = al[:, :3] - bfl                        # ud uses al - bfl, S/FLUID.f:236-238
    ux = Nx.T @ yl[:, :3]                        # ux[j][i] = d u_i / d x_j
    px = Nx.T @ yl[:, 3]
    divU = np.trace(ux)
    es = ux + ux.T
    kT = 4.0 / dt ** 2
    kS = 36.0 * (ks * ks).sum() * (mu / rho) ** 2
    trks = np.trace(ks)
    N = gauss_N()
    A = np.zeros((4, 4)); c2 = np.zeros(4); r2 = np.zeros(4); lR4 = np.zeros(4)
    sNrV = np.zeros((4, 3)); sRM = np.zeros((3, 3)); sTC = sTM = 0.0
    for g in range(4):
        Ng = N[g]
        ud = -np.asarray(f) + Ng @ ald
        u = Ng @ yl[:, :3]
        p = Ng @ yl[:, 3]
        kU = u @ ks @ u
        tauM = 1.0 / (rho * np.sqrt(kT + kU + kS))
        rV = ud + u @ ux
        up = -tauM * (rho * rV + px)
        tauC = 1.0 / (tauM * trks)
        tB = up @ ks @ up
        if abs(tB) / max(abs(tB), EPS) < 10 * EPS:
            tB = EPS
        tauB = rho / np.sqrt(tB)
        ua = u + up
        pa = p - tauC * divU
        rVb = tauB * (up @ ux)
        # rM(j,i), S/FLUID.f:427-439
        sRM += mu * es - rho * np.outer(ua, up) + np.outer(up, rVb) - pa * np.eye(3)
        rV2 = ud + ua @ ux
        uNx, upNx = Nx @ u, Nx @ up
        uaNx = uNx + upNx
        sNrV += np.outer(Ng, rV2)
        lR4 += Ng * divU - upNx                  # S/FLUID.f:1046-1049
        c2 += tauM * uaNx
        r2 += tauM * (uNx + amd * Ng)
        sTC += tauC; sTM += tauM
        for a in range(4):
            for b in range(4):
                A[a, b] += (rho * amd * Ng[b] * (Ng[a] + rho * tauM * uaNx[a]) + rho * Ng[a] * (uNx[b] + upNx[b])
                            + tauB * upNx[a] * upNx[b] + rho * tauM * uaNx[a] * (rho * uNx[b]))
    nn = Nx @ Nx.T
    rec = dict(Nx=Nx, C2=rho * c2, R2=rho * r2, sTC=sTC, sTM=sTM, wl=wl,
               D=4.0 * mu * nn + A, E=sTM * nn)
    lR = np.zeros((4, 4))
    lR[:, :3] = wr * sNrV + w * (Nx @ sRM)
    lR[:, 3] = w * lR4
    rec["lR"] = lR
    return rec


def expand_block(rec, a, b, mu, exact_sN=False):
    """tangent_pair of asm_kernels.cu: the 4x4 block (row node a, column node b), row i, column j.
    exact_sN: sum_g N_a(g) taken as 1 (the lean / wide / quad kernels)."""
    Nx, wl, sTC = rec["Nx"], rec["wl"], rec["sTC"]
    sN = np.ones(4) if exact_sN else gauss_N().sum(axis=0)
    mu4 = 4.0 * mu
    K = np.zeros((4, 4))
    for i in range(3):
        for j in range(3):
            K[i, j] = wl * (mu4 * (Nx[a, j] * Nx[b, i]) + sTC * (Nx[a, i] * Nx[b, j]) + (rec["D"][a, b] if i == j else 0.0))
        K[i, 3] = -wl * (Nx[a, i] * sN[b] - Nx[b, i] * rec["C2"][a])
    for j in range(3):
        K[3, j] = wl * (sN[a] * Nx[b, j] + Nx[a, j] * rec["R2"][b])
    K[3, 3] = wl * rec["E"][a, b]
    return K


def _elements(n, seed):
    rng = np.random.default_rng(seed)
    base = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0.0]])
    for _ in range(n):
        xl = base * rng.uniform(0.05, 2.0) + 0.1 * rng.standard_normal((4, 3))
        if np.linalg.det((xl[:3] - xl[3]).T) < 0:
            xl[[0, 1]] = xl[[1, 0]]
        al = rng.standard_normal((4, 4)); al[:, 3] = 0.0
        yl = 10.0 * rng.standard_normal((4, 4))
        yield xl, al, yl, rng.standard_normal((4, 3))


@pytest.mark.parametrize("exact_sN", [False, True])
def test_compact_record_expansion_equals_the_gauss_point_loop(exact_sN):
    ga = cm.GA
    rho, mu, f, dt = 1.06, 0.04, (0.1, -0.2, 0.3), 5e-3
    par = ora.fluid_par(rho, mu, f, dt, ga["af"], ga["am"], ga["gam"])
    worst = 0.0
    for xl, al, yl, bfl in _elements(24, 7):
        lR, lK, flag = ora.fluid_element(par, xl, al, yl, bfl)
        assert flag == 0
        rec = compact_record(rho, mu, f, dt, ga["af"], ga["am"], ga["gam"], xl, al, yl, bfl)
        assert np.abs(rec["lR"] - lR).max() <= 1e-13 * np.abs(lR).max()
        ref = np.abs(lK).max()
        for a in range(4):
            for b in range(4):
                K = expand_block(rec, a, b, mu, exact_sN)
                Ko = lK[b, a].reshape(4, 4)            # lK(16, a, b): block (row node a, col node b), entry i*4+j
                worst = max(worst, np.abs(K - Ko).max() / ref)
    assert worst <= 1e-13, worst


def test_blocks_ab_and_ba_share_their_operands():
    """what the pair-owner kernel uses: block (b,a) is block (a,b) with the two nodes' roles exchanged,
    so one fetch of (Nx_a, Nx_b, C2, R2, sum tauC, wl) serves both; only (D,E) differs"""
    ga = cm.GA
    rho, mu, f, dt = 1.06, 0.04, (0.0, 0.0, 0.0), 5e-3
    xl, al, yl, bfl = next(_elements(1, 3))
    rec = compact_record(rho, mu, f, dt, ga["af"], ga["am"], ga["gam"], xl, al, yl, bfl)
    swapped = dict(rec)
    perm = [1, 0, 2, 3]
    for k in ("Nx", "C2", "R2"):
        swapped[k] = rec[k][perm]
    swapped["D"] = rec["D"][np.ix_(perm, perm)]; swapped["E"] = rec["E"][np.ix_(perm, perm)]
    assert np.array_equal(expand_block(rec, 1, 0, mu, True), expand_block(swapped, 0, 1, mu, True))
    assert not np.allclose(rec["D"][0, 1], rec["D"][1, 0])      # D_ab != D_ba: advection is not symmetric
    assert np.allclose(rec["E"], rec["E"].T)
