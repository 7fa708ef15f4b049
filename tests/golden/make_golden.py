"""Generates the committed regression fixtures in tests/golden/ FROM THE ORACLE ITSELF.

The reference (svFSI) ships no golden vectors and cannot be compiled in this image (no Fortran
compiler / MPI), so these fixtures do not pin the oracle to the reference -- they pin the oracle
(and through it the CUDA path) against silent drift between rounds.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as cm  # noqa: E402
from oracle import oracle as ora  # noqa: E402
from svfsi_b200 import mesh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261017)
    ga = cm.GA
    parv = np.array([1.06, 0.04, 0.1, -0.2, 0.3, 5e-3, ga["af"], ga["am"], ga["gam"]])
    par = ora.fluid_par(parv[0], parv[1], tuple(parv[2:5]), *parv[5:9])
    n = 8
    xl = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0.0]])[None] * rng.uniform(0.05, 2, (n, 1, 1))
    xl = xl + 0.1 * rng.standard_normal((n, 4, 3)) * np.abs(xl).max(axis=(1, 2), keepdims=True)
    al = rng.standard_normal((n, 4, 4)); yl = 10 * rng.standard_normal((n, 4, 4))
    bfl = rng.standard_normal((n, 4, 3))
    lR = np.zeros((n, 4, 4)); lK = np.zeros((n, 4, 4, 16))
    for k in range(n):
        lR[k], lK[k], _ = ora.fluid_element(par, xl[k], al[k], yl[k], bfl[k])
    np.savez(os.path.join(HERE, "fluid_elements.npz"), par=parv, xl=xl, al=al, yl=yl, bfl=bfl,
             lR=lR, lK=lK)
    hparv = np.array([0.7, 0.3, 1.3, 5e-3, ga["af"], ga["am"], ga["gam"]])
    hp = ora.heat_par(*hparv)
    hal = rng.standard_normal((n, 4)); hyl = rng.standard_normal((n, 4))
    hR = np.zeros((n, 4)); hK = np.zeros((n, 4, 4))
    for k in range(n):
        hR[k], hK[k], _ = ora.heat_element(hp, xl[k], hal[k], hyl[k])
    np.savez(os.path.join(HERE, "heat_elements.npz"), par=hparv, xl=xl, al=hal, yl=hyl, lR=hR, lK=hK)

    # a small end-to-end system: 2-rank pipe, assembled + GMRES / NS / CG solves
    m, probs, _ = mesh.build_problem(4, 4, 6, nparts=2, L=3.0)
    Rs, Vs = cm.oracle_assemble(probs)
    w = cm.oracle_world(probs, m.nNo)
    Rc = cm.oracle_commu(w, probs, Rs)
    out = dict(R0=Rs[0], R1=Rs[1], V0=Vs[0], V1=Vs[1], Rc0=Rc[0], Rc1=Rc[1])
    for tag, lst in (("gmres", ora.LS_TYPE_GMRES), ("ns", ora.LS_TYPE_NS)):
        ls = ora.ls_create(lst, relTol=1e-8, absTol=1e-14, maxItr=20 if lst == ora.LS_TYPE_NS else 6,
                           dimKry=60)
        X = [r.copy() for r in Rc]
        w.solve(ls, 4, X, [v.copy() for v in Vs], incL=[1, 1, 1], res=[0.0, 0.0, 3.0])
        out[f"{tag}_X0"], out[f"{tag}_X1"] = X
        out[f"{tag}_stats"] = np.array([ls.RI.itr, ls.RI.suc, ls.RI.iNorm, ls.RI.fNorm, ls.GM.itr,
                                        ls.CG.itr])
    np.savez(os.path.join(HERE, "pipe_2rank.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
