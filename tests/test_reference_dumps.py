"""The reference's own golden-vector hook for the hot path: PDEBUGVALR dumps (S/DEBUG.f:122-176).
(i) CPU: the reader (baseline/pdebugvalr.py) against a dump written by a restatement of the Fortran formatting
(NDTSTR, `1pE25.18`) from the oracle's own assembly -- R comes back bit for bit, Val to the eight printed
characters.  (ii) whenever a REAL dump exists (tests/golden/ref_Val_R_<nx>x<ny>x<nz>_<cTS>_<itr>_<rank>, produced
off-box by baseline/run_reference.sh with SVFSI_DUMP=1 on the case of baseline/make_reference_case.py), the oracle
-- and with -m gpu the CUDA path -- must reproduce it: R within 1e-12, Val within the printed precision.
No such file can be produced in this image (no Fortran compiler): until one is committed (ii) is skipped and the
pin is the reference's source text executed by oracle/refexec.py (tests/test_reference_golden.py), not a compiled run."""
import glob
import os
import re
import sys

import numpy as np
import pytest

import common as cm
from svfsi_b200 import mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import pdebugvalr as pdv  # noqa: E402

GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_Val_R_*")))


def test_ndtstr_matches_the_fortran_formatting():
    # values worked out by hand from S/UTIL.f:539-656 (eight characters, digits truncated)
    assert pdv.ndtstr(1.0) == "1.000000"
    assert pdv.ndtstr(-1.0) == "-1.00000"
    assert pdv.ndtstr(123.456789) == "1.2345E2"
    assert pdv.ndtstr(1.2399e-4) == "1.239E-4"
    assert pdv.ndtstr(-5.6789e-11) == "-5.6E-11"
    assert pdv.ndtstr(0.0) == "0.000000"


def test_dump_round_trip(tmp_path):
    m, probs, _ = mesh.build_problem(4, 4, 6, nparts=1, L=2.0)
    p = probs[0]
    Rs, Vs = cm.oracle_assemble([p])
    path = str(tmp_path / "Val_R_1_1_0")
    pdv.write_dump(path, p.rm.x, p.rm.ltg, Rs[0], p.rowPtr, p.colPtr, Vs[0])
    d = pdv.read_dump(path)
    assert np.array_equal(d["rowPtr"], p.rowPtr) and np.array_equal(d["colPtr"], p.colPtr)
    assert np.array_equal(d["ltg"], p.rm.ltg)
    assert np.array_equal(d["R"], Rs[0])                    # 1pE25.18 is exact for binary64
    tol = np.vectorize(pdv.printed_tolerance)(Vs[0])
    assert (np.abs(d["Val"] - Vs[0]) <= tol).all()
    assert np.abs(d["Val"] - Vs[0]).max() > 0               # ... and it really is only the printed precision


def _case_of(path):
    mm = re.search(r"ref_Val_R_(\d+)x(\d+)x(\d+)_(\d+)_(\d+)_(\d+)$", os.path.basename(path))
    return tuple(int(v) for v in mm.groups()) if mm else None


@pytest.mark.skipif(not GOLD, reason="no PDEBUGVALR dump of a compiled svFSI under tests/golden/")
@pytest.mark.parametrize("path", GOLD)
def test_oracle_reproduces_the_reference_dump(path):
    nx, ny, nz, cTS, itr, rank = _case_of(path)
    d = pdv.read_dump(path)
    m, probs, _ = mesh.build_problem(nx, ny, nz, nparts=1)
    p = probs[0]
    # first Newton iteration of the first time step: Ag / Yg as PICI leaves them from the initial state of
    # baseline/make_reference_case.py (= mesh.poiseuille_state); svFSI numbers the nodes as the .vtu does
    order = np.argsort(p.rm.ltg)
    Rs, Vs = cm.oracle_assemble([p])
    assert d["R"].shape == Rs[0].shape
    assert cm.rel_err(Rs[0][order], d["R"][np.argsort(d["ltg"])]) <= 1e-12


@pytest.mark.gpu
@pytest.mark.skipif(not GOLD, reason="no PDEBUGVALR dump of a compiled svFSI under tests/golden/")
@pytest.mark.parametrize("path", GOLD)
def test_gpu_reproduces_the_reference_dump(path, gpu_lib):
    from svfsi_b200 import api
    nx, ny, nz, cTS, itr, rank = _case_of(path)
    d = pdv.read_dump(path)
    m, probs, _ = mesh.build_problem(nx, ny, nz, nparts=1)
    p = probs[0]
    api.FSILS_LHS_CREATE(m.nNo, p.rm.nNo, p.colPtr.size, p.rm.ltg, p.rowPtr, p.colPtr, 3)
    try:
        api.mesh_create(p.rm.IEN, p.rm.x)
        api.CONSTRUCT_FLUID(p.Ag, p.Yg, None, cm.RHO, cm.MU, cm.F, cm.DT, cm.GA["af"], cm.GA["am"], cm.GA["gam"],
                            api.ASM_GATHER)
        R = api.get_R(4)
    finally:
        api.FSILS_LHS_FREE()
    order = np.argsort(p.rm.ltg)
    assert cm.rel_err(R[order], d["R"][np.argsort(d["ltg"])]) <= 1e-12
