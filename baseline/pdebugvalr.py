#!/usr/bin/env python
"""Reader (and test-only writer) of the dumps svFSI's PDEBUGVALR writes (S/DEBUG.f:122-176): the assembled
residual R and block-CSR tangent Val of one rank at one Newton iteration, file `Val_R_<cTS>_<itr>_<rank>`.
This is the reference's own golden-vector hook for the hot path: R is printed with
`1pE25.18` (every bit), the Val entries and the column residuals through STR() = NDTSTR(x, 8)
(S/UTIL.f:528-656: EIGHT characters, digits truncated, e.g. `1.23E-04`, `-5.6E-11`), i.e. to 2-5 digits.

PDEBUGVALR is not called anywhere in the reference; to produce a dump insert `CALL PDEBUGVALR()` in
S/MAIN.f after `CALL COMMU(R)` (:161) -- baseline/run_reference.sh does it with SVFSI_DUMP=1 -- and run the
case of baseline/make_reference_case.py on ONE rank.  Drop the file as tests/golden/ref_Val_R_<nx>x<ny>x<nz>_
<cTS>_<itr>_<rank> and tests/test_reference_dumps.py compares the oracle (CPU) and the CUDA path (-m gpu)
against it: R at 1e-12, Val to the printed precision.

    python baseline/pdebugvalr.py Val_R_1_1_0        # summary of a dump
"""
import math
import re
import sys

import numpy as np

ROW = re.compile(r"^Row:\s*(\d+)\s+grow:\s*(\d+)\s+x:\s*(.*)$")
COL = re.compile(r"^\s*Col:\s*(\d+)\s+gcol:\s*(\d+)")


def _f(tok):
    # NDTSTR never writes a '+' and may write '*' when the value does not fit
    if "*" in tok:
        return float("nan")
    return float(tok)


def read_dump(path):
    """-> dict(x (n,3), ltg (n,), R (n,dof) full precision, rowPtr (n+1,), colPtr (nnz,), Val (nnz,dof*dof)
    low precision, Rcol (nnz,dof) low precision); ids 1-based as svFSI prints them"""
    xs, ltg, R, rowPtr, colPtr, Val = [], [], [], [1], [], []
    with open(path) as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    i, n = 0, len(lines)
    dof = None
    while i < n:
        ln = lines[i]
        m = ROW.match(ln.strip())
        if not m:
            i += 1
            continue
        ltg.append(int(m.group(2)))
        xs.append([_f(t) for t in m.group(3).split()])
        i += 1
        # "    R: v1" then one value per line until a dashed line / next row
        r = []
        first = lines[i].strip()
        assert first.startswith("R:"), f"{path}:{i + 1}: expected the R block"
        r.append(float(first[2:]))
        i += 1
        while i < n and not lines[i].strip().startswith("----") and not lines[i].startswith("="):
            r.append(float(lines[i]))
            i += 1
        dof = dof or len(r)
        R.append(r)
        # column blocks
        while i < n and lines[i].strip().startswith("----"):
            i += 1
            mc = COL.match(lines[i])
            assert mc, f"{path}:{i + 1}: expected 'Col:'"
            colPtr.append(int(mc.group(1)))
            i += 2                                   # skip the "R: ..." line of the column node
            blk = []
            for _ in range(dof):
                blk.extend(_f(t) for t in lines[i].split())
                i += 1
            Val.append(blk)
        rowPtr.append(len(colPtr) + 1)
    return dict(x=np.array(xs), ltg=np.array(ltg, dtype=np.int64), R=np.array(R),
                rowPtr=np.array(rowPtr, dtype=np.int32), colPtr=np.array(colPtr, dtype=np.int32),
                Val=np.array(Val), dof=dof)


# ------------------------------------------------------------------------------------------------
# test-only: the Fortran side restated, so that the reader can be exercised without a Fortran compiler
def ndtstr(v, l=8):
    """NDTSTR(dVal, l), S/UTIL.f:539-656 (digits are TRUNCATED, not rounded)"""
    if v != v:
        return ("NaN".rjust(l)) if l >= 3 else "NaN"[:l]
    absn = abs(v)
    if math.isinf(absn):
        return "Infinity".rjust(l) if l >= 8 else "Infinity"[:l]
    if absn < sys.float_info.min:
        s = list("0.0" + "0" * (l - 3)) if l >= 3 else list("0" * l)
        return "".join(s)
    ex = int(math.floor(math.log10(absn)))
    abex = abs(ex)
    exex = int(math.floor(math.log10(float(abex)))) + 1 if ex != 0 else 0
    i = exex + 1 + (1 if v < 0 else 0) + (1 if ex < 0 else 0) + (1 if ex != 0 else 0)
    if i > l:
        return "*" * l
    s = [" "] * l
    pos = l - 1
    if ex != 0:
        for _ in range(exex):
            s[pos] = "0123456789"[abex % 10]
            abex //= 10
            pos -= 1
        if ex < 0:
            s[pos] = "-"
            pos -= 1
        s[pos] = "E"
        pos -= 1
    if l - i >= 1:
        absn = absn * (10.0 ** float(int(-ex / 2)))
        absn = absn * (10.0 ** float(l - i - 1 - ex + int(ex / 2)))
        for _ in range(l - i - 1):
            s[pos] = "0123456789"[int(math.floor(absn % 10.0))]
            absn = absn / 10.0
            pos -= 1
        s[pos] = "."
        pos -= 1
        s[pos] = "0123456789"[int(math.floor(absn % 10.0))]
    else:
        absn = absn * (10.0 ** float(l - i - ex))
        s[pos] = "0123456789"[int(math.floor(absn % 10.0))]
    if v < 0:
        s[0] = "-"
    return "".join(s)


def write_dump(path, x, ltg, R, rowPtr, colPtr, Val):
    """what PDEBUGVALR writes for one rank (x (n,3), R (n,dof), Val (nnz,dof*dof), 1-based CSR)"""
    n, dof = R.shape
    with open(path, "w") as fh:
        for a in range(n):
            fh.write("=" * 32 + "\n")
            fh.write(f"Row: {a + 1} grow: {int(ltg[a])} x: " + "".join(" " + ndtstr(v) for v in x[a]) + "\n")
            fh.write("    R:")
            for i in range(dof):
                fh.write((" " if i == 0 else " " * 7) + f"{R[a, i]:25.18E}" + "\n")
            for p in range(rowPtr[a] - 1, rowPtr[a + 1] - 1):
                b = colPtr[p] - 1
                fh.write("    " + "-" * 36 + "\n")
                fh.write(f"    Col: {b + 1} gcol: {int(ltg[b])}\n")
                fh.write("    R: " + "".join(" " + ndtstr(v) for v in R[b]) + "\n")
                for k in range(dof):
                    fh.write("    " + "".join(" " + ndtstr(v) for v in Val[p, k * dof:(k + 1) * dof]) + "\n")


def printed_tolerance(v, l=8):
    """|v - value printed by NDTSTR(v, l)| is below this (truncation of the last printed digit)"""
    if v == 0 or v != v:
        return 0.0
    s = ndtstr(v, l)
    if "*" in s:
        return float("inf")
    mant = s.split("E")[0].lstrip("-")
    digits = len(mant.replace(".", ""))
    ex = int(math.floor(math.log10(abs(v))))
    return 10.0 ** (ex - digits + 1) * 1.0000001


if __name__ == "__main__":
    d = read_dump(sys.argv[1])
    print(f"{sys.argv[1]}: nNo={d['R'].shape[0]} dof={d['dof']} nnz={d['colPtr'].size} "
          f"max|R|={np.abs(d['R']).max():.6e} max|Val|={np.nanmax(np.abs(d['Val'])):.3e}")
