#!/usr/bin/env python
"""Parses the iteration table svFSI prints (S/OUTPUT.f:66-120):

     Eq  N-i     T       dB  Ri/R1   Ri/R0    R/Ri     lsIt   dB  %t
     NS 1-1  1.230e+00  [0 1.000e+00 1.000e+00 6.6e-04]  [34 -63 71]

`N-i` = time step - Newton iteration (an `s` suffix marks the last iteration of a step), `T` =
cumulative wall time in seconds, `lsIt` = RI%itr (the SpMV count the parity tests compare), `%t` = share
of the iteration spent in FSILS_SOLVE.  Prints one JSON line with Newton-iterations/s (first time
step dropped as warm-up) and the lsIt list.   Usage: parse_svfsi_log.py svfsi_stdout.txt [ncores]"""
import json
import re
import sys

# the brackets turn into '!' when the residual grew (i > 20) / the linear solver did not converge
ROW = re.compile(r"^\s*(\w+)\s+(\d+)-(\d+)(s?)\s+([0-9.eE+-]+)\s+[\[!](.*?)[\]!]\s+[\[!](.*?)[\]!]")


def parse(text):
    rows = []
    for ln in text.splitlines():
        m = ROW.match(ln)
        if not m:
            continue
        ls = m.group(7).split()
        rows.append(dict(eq=m.group(1), step=int(m.group(2)), it=int(m.group(3)), T=float(m.group(5)),
                         lsIt=int(ls[0]) if ls else None,
                         pct_solve=float(ls[2]) if len(ls) > 2 else None))
    return rows


def summarise(rows, cores=None):
    if not rows:
        return dict(error="no iteration rows found")
    first = min(r["step"] for r in rows)
    timed = [r for r in rows if r["step"] > first] or rows
    prev = [r for r in rows if r["step"] == first]
    t0 = prev[-1]["T"] if (prev and timed is not rows) else 0.0
    dt = timed[-1]["T"] - t0
    n = len(timed)
    return dict(metric="fluid_newton_iters_per_sec", value=n / dt if dt > 0 else None,
                unit="Newton-iter/s", iterations=n, seconds=dt, cores=cores,
                lsIt=[r["lsIt"] for r in rows],
                pct_solve=[r["pct_solve"] for r in rows], kind="reference (mpiexec svFSI)")


if __name__ == "__main__":
    txt = open(sys.argv[1]).read()
    print(json.dumps(summarise(parse(txt), int(sys.argv[2]) if len(sys.argv) > 2 else None)))
