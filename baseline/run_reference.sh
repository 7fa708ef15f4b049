#!/bin/bash
# Off-box recipe: build the REAL svFSI and time it on the same synthetic pipe.
# Needs what Docker/Dockerfile:8-11 of the reference installs (gfortran, OpenMPI, BLAS/LAPACK, cmake);
# none of it exists in the B200 image, so bench.py's CPU arm is the C port of the same routines
# (oracle/, "kind": "port").  Run this on any workstation with that toolchain:
#
#   baseline/run_reference.sh /path/to/svFSI-source /tmp/pipe10M 64 64 408 [nprocs]
#
# Output: <case>/svfsi_stdout.txt and one JSON line (Newton-iterations/s on `nprocs` cores, lsIt list)
# comparable with bench.py's "cpu_baseline" / the GMRES counts of tests/test_gpu_parity.py.
set -euo pipefail
SRC=${1:?svFSI source tree}
CASE=${2:?case directory}
NX=${3:-64}; NY=${4:-64}; NZ=${5:-408}
NP=${6:-$(nproc)}
HERE=$(cd "$(dirname "$0")" && pwd)
BUILD=${SVFSI_BUILD:-$CASE/build}

# SVFSI_DUMP=1: build from a copy of the source with `CALL PDEBUGVALR()` inserted after COMMU(R) (S/MAIN.f:161), so
# that every Newton iteration writes the assembled R / Val of each rank (S/DEBUG.f:122-176, file
# Val_R_<cTS>_<itr>_<rank>).  Run with NP=1 and copy Val_R_1_1_0 to tests/golden/ref_Val_R_${NX}x${NY}x${NZ}_1_1_0:
# tests/test_reference_dumps.py then pins the oracle and the CUDA path to the real svFSI.
if [ "${SVFSI_DUMP:-0}" = "1" ]; then
  cp -r "$SRC" "$CASE/src_dump"
  SRC="$CASE/src_dump"
  # S/MAIN.f:161 reads "IF (.NOT.eq(cEq)%assmTLS) CALL COMMU(R)"
  sed -i '/CALL COMMU(R)/a\            CALL PDEBUGVALR()' "$SRC/Code/Source/svFSI/MAIN.f"
  BUILD="$CASE/build_dump"
fi
if [ ! -x "$BUILD/svFSI-build/bin/svFSI" ]; then
  mkdir -p "$BUILD"
  (cd "$BUILD" && cmake "$SRC" -DCMAKE_BUILD_TYPE=Release && make -j"$NP")
fi
python "$HERE/make_reference_case.py" --dims "$NX" "$NY" "$NZ" --out "$CASE"
(cd "$CASE" && mpiexec -np "$NP" "$BUILD/svFSI-build/bin/svFSI" svFSI.inp | tee svfsi_stdout.txt)
python "$HERE/parse_svfsi_log.py" "$CASE/svfsi_stdout.txt" "$NP"
