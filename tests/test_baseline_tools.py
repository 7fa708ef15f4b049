"""baseline/: the off-box recipe that runs the REAL svFSI on the synthetic pipe.  CPU
checks: the generated case is readable by our own VTK reader and names every face the deck uses; the
stdout parser understands the table of S/OUTPUT.f:66-120."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import parse_svfsi_log as psl  # noqa: E402
from svfsi_b200 import mesh, vtkio  # noqa: E402


def test_reference_case_round_trip(tmp_path):
    out = str(tmp_path / "case")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "baseline", "make_reference_case.py"),
                           "--dims", "4", "4", "6", "--out", out], stdout=subprocess.DEVNULL)
    deck = open(os.path.join(out, "svFSI.inp")).read()
    x, IEN, faces = vtkio.read_mesh_complete(os.path.join(out, "mesh"))
    m = mesh.make_cylinder(4, 4, 6)
    assert np.array_equal(IEN, m.IEN) and np.allclose(x, m.x, rtol=0, atol=0)
    for name in ("inlet", "outlet", "wall"):
        assert name in faces and f"mesh/mesh-surfaces/{name}.vtp" in deck and f"Add BC: {name}" in deck
    assert deck.count("{") == deck.count("}")
    assert "LS type: GMRES" in deck and "Krylov space dimension: 50" in deck


def test_parse_svfsi_iteration_table():
    log = """
 ---------------------------------------------------------------------
 Eq     N-i     T       dB  Ri/R1   Ri/R0    R/Ri     lsIt   dB  %t
 ---------------------------------------------------------------------
 NS 1-1  9.100e+00  [0 1.000e+00 1.000e+00 8.1e-04]  [33 -62 70]
 NS 1-2  1.820e+01  [-20 9.1e-02 9.1e-02 7.7e-04]  [34 -62 71]
 NS 1-3s 2.730e+01  [-41 8.0e-03 8.0e-03 9.0e-04]  [34 -61 71]
 NS 2-1  3.630e+01  [0 1.000e+00 6.1e-01 8.5e-04]  [35 -61 70]
 NS 2-2  4.530e+01  !25 1.9e+01 1.1e+01 1.0e-03!  !500 -3 90!
 NS 2-3s 5.430e+01  [-40 1.0e-02 6.0e-03 9.0e-04]  [34 -61 71]
"""
    rows = psl.parse(log)
    assert [(r["step"], r["it"]) for r in rows] == [(1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3)]
    assert [r["lsIt"] for r in rows] == [33, 34, 34, 35, 500, 34]
    s = psl.summarise(rows, cores=8)
    assert s["iterations"] == 3 and abs(s["seconds"] - 27.0) < 1e-9
    assert abs(s["value"] - 3 / 27.0) < 1e-12 and s["cores"] == 8


def test_pick_asm_tune_prefers_the_fastest_tested_combination(tmp_path):
    """tools/pick_asm_tune.py: fastest records kernel | fastest gather kernel among the variants whose
    parity test passed; the 8-lane kernels pay for the separate residual gather"""
    import json
    v = tmp_path / "v.json"; ok = tmp_path / "ok.txt"
    v.write_text(json.dumps({"nEl": 1, "nnz": 1, "record_tune0_ms": 3.2, "record_tune128_ms": 2.6,
                             "gather_val_tune8_ms": 6.1, "gather_val_tune40_ms": 6.3,
                             "gather_val_tune790528_ms": 4.0, "gather_r_tune0_ms": 0.33}))
    tool = os.path.join(ROOT, "tools", "pick_asm_tune.py")

    def pick():
        return int(subprocess.check_output([sys.executable, tool, str(v), str(ok)], text=True))
    ok.write_text("0\n8\n40\n136\n168\n790528\n790656\n")
    assert pick() == 128 + 790528
    ok.write_text("0\n8\n40\n136\n168\n")            # quad variant failed its test: not eligible
    assert pick() == 128 + 40                         # 6.3 (B + C in one launch) beats 6.1 + 0.33
    ok.write_text("")
    assert pick() == 8                                # nothing verified: the plain 8-lane kernel


def test_launch_list_summary(tmp_path):
    csvf = tmp_path / "l.csv"
    csvf.write_text('==PROF== Connected\n"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream",'
                    '"Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"\n'
                    '"0","1","p","h","svfsi::a_kernel(int)","1","13","(256, 1, 1)","(1, 1, 1)","0","10.0","s",'
                    '"gpu__time_duration.sum","ns","1500"\n'
                    '"1","1","p","h","svfsi::a_kernel(int)","1","13","(256, 1, 1)","(1, 1, 1)","0","10.0","s",'
                    '"gpu__time_duration.sum","us","2.5"\n'
                    '"2","1","p","h","svfsi::b_kernel(int)","1","13","(256, 1, 1)","(1, 1, 1)","0","10.0","s",'
                    '"gpu__time_duration.sum","ns","1000"\n')
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), str(csvf)],
                                  text=True)
    assert "| `svfsi::a_kernel` | 2 | 4.0 | 2.0 | 80.0% |" in out
    assert "| `svfsi::b_kernel` | 1 | 1.0 | 1.0 | 20.0% |" in out
