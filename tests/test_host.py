"""CPU tests of the host side of the product: the C-ABI library loads and exports every symbol the
header declares, fails loudly without a GPU, and the host-only part of FSILS_LHS_CREATE (node
reordering + halo schedule) matches the oracle -- also across two real processes (gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import common as cm
from oracle import oracle as ora
from svfsi_b200 import api, mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "svfsi_b200.h")).read()
    declared = set(re.findall(r"\bint32_t\s+((?:gpu|svfsi)_\w+_)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(api.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)


def test_no_oracle_or_cpu_fallback_in_the_product():
    """the product must not import/link the oracle: grep the package and the shared object"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "svfsi_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libsvfsi_oracle" not in txt, f
    out = subprocess.run(["ldd", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_init_fails_loudly_without_gpu():
    with pytest.raises(api.SvfsiError) as ei:
        api.init(device=0, rank=0, nranks=1)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    with pytest.raises(api.SvfsiError):
        api.FSILS_LHS_CREATE(1, 1, 1, [1], [1, 2], [1], 0)


def test_ls_create_defaults_match_reference():
    ls = api.FSILS_LS_CREATE(api.LS_TYPE_NS)
    lo = ora.ls_create(ora.LS_TYPE_NS)
    for sub in ("RI", "GM", "CG"):
        a, b = getattr(ls, sub), getattr(lo, sub)
        assert (a.relTol, a.absTol, a.mItr, a.sD) == (b.relTol, b.absTol, b.mItr, b.sD)
    with pytest.raises(api.SvfsiError):
        api.FSILS_LS_CREATE(123)


@pytest.mark.parametrize("nparts", [1, 2, 3, 5])
def test_lhs_plan_matches_oracle(nparts):
    m, probs, _ = mesh.build_problem(4, 4, 10, nparts=nparts, L=3.0)
    w = cm.oracle_world(probs, m.nNo, with_faces=False)
    ltgs = [p.rm.ltg for p in probs]
    for r in range(nparts):
        plan = api.lhs_plan(r, nparts, m.nNo, ltgs)
        oi = w.info(r)
        assert (plan["mynNo"], plan["nReq"]) == (oi["mynNo"], oi["nReq"])
        if nparts > 1:
            assert plan["shnNo"] == oi["shnNo"]
        assert np.array_equal(plan["map"], w.map(r))
        for (iPa, pa), (iPb, pb) in zip(plan["cS"], w.cs(r)):
            assert iPa == iPb and np.array_equal(pa, pb)


def _quadrant_problems(nx=4, ny=4, nz=8, L=3.0):
    """an element partition whose cut nodes are shared by up to four ranks (the slab partitions of
    the other tests never share a node between more than two): sign of the element centroid's x and
    y picks the rank"""
    m = mesh.make_cylinder(nx, ny, nz, L=L)
    c = m.x[m.IEN.astype(np.int64) - 1].mean(axis=1)
    part = ((c[:, 0] > 0).astype(np.int32) + 2 * (c[:, 1] > 0).astype(np.int32))
    rms = mesh.split_mesh(m, part, 4)
    Ag, Yg = mesh.poiseuille_state(m)
    probs = []
    for rm in rms:
        rp, cp = mesh.csr_pattern(rm.nNo, rm.IEN)
        probs.append(mesh.RankProblem(rm, rp, cp, mesh.scatter_nodal(rm, Ag), mesh.scatter_nodal(rm, Yg), {}))
    return m, probs


def test_lhs_plan_matches_oracle_when_nodes_are_shared_by_four_ranks():
    m, probs = _quadrant_problems()
    ltgs = [p.rm.ltg for p in probs]
    mult = np.zeros(m.nNo + 1, dtype=int)
    for l in ltgs:
        mult[l] += 1
    assert mult.max() == 4                      # the axis nodes
    w = cm.oracle_world(probs, m.nNo, with_faces=False)
    for r in range(4):
        plan = api.lhs_plan(r, 4, m.nNo, ltgs)
        oi = w.info(r)
        assert (plan["mynNo"], plan["nReq"], plan["shnNo"]) == (oi["mynNo"], oi["nReq"], oi["shnNo"])
        assert np.array_equal(plan["map"], w.map(r))
        assert len(plan["cS"]) == len(w.cs(r)) == 3
        for (iPa, pa), (iPb, pb) in zip(plan["cS"], w.cs(r)):
            assert iPa == iPb and np.array_equal(pa, pb)
    # every global node is owned by exactly one rank (owned = the first mynNo of the reordered numbering)
    owned = np.zeros(m.nNo + 1, dtype=int)
    for r in range(4):
        plan = api.lhs_plan(r, 4, m.nNo, ltgs)
        inv = np.argsort(plan["map"])
        owned[ltgs[r][inv[: plan["mynNo"]]]] += 1
    assert (owned[1:] == 1).all()


def test_oracle_four_rank_solve_matches_one_rank_on_the_quadrant_partition():
    """COMMUV with three neighbours per rank and nodes summed over four ranks: the 4-rank oracle's
    assembled + halo-summed residual and its GMRES step equal the 1-rank oracle's."""
    m, probs = _quadrant_problems()
    Rs, Vs = cm.oracle_assemble(probs)
    w = cm.oracle_world(probs, m.nNo, nFaces=0, with_faces=False)
    Rc = cm.oracle_commu(w, probs, Rs)
    m1, probs1, _ = mesh.build_problem(4, 4, 8, nparts=1, L=3.0)
    R1, V1 = cm.oracle_assemble(probs1)
    G = np.zeros((m.nNo, 4))
    for p, r in zip(probs, Rc):
        G[p.rm.ltg - 1] = r
    assert cm.rel_err(G, R1[0][np.argsort(probs1[0].rm.ltg)]) <= 1e-13
    ls4 = ora.ls_create(ora.LS_TYPE_GMRES, relTol=1e-4, absTol=1e-14, maxItr=10, dimKry=80)
    X = [r.copy() for r in Rc]
    w.solve(ls4, 4, X, [v.copy() for v in Vs])
    w1 = cm.oracle_world(probs1, m.nNo, nFaces=0, with_faces=False)
    ls1 = ora.ls_create(ora.LS_TYPE_GMRES, relTol=1e-4, absTol=1e-14, maxItr=10, dimKry=80)
    X1 = [R1[0].copy()]
    w1.solve(ls1, 4, X1, [V1[0].copy()])
    G4 = np.zeros((m.nNo, 4))
    for p, x in zip(probs, X):
        G4[p.rm.ltg - 1] = x
    G1 = np.zeros((m.nNo, 4)); G1[probs1[0].rm.ltg - 1] = X1[0]
    assert abs(ls4.RI.itr - ls1.RI.itr) <= 1
    assert np.linalg.norm(G4 - G1) / np.linalg.norm(G1) <= 1e-8


def test_csr_pattern_matches_lhsa():
    m = mesh.make_cylinder(4, 4, 5)
    rp, cp = mesh.csr_pattern(m.nNo, m.IEN)
    rp2, cp2 = ora.lhsa(m.nNo, m.IEN)
    assert np.array_equal(rp, rp2) and np.array_equal(cp, cp2)


def test_rank_local_generator_matches_global_split():
    m, probs, _ = mesh.build_problem(4, 4, 9, 3, L=3.0)
    for r in range(3):
        g, p = mesh.build_rank_problem(4, 4, 9, r, 3, L=3.0)
        q = probs[r]
        assert g == m.nNo and np.array_equal(p.rm.ltg, q.rm.ltg) and np.array_equal(p.rm.IEN, q.rm.IEN)
        assert np.array_equal(p.rowPtr, q.rowPtr) and np.array_equal(p.colPtr, q.colPtr)


_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
import common as cm
from svfsi_b200 import api, mesh
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
m, probs, _ = mesh.build_problem(4, 4, 8, nparts=2, L=3.0)
p = probs[rank]
# the MPI_ALLGATHERV of FSILS_LHS_CREATE (L/LHS.f:113-126) done with gloo
cnt = [torch.zeros(1, dtype=torch.int32) for _ in range(2)]
dist.all_gather(cnt, torch.tensor([p.rm.nNo], dtype=torch.int32))
maxn = int(max(c.item() for c in cnt))
mine = torch.zeros(maxn, dtype=torch.int32); mine[:p.rm.nNo] = torch.from_numpy(p.rm.ltg.copy())
allv = [torch.zeros(maxn, dtype=torch.int32) for _ in range(2)]
dist.all_gather(allv, mine)
ltgs = [v.numpy()[: int(c.item())] for v, c in zip(allv, cnt)]
plan = api.lhs_plan(rank, 2, m.nNo, ltgs)
w = cm.oracle_world(probs, m.nNo, with_faces=False)
assert np.array_equal(plan["map"], w.map(rank))
assert plan["mynNo"] == w.info(rank)["mynNo"]
# pair-consistent halo lists: both sides must list the same GLOBAL ids in the same order
inv = np.argsort(plan["map"])            # reordered id -> local id
gl = p.rm.ltg[inv[plan["cS"][0][1] - 1]]
other = [torch.zeros(gl.size, dtype=torch.int32) for _ in range(2)]
dist.all_gather(other, torch.from_numpy(gl.astype(np.int32)))
assert torch.equal(other[0], other[1])
# owned-node dot over both ranks == global dot (FSILS_DOTV semantics, L/DOT.f:51-55)
x = np.random.default_rng(5).standard_normal(m.nNo)
loc = x[p.rm.ltg[inv[: plan["mynNo"]]] - 1]
s = torch.tensor([float((loc * loc).sum())], dtype=torch.float64)
dist.all_reduce(s)
assert abs(s.item() - float((x * x).sum())) < 1e-10
dist.destroy_process_group()
print("WORKER_OK", rank)
'''


def test_two_process_gloo_lhs_plan(tmp_path):
    """world_size-2 gloo: the N>1 host path (all-gather of node lists -> plan) on CPU"""
    port = 29650 + (os.getpid() % 200)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER_OK {r}" in o, o[-3000:]
