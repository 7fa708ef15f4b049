"""the on-disk formats either side of the hot path (S/VTKXML.f:39-150 READVTU/READVTP,
S/LOADMSH.f:39-91 mesh-complete layout, S/OUTPUT.f:132-232 + S/INITIALIZE.f:146-170 restart record).
CPU only: the reference ships no fixture files, so the pins are (i) round trips through every
DataArray encoding of the VTK XML format, (ii) hand-written ascii files with known content (what
svFSI-Tests' small meshes look like), (iii) the byte layout of the direct-access record.
"""
import base64
import os
import struct

import numpy as np
import pytest

from svfsi_b200 import mesh as M
from svfsi_b200 import vtkio as V


ENCODINGS = [("appended", False), ("appended", True), ("appended-base64", False), ("appended-base64", True),
             ("binary", False), ("binary", True), ("ascii", False)]


@pytest.fixture(scope="module")
def small_mesh():
    return M.make_cylinder(4, 4, 3)


@pytest.mark.parametrize("encoding,compress", ENCODINGS)
@pytest.mark.parametrize("header_type", ["UInt32", "UInt64"])
def test_vtu_round_trip_every_encoding(tmp_path, small_mesh, encoding, compress, header_type):
    m = small_mesh
    rng = np.random.default_rng(M.SEED)
    vel = rng.standard_normal((m.nNo, 3))
    prs = rng.standard_normal(m.nNo)
    dom = np.arange(m.nEl, dtype=np.int32) % 3
    p = str(tmp_path / "a.vtu")
    V.write_vtu(p, m.x, m.IEN, point_data={"Velocity": vel, "Pressure": prs}, cell_data={"Domain_ID": dom},
                encoding=encoding, compress=compress, header_type=header_type)
    x, IEN, piece = V.read_vtu(p)
    assert IEN.dtype == np.int32 and IEN.min() == 1            # READVTU: gIEN = gIEN + 1
    np.testing.assert_array_equal(IEN, m.IEN)
    np.testing.assert_array_equal(x, m.x)                      # bit-exact, ascii uses repr()
    np.testing.assert_array_equal(piece.point_data["Velocity"], vel)
    np.testing.assert_array_equal(piece.point_data["Pressure"], prs)
    np.testing.assert_array_equal(piece.cell_data["Domain_ID"], dom)
    assert (piece.types == V.VTK_TETRA).all() and piece.nodes_per_cell() == 4


@pytest.mark.parametrize("encoding,compress", ENCODINGS)
def test_mesh_complete_round_trip(tmp_path, small_mesh, encoding, compress):
    m = small_mesh
    d = str(tmp_path / "mesh")
    V.write_mesh_complete(m, d, encoding=encoding, compress=compress)
    assert sorted(os.listdir(os.path.join(d, "mesh-surfaces"))) == ["inlet.vtp", "outlet.vtp", "wall.vtp"]
    x, IEN, faces = V.read_mesh_complete(d)
    np.testing.assert_array_equal(x, m.x)
    np.testing.assert_array_equal(IEN, m.IEN)
    for name, fa in m.faces.items():
        np.testing.assert_array_equal(faces[name]["gN"], fa.gN)
        np.testing.assert_array_equal(faces[name]["IEN"], fa.tri)       # READVTP maps through GlobalNodeID
        np.testing.assert_array_equal(faces[name]["gE"], fa.parent + 1)
        # every face element is a face of its parent tet (what svFSI's SETFACEEBC relies on)
        par = m.IEN[faces[name]["gE"] - 1]
        assert all(set(t) <= set(q) for t, q in zip(faces[name]["IEN"], par))


def test_hand_written_ascii_vtu(tmp_path):
    """A two-tet file typed by hand the way VTK's ascii writer lays it out (Float32 points, Int64 ids)."""
    p = tmp_path / "two.vtu"
    p.write_text("""<?xml version="1.0"?>
<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">
  <UnstructuredGrid>
    <Piece NumberOfPoints="5" NumberOfCells="2">
      <PointData Scalars="GlobalNodeID">
        <DataArray type="Int32" Name="GlobalNodeID" format="ascii">
          1 2 3 4 5
        </DataArray>
      </PointData>
      <Points>
        <DataArray type="Float32" NumberOfComponents="3" format="ascii">
          0 0 0  1 0 0  0 1 0
          0 0 1  1 1 1
        </DataArray>
      </Points>
      <Cells>
        <DataArray type="Int64" Name="connectivity" format="ascii">0 1 2 3 1 2 3 4</DataArray>
        <DataArray type="Int64" Name="offsets" format="ascii">4 8</DataArray>
        <DataArray type="UInt8" Name="types" format="ascii">10 10</DataArray>
      </Cells>
    </Piece>
  </UnstructuredGrid>
</VTKFile>
""")
    x, IEN, piece = V.read_vtu(str(p))
    assert x.dtype == np.float64 and x.shape == (5, 3)
    np.testing.assert_array_equal(IEN, [[1, 2, 3, 4], [2, 3, 4, 5]])
    np.testing.assert_array_equal(x[4], [1, 1, 1])
    np.testing.assert_array_equal(piece.point_data["GlobalNodeID"], [1, 2, 3, 4, 5])


def test_hand_packed_binary_header_layout(tmp_path):
    """format="binary" as VTK writes it: base64( UInt32 nbytes | payload ) in ONE stream; built here with
    struct/base64 directly so that the reader is checked against the format, not against our writer."""
    pts = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0]], dtype="<f8")
    conn = np.array([0, 1, 2], dtype="<i4")

    def b64(a):
        raw = a.tobytes()
        return base64.b64encode(struct.pack("<I", len(raw)) + raw).decode()

    p = tmp_path / "tri.vtp"
    p.write_text(f"""<?xml version="1.0"?>
<VTKFile type="PolyData" version="0.1" byte_order="LittleEndian" header_type="UInt32">
<PolyData><Piece NumberOfPoints="3" NumberOfVerts="0" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="1">
<PointData><DataArray type="Int32" Name="GlobalNodeID" format="binary">{b64(np.array([7, 9, 4], dtype="<i4"))}</DataArray></PointData>
<CellData><DataArray type="Int32" Name="GlobalElementID" format="binary">{b64(np.array([12], dtype="<i4"))}</DataArray></CellData>
<Points><DataArray type="Float64" NumberOfComponents="3" format="binary">{b64(pts)}</DataArray></Points>
<Polys><DataArray type="Int32" Name="connectivity" format="binary">{b64(conn)}</DataArray>
<DataArray type="Int32" Name="offsets" format="binary">{b64(np.array([3], dtype="<i4"))}</DataArray></Polys>
</Piece></PolyData></VTKFile>
""")
    x, ien, gN, gE, _ = V.read_vtp(str(p))
    np.testing.assert_array_equal(x, pts)
    np.testing.assert_array_equal(ien, [[7, 9, 4]])      # S/VTKXML.f:124-133: IEN -> gN(IEN+1)
    np.testing.assert_array_equal(gN, [7, 9, 4])
    np.testing.assert_array_equal(gE, [12])


def test_vtp_without_global_ids_keeps_local_connectivity(tmp_path):
    """S/VTKXML.f:120-123,136-139: missing GlobalNodeID / GlobalElementID is a warning, not an error"""
    p = str(tmp_path / "f.vtp")
    V.write_vtp(p, np.eye(3), np.array([[1, 2, 3]]))
    x, ien, gN, gE, _ = V.read_vtp(p)
    assert gN is None and gE is None
    np.testing.assert_array_equal(ien, [[1, 2, 3]])
    d = tmp_path / "mc"
    (d / "mesh-surfaces").mkdir(parents=True)
    V.write_vtu(str(d / "mesh-complete.mesh.vtu"), np.vstack([np.zeros(3), np.eye(3)]), np.array([[1, 2, 3, 4]]))
    V.write_vtp(str(d / "mesh-surfaces" / "f.vtp"), np.eye(3), np.array([[1, 2, 3]]))
    with pytest.raises(V.VtkError, match="GlobalNodeID"):
        V.read_mesh_complete(str(d))


def test_error_behaviour(tmp_path, small_mesh):
    bad = tmp_path / "bad.vtu"
    bad.write_text("this is not xml")
    with pytest.raises(V.VtkError):
        V.read_vtu(str(bad))
    p = str(tmp_path / "a.vtp")
    V.write_vtp(p, np.eye(3), np.array([[1, 2, 3]]))
    with pytest.raises(V.VtkError, match="UnstructuredGrid"):
        V.read_vtu(p)                                          # a .vtp handed to READVTU
    q = str(tmp_path / "a.vtu")
    V.write_vtu(q, small_mesh.x, small_mesh.IEN)
    with pytest.raises(V.VtkError, match="PolyData"):
        V.read_vtp(q)
    with pytest.raises(V.VtkError, match="nodes"):
        V.write_vtu(q, small_mesh.x, small_mesh.IEN[:, :3])    # TETRA with 3 nodes
    mixed = V.VtkPiece("UnstructuredGrid", np.zeros((5, 3)), np.arange(7), np.array([4, 7]))
    with pytest.raises(V.VtkError, match="mixed"):
        mixed.nodes_per_cell()                                  # getVTK_nodesPerElem needs one cell size


def test_empty_face(tmp_path):
    p = str(tmp_path / "e.vtp")
    V.write_vtp(p, np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64), gN=np.zeros(0, np.int32),
                gE=np.zeros(0, np.int32))
    x, ien, gN, gE, piece = V.read_vtp(p)
    assert x.shape == (0, 3) and piece.n_cells == 0 and ien.size == 0


def test_restart_record_layout(tmp_path):
    """Byte layout of WRITERESTART's record (S/OUTPUT.f:204-205: stamp, cTS, time, tt-timeP(1), eq%iNorm,
    cplBC%xn, Yn, An) and recLn of S/INITIALIZE.f:154-163; two ranks in one direct-access file."""
    nEq, nX, tDof = 1, 2, 4
    tn = [5, 7]
    recLn = max(V.restart_reclen(nEq, nX, tDof, n) for n in tn)     # MPI_ALLREDUCE(MAX), INITIALIZE.f:168
    assert V.restart_reclen(nEq, nX, tDof, 7) == 4 * 8 + 8 * (2 + 1 + 2 + 2 * 4 * 7)
    assert V.restart_reclen(nEq, nX, tDof, 7, dFlag=True) == 4 * 8 + 8 * (2 + 1 + 2 + 3 * 4 * 7)
    p = str(tmp_path / "stFile_last.bin")
    rng = np.random.default_rng(M.SEED)
    state = []
    for r, n in enumerate(tn, start=1):
        Yn, An = rng.standard_normal((n, tDof)), rng.standard_normal((n, tDof))
        stamp = [2, nEq, 1, n, nX, tDof, 0]
        V.write_restart(p, r, recLn, stamp, cTS=40 + r, time=0.2 * r, timeP=1.5, iNorm=[3.25],
                        xn=[0.5, -0.5], Yn=Yn, An=An)
        state.append((stamp, Yn, An))
    assert os.path.getsize(p) == 2 * recLn
    with open(p, "rb") as fh:
        blob = fh.read()
    rec2 = blob[recLn:]
    assert struct.unpack_from("<7i", rec2, 0) == (2, 1, 1, 7, 2, 4, 0)
    assert struct.unpack_from("<i", rec2, 28)[0] == 42
    assert struct.unpack_from("<2d", rec2, 32) == (0.4, 1.5)
    assert struct.unpack_from("<d", rec2, 48)[0] == 3.25
    assert struct.unpack_from("<2d", rec2, 56) == (0.5, -0.5)
    # Fortran Yn(tDof, tnNo) column-major == our (tnNo, tDof) row-major: node 0's four dofs come first
    np.testing.assert_array_equal(np.frombuffer(rec2, "<f8", 4, 72), state[1][1][0])
    for r, n in enumerate(tn, start=1):
        got = V.read_restart(p, r, recLn, nEq, nX, tDof, n, expect_stamp=state[r - 1][0])
        assert got["cTS"] == 40 + r and got["time"] == 0.2 * r and got["timeP"] == 1.5
        np.testing.assert_array_equal(got["Yo"], state[r - 1][1])
        np.testing.assert_array_equal(got["Ao"], state[r - 1][2])
        np.testing.assert_array_equal(got["xo"], [0.5, -0.5])
        assert got["Do"] is None
    with pytest.raises(ValueError, match="Number of dof"):       # S/INITIALIZE.f:593-617
        V.read_restart(p, 1, recLn, nEq, nX, tDof, 5, expect_stamp=[2, nEq, 1, 5, nX, 3, 0])
    with pytest.raises(ValueError, match="exceeds"):
        V.write_restart(p, 1, 64, [2, 1, 1, 5, 2, 4, 0], 1, 0.0, 0.0, [0.0], [0.0, 0.0],
                        np.zeros((5, 4)), np.zeros((5, 4)))


def test_restart_with_displacement(tmp_path):
    nEq, nX, tDof, n = 2, 0, 7, 3
    recLn = V.restart_reclen(nEq, nX, tDof, n, dFlag=True)
    rng = np.random.default_rng(1)
    Yn, An, Dn = (rng.standard_normal((n, tDof)) for _ in range(3))
    p = str(tmp_path / "r.bin")
    V.write_restart(p, 1, recLn, [1, nEq, 1, n, nX, tDof, 1], 3, 0.1, 0.0, [1.0, 2.0], [], Yn, An, Dn)
    got = V.read_restart(p, 1, recLn, nEq, nX, tDof, n, dFlag=True)
    np.testing.assert_array_equal(got["Do"], Dn)
    np.testing.assert_array_equal(got["iNorm"], [1.0, 2.0])
