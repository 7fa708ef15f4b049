"""On-disk formats either side of the hot path: VTK XML unstructured grids
(`.vtu`, volume meshes and results) and polydata (`.vtp`, boundary faces) as svFSI reads and
writes them (S/LOADMSH.f:39-91, S/VTKXML.f:39-410: READVTU / READVTP / WRITEVTUS; node ids in the
files are 0-based, svFSI's arrays 1-based), the `mesh-complete` directory layout of svFSI-Tests
cases, and the direct-access restart record (S/OUTPUT.f:132-232, S/INITIALIZE.f:146-170,512-620).

Host-side NumPy code: these files are read once at start-up and written every few hundred time
steps; nothing here is on the Newton-iteration path.  The reader accepts every DataArray encoding
of the VTK XML format (ascii, inline base64 with optional zlib blocks, appended raw or base64,
UInt32 / UInt64 headers, either byte order); the writer emits appended-raw by default (what
ParaView and svFSI's own parser read fastest) and the other encodings on request.
"""
from __future__ import annotations

import base64
import os
import re
import struct
import sys
import xml.etree.ElementTree as ET
import zlib
from dataclasses import dataclass, field

import numpy as np

_VTK2NP = {"Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2", "Int32": "i4", "UInt32": "u4",
           "Int64": "i8", "UInt64": "u8", "Float32": "f4", "Float64": "f8"}
_NP2VTK = {np.dtype(v).str[1:]: k for k, v in _VTK2NP.items()}
VTK_TRIANGLE, VTK_TETRA, VTK_LINE, VTK_QUAD, VTK_HEXAHEDRON = 5, 10, 3, 9, 12
_NODES_OF = {VTK_TRIANGLE: 3, VTK_TETRA: 4, VTK_LINE: 2, VTK_QUAD: 4, VTK_HEXAHEDRON: 8}


class VtkError(ValueError):
    pass


@dataclass
class VtkPiece:
    kind: str                                  # "UnstructuredGrid" | "PolyData"
    points: np.ndarray                         # (nNo, 3) float64
    connectivity: np.ndarray                   # flat, 0-based
    offsets: np.ndarray                        # (nEl,) end offsets
    types: np.ndarray | None = None            # (nEl,) VTK cell types (UnstructuredGrid only)
    point_data: dict = field(default_factory=dict)
    cell_data: dict = field(default_factory=dict)

    @property
    def n_cells(self):
        return int(self.offsets.size)

    def nodes_per_cell(self):
        """getVTK_nodesPerElem: all cells must have the same size (svFSI meshes are homogeneous)"""
        sizes = np.diff(np.concatenate([[0], self.offsets]))
        if sizes.size == 0:
            return 0
        if not (sizes == sizes[0]).all():
            raise VtkError("cells of mixed size: svFSI meshes have one element type")
        return int(sizes[0])

    def ien(self):
        """(nEl, eNoN) 0-based connectivity"""
        n = self.nodes_per_cell()
        return self.connectivity.reshape(-1, n) if n else self.connectivity.reshape(0, 0)


# ------------------------------------------------------------------------------------ decoding
def _b64len(nbytes):
    return ((nbytes + 2) // 3) * 4


def _decode_blocks(raw_header_and_data, hdt, compressed, from_base64):
    """raw_header_and_data: str (base64 text) or bytes (appended raw).  Returns the payload bytes."""
    hsz = hdt.itemsize
    if not compressed:
        if from_base64:
            txt = raw_header_and_data
            n = int(np.frombuffer(base64.b64decode(txt[:_b64len(hsz)])[:hsz], hdt)[0])
            # VTK encodes header and data in ONE base64 stream; some writers encode them apart
            one = base64.b64decode(txt[:_b64len(hsz + n)])
            if len(one) >= hsz + n:
                return one[hsz:hsz + n]
            return base64.b64decode(txt[_b64len(hsz):])[:n]
        buf = raw_header_and_data
        n = int(np.frombuffer(buf[:hsz], hdt)[0])
        return bytes(buf[hsz:hsz + n])
    # zlib: header = [nblocks, blocksize, lastblocksize, csize_0 ... csize_{nblocks-1}]
    if from_base64:
        txt = raw_header_and_data
        nb = int(np.frombuffer(base64.b64decode(txt[:_b64len(3 * hsz)])[:hsz], hdt)[0])
        hbytes = hsz * (3 + nb)
        hchars = _b64len(hbytes)
        head = np.frombuffer(base64.b64decode(txt[:hchars])[:hbytes], hdt)
        body = base64.b64decode(txt[hchars:])
    else:
        buf = raw_header_and_data
        nb = int(np.frombuffer(buf[:hsz], hdt)[0])
        hbytes = hsz * (3 + nb)
        head = np.frombuffer(buf[:hbytes], hdt)
        body = buf[hbytes:]
    out, at = [], 0
    for cs in head[3:3 + nb]:
        cs = int(cs)
        out.append(zlib.decompress(bytes(body[at:at + cs])))
        at += cs
    return b"".join(out)


def _appended_span(buf, off, hdt, compressed):
    """number of bytes (header + data) of the appended-raw array starting at off"""
    hsz = hdt.itemsize
    if not compressed:
        n = int(np.frombuffer(buf[off:off + hsz], hdt)[0])
        return hsz + n
    nb = int(np.frombuffer(buf[off:off + hsz], hdt)[0])
    head = np.frombuffer(buf[off:off + hsz * (3 + nb)], hdt)
    return hsz * (3 + nb) + int(head[3:].sum())


class _Reader:
    def __init__(self, path):
        with open(path, "rb") as fh:
            blob = fh.read()
        # the AppendedData section may hold raw bytes that are not XML: cut it out before parsing
        self.app_raw = None
        self.app_enc = None
        m = re.search(rb"<AppendedData[^>]*>", blob)
        if m:
            enc = re.search(rb'encoding\s*=\s*"([^"]+)"', m.group(0))
            self.app_enc = enc.group(1).decode() if enc else "base64"
            start = blob.index(b"_", m.end()) + 1
            end = blob.rindex(b"</AppendedData>")
            self.app_raw = blob[start:end]
            blob = blob[:m.end()] + b"</AppendedData>" + blob[end + len(b"</AppendedData>"):]
        try:
            self.root = ET.fromstring(blob)
        except ET.ParseError as e:
            raise VtkError(f"{path}: not a VTK XML file ({e})") from None
        if self.root.tag != "VTKFile":
            raise VtkError(f"{path}: root element is <{self.root.tag}>, expected <VTKFile>")
        self.kind = self.root.get("type")
        bo = "<" if self.root.get("byte_order", "LittleEndian") == "LittleEndian" else ">"
        self.bo = bo
        self.hdt = np.dtype(bo + _VTK2NP[self.root.get("header_type", "UInt32")])
        self.compressed = self.root.get("compressor") is not None
        if self.compressed and "ZLib" not in self.root.get("compressor"):
            raise VtkError(f"{path}: compressor {self.root.get('compressor')} is not supported (zlib only)")
        if self.app_enc == "base64" and self.app_raw is not None:
            self.app_raw = b"".join(self.app_raw.split())

    def array(self, da):
        typ = da.get("type")
        if typ not in _VTK2NP:
            raise VtkError(f"DataArray type {typ} is not supported")
        dt = np.dtype(self.bo + _VTK2NP[typ])
        fmt = da.get("format", "ascii")
        ncomp = int(da.get("NumberOfComponents", "1"))
        if fmt == "ascii":
            txt = da.text or ""
            a = np.array(txt.split(), dtype=np.float64 if dt.kind == "f" else np.int64).astype(dt.newbyteorder("="))
        elif fmt == "binary":
            payload = _decode_blocks("".join((da.text or "").split()), self.hdt, self.compressed, True)
            a = np.frombuffer(payload, dt).astype(dt.newbyteorder("="))
        elif fmt == "appended":
            if self.app_raw is None:
                raise VtkError("format=appended but the file has no AppendedData section")
            off = int(da.get("offset", "0"))
            if self.app_enc == "raw":
                span = _appended_span(self.app_raw, off, self.hdt, self.compressed)
                payload = _decode_blocks(self.app_raw[off:off + span], self.hdt, self.compressed, False)
            else:
                payload = _decode_blocks(self.app_raw[off:].decode("ascii"), self.hdt, self.compressed, True)
            a = np.frombuffer(payload, dt).astype(dt.newbyteorder("="))
        else:
            raise VtkError(f"DataArray format {fmt} is not supported")
        return a.reshape(-1, ncomp) if ncomp > 1 else a

    def data_section(self, piece, name):
        sec = piece.find(name)
        out = {}
        if sec is not None:
            for da in sec.findall("DataArray"):
                out[da.get("Name")] = self.array(da)
        return out

    def named(self, sec, name):
        for da in sec.findall("DataArray"):
            if da.get("Name") == name:
                return self.array(da)
        raise VtkError(f"<{sec.tag}> has no DataArray named {name}")


def read_vtk_xml(path) -> VtkPiece:
    """loadVTK + the getVTK_* accessors (S/vtkXMLParser.f90 is svFSI's own parser): one piece."""
    r = _Reader(path)
    if r.kind not in ("UnstructuredGrid", "PolyData"):
        raise VtkError(f"{path}: VTKFile type {r.kind} is not supported")
    grid = r.root.find(r.kind)
    pieces = grid.findall("Piece")
    if len(pieces) != 1:
        raise VtkError(f"{path}: {len(pieces)} pieces; svFSI meshes are written as one piece")
    pc = pieces[0]
    pts = r.array(pc.find("Points").find("DataArray")).astype(np.float64)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 3)
    if r.kind == "UnstructuredGrid":
        cells = pc.find("Cells")
        conn = r.named(cells, "connectivity").astype(np.int64)
        offs = r.named(cells, "offsets").astype(np.int64)
        types = r.named(cells, "types").astype(np.int64)
    else:
        conn = offs = None
        for tag in ("Polys", "Lines", "Strips", "Verts"):
            sec = pc.find(tag)
            if sec is not None and int(pc.get("NumberOf" + tag, "0")) > 0:
                conn = r.named(sec, "connectivity").astype(np.int64)
                offs = r.named(sec, "offsets").astype(np.int64)
                break
        if conn is None:
            conn, offs = np.zeros(0, np.int64), np.zeros(0, np.int64)
        types = None
    npts = int(pc.get("NumberOfPoints", pts.shape[0]))
    if npts != pts.shape[0]:
        raise VtkError(f"{path}: NumberOfPoints={npts} but {pts.shape[0]} coordinates")
    return VtkPiece(r.kind, pts, conn, offs, types, r.data_section(pc, "PointData"),
                    r.data_section(pc, "CellData"))


def read_vtu(path):
    """READVTU (S/VTKXML.f:39-81): x (gnNo, 3), gIEN (gnEl, eNoN) 1-based."""
    p = read_vtk_xml(path)
    if p.kind != "UnstructuredGrid":
        raise VtkError(f"{path}: expected an UnstructuredGrid (.vtu)")
    return p.points, (p.ien() + 1).astype(np.int32), p


def read_vtp(path):
    """READVTP (S/VTKXML.f:83-150): face x, IEN (1-based GLOBAL node ids when GlobalNodeID is present,
    else local 1-based), gN (GlobalNodeID, 1-based) and gE (GlobalElementID, 1-based) or None."""
    p = read_vtk_xml(path)
    if p.kind != "PolyData":
        raise VtkError(f"{path}: expected PolyData (.vtp)")
    ien = p.ien()
    gN = p.point_data.get("GlobalNodeID")
    gE = p.cell_data.get("GlobalElementID")
    if gN is not None:
        gN = gN.astype(np.int32).reshape(-1)
        ien = gN[ien]
    else:
        ien = ien + 1
    return p.points, ien.astype(np.int32), gN, (None if gE is None else gE.astype(np.int32).reshape(-1)), p


# ------------------------------------------------------------------------------------ encoding
def _encode(a, mode, compress, hdt, block=1 << 15):
    raw = np.ascontiguousarray(a).tobytes()
    if compress:
        blocks = [raw[i:i + block] for i in range(0, len(raw), block)] or [b""]
        comp = [zlib.compress(b) for b in blocks]
        last = len(blocks[-1]) if len(blocks[-1]) != block else 0
        head = np.array([len(blocks), block, last] + [len(c) for c in comp], dtype=hdt).tobytes()
        body = b"".join(comp)
        if mode == "raw":
            return head + body
        return base64.b64encode(head) + base64.b64encode(body)
    head = np.array([len(raw)], dtype=hdt).tobytes()
    if mode == "raw":
        return head + raw
    return base64.b64encode(head + raw)


def _write(path, kind, points, conn, offs, types, point_data, cell_data, encoding, compress, header_type):
    if encoding not in ("appended", "appended-base64", "binary", "ascii"):
        raise VtkError(f"unknown encoding {encoding}")
    hdt = np.dtype("<" + _VTK2NP[header_type])
    appended, chunks = [], []

    def darray(name, a, ncomp=None):
        a = np.asarray(a)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        a = a.astype(a.dtype.newbyteorder("<"), copy=False)
        key = a.dtype.str[1:]
        if key not in _NP2VTK:
            raise VtkError(f"array {name}: dtype {a.dtype} has no VTK type")
        nc = ncomp if ncomp is not None else (a.shape[1] if a.ndim == 2 else 1)
        attrs = f'type="{_NP2VTK[key]}"' + (f' Name="{name}"' if name else "") + \
            (f' NumberOfComponents="{nc}"' if nc != 1 else "")
        if encoding == "ascii":
            body = " ".join(repr(float(v)) if a.dtype.kind == "f" else str(int(v)) for v in a.reshape(-1))
            return f'<DataArray {attrs} format="ascii">{body}</DataArray>\n'
        if encoding == "binary":
            return f'<DataArray {attrs} format="binary">{_encode(a, "b64", compress, hdt).decode()}</DataArray>\n'
        mode = "raw" if encoding == "appended" else "b64"
        off = sum(len(c) for c in appended)
        appended.append(_encode(a, mode, compress, hdt))
        return f'<DataArray {attrs} format="appended" offset="{off}"/>\n'

    def section(tag, data):
        if not data:
            return ""
        s = f"<{tag}>\n"
        for k, v in data.items():
            s += darray(k, v)
        return s + f"</{tag}>\n"

    nNo, nEl = points.shape[0], offs.size
    comp_attr = ' compressor="vtkZLibDataCompressor"' if (compress and encoding != "ascii") else ""
    s = (f'<?xml version="1.0"?>\n<VTKFile type="{kind}" version="0.1" byte_order="LittleEndian" '
         f'header_type="{header_type}"{comp_attr}>\n<{kind}>\n')
    if kind == "UnstructuredGrid":
        s += f'<Piece NumberOfPoints="{nNo}" NumberOfCells="{nEl}">\n'
    else:
        s += (f'<Piece NumberOfPoints="{nNo}" NumberOfVerts="0" NumberOfLines="0" NumberOfStrips="0" '
              f'NumberOfPolys="{nEl}">\n')
    s += section("PointData", point_data) + section("CellData", cell_data)
    s += "<Points>\n" + darray("Points", np.asarray(points, dtype=np.float64), 3) + "</Points>\n"
    tag = "Cells" if kind == "UnstructuredGrid" else "Polys"
    s += f"<{tag}>\n" + darray("connectivity", conn.astype(np.int64)) + darray("offsets", offs.astype(np.int64))
    if kind == "UnstructuredGrid":
        s += darray("types", types.astype(np.uint8))
    s += f"</{tag}>\n</Piece>\n</{kind}>\n"
    with open(path, "wb") as fh:
        fh.write(s.encode())
        if appended:
            enc = "raw" if encoding == "appended" else "base64"
            fh.write(f'<AppendedData encoding="{enc}">\n_'.encode())
            for c in appended:
                fh.write(c)
            fh.write(b"\n</AppendedData>\n")
        fh.write(b"</VTKFile>\n")


def write_vtu(path, x, IEN, point_data=None, cell_data=None, cell_type=VTK_TETRA, encoding="appended",
              compress=False, header_type="UInt64"):
    """WRITEVTUS's file format (S/VTKXML.f:411-1000 writes one piece with PointData results).
    IEN is 1-based (svFSI); the file stores 0-based ids."""
    IEN = np.asarray(IEN)
    nEl, eNoN = IEN.shape
    if _NODES_OF.get(cell_type) != eNoN:
        raise VtkError(f"cell type {cell_type} does not have {eNoN} nodes")
    conn = (IEN.astype(np.int64) - 1).reshape(-1)
    offs = np.arange(1, nEl + 1, dtype=np.int64) * eNoN
    types = np.full(nEl, cell_type, dtype=np.uint8)
    _write(path, "UnstructuredGrid", np.asarray(x, dtype=np.float64), conn, offs, types,
           point_data or {}, cell_data or {}, encoding, compress, header_type)


def write_vtp(path, x, IEN_local, gN=None, gE=None, encoding="appended", compress=False,
              header_type="UInt64"):
    """A boundary face as svFSI-Tests' mesh-surfaces/*.vtp hold it: local 1-based connectivity into
    the face's own points plus GlobalNodeID / GlobalElementID (1-based) arrays."""
    IEN_local = np.asarray(IEN_local)
    nEl, eNoN = IEN_local.shape
    conn = (IEN_local.astype(np.int64) - 1).reshape(-1)
    offs = np.arange(1, nEl + 1, dtype=np.int64) * eNoN
    pd = {} if gN is None else {"GlobalNodeID": np.asarray(gN, dtype=np.int32)}
    cd = {} if gE is None else {"GlobalElementID": np.asarray(gE, dtype=np.int32)}
    _write(path, "PolyData", np.asarray(x, dtype=np.float64), conn, offs, None, pd, cd, encoding,
           compress, header_type)


# ------------------------------------------------------------------------------------ mesh-complete
def write_mesh_complete(m, dirname, encoding="appended", compress=False):
    """Write a svfsi_b200.mesh.Mesh as a svFSI-Tests style directory: mesh-complete.mesh.vtu and
    mesh-surfaces/<face>.vtp (what `Mesh file path` / `Face file path` of the input deck name,
    S/READMSH.f -> LOADMSH.f:39-91)."""
    os.makedirs(os.path.join(dirname, "mesh-surfaces"), exist_ok=True)
    gid = np.arange(1, m.nNo + 1, dtype=np.int32)
    eid = np.arange(1, m.nEl + 1, dtype=np.int32)
    write_vtu(os.path.join(dirname, "mesh-complete.mesh.vtu"), m.x, m.IEN,
              point_data={"GlobalNodeID": gid}, cell_data={"GlobalElementID": eid}, encoding=encoding,
              compress=compress)
    for name, fa in m.faces.items():
        gN = np.asarray(fa.gN, dtype=np.int64)
        loc = np.zeros(m.nNo + 1, dtype=np.int64)
        loc[gN] = np.arange(1, gN.size + 1)
        write_vtp(os.path.join(dirname, "mesh-surfaces", name + ".vtp"), m.x[gN - 1], loc[fa.tri.astype(np.int64)],
                  gN=gN, gE=np.asarray(fa.parent, dtype=np.int64) + 1, encoding=encoding, compress=compress)


def read_mesh_complete(dirname):
    """-> (x (nNo,3), IEN (nEl,4) 1-based, faces: name -> dict(gN, IEN (global 1-based), gE (1-based)))"""
    x, IEN, _ = read_vtu(os.path.join(dirname, "mesh-complete.mesh.vtu"))
    faces = {}
    sdir = os.path.join(dirname, "mesh-surfaces")
    for fn in sorted(os.listdir(sdir)):
        if not fn.endswith(".vtp"):
            continue
        fx, fien, gN, gE, _ = read_vtp(os.path.join(sdir, fn))
        if gN is None or gE is None:
            raise VtkError(f"{fn}: svFSI needs GlobalNodeID and GlobalElementID on mesh faces")
        if np.abs(x[gN - 1] - fx).max() > 1e-12 * max(1.0, np.abs(x).max()):
            raise VtkError(f"{fn}: face coordinates do not match the volume mesh at GlobalNodeID")
        faces[fn[:-4]] = dict(gN=gN, IEN=fien, gE=gE)
    return x, IEN, faces


# ------------------------------------------------------------------------------------ restart record
def restart_reclen(nEq, nX, tDof, tnNo, dFlag=False):
    """recLn of S/INITIALIZE.f:155-170 for the non-prestress, non-CEP, non-IB cases (IKIND=4, RKIND=8)"""
    i = 3 * tDof if dFlag else 2 * tDof
    return 4 * (1 + 7) + 8 * (2 + nEq + nX + i * tnNo)


def write_restart(path, rank, recLn, stamp, cTS, time, timeP, iNorm, xn, Yn, An, Dn=None):
    """One rank's record of WRITERESTART (S/OUTPUT.f:132-232): direct access, record `rank` (1-based)
    of length recLn: stamp(7) int32, cTS int32, time, timeP(1), eq%iNorm(nEq), cplBC%xn(nX),
    Yn(tDof,tnNo), An(tDof,tnNo)[, Dn].  Arrays are (tnNo, tDof) row-major = Fortran (tDof,tnNo)."""
    stamp = np.asarray(stamp, dtype="<i4")
    if stamp.size != 7:
        raise ValueError("stamp = (/np, nEq, nMsh, tnNo, nX, tDof, dFlag/)")
    rec = stamp.tobytes() + struct.pack("<i", int(cTS)) + struct.pack("<dd", float(time), float(timeP))
    rec += np.asarray(iNorm, dtype="<f8").tobytes() + np.asarray(xn, dtype="<f8").tobytes()
    rec += np.ascontiguousarray(Yn, dtype="<f8").tobytes() + np.ascontiguousarray(An, dtype="<f8").tobytes()
    if Dn is not None:
        rec += np.ascontiguousarray(Dn, dtype="<f8").tobytes()
    if len(rec) > recLn:
        raise ValueError(f"record of {len(rec)} bytes exceeds recLn={recLn}")
    mode = "r+b" if os.path.exists(path) else "w+b"
    with open(path, mode) as fh:
        fh.seek((rank - 1) * recLn)
        fh.write(rec + b"\0" * (recLn - len(rec)))


def read_restart(path, rank, recLn, nEq, nX, tDof, tnNo, dFlag=False, expect_stamp=None):
    """INITFROMBIN (S/INITIALIZE.f:512-620) for the same cases; checks the stamp like :593-617."""
    with open(path, "rb") as fh:
        fh.seek((rank - 1) * recLn)
        rec = fh.read(recLn)
    need = restart_reclen(nEq, nX, tDof, tnNo, dFlag)
    if len(rec) < need:
        raise ValueError(f"{path}: record {rank} is {len(rec)} bytes, need {need}")
    at = 0
    stamp = np.frombuffer(rec, "<i4", 7, at); at += 28
    cTS = struct.unpack_from("<i", rec, at)[0]; at += 4
    time, timeP = struct.unpack_from("<dd", rec, at); at += 16
    iNorm = np.frombuffer(rec, "<f8", nEq, at).copy(); at += 8 * nEq
    xo = np.frombuffer(rec, "<f8", nX, at).copy(); at += 8 * nX
    n = tDof * tnNo
    Yo = np.frombuffer(rec, "<f8", n, at).reshape(tnNo, tDof).copy(); at += 8 * n
    Ao = np.frombuffer(rec, "<f8", n, at).reshape(tnNo, tDof).copy(); at += 8 * n
    Do = None
    if dFlag:
        Do = np.frombuffer(rec, "<f8", n, at).reshape(tnNo, tDof).copy()
    if expect_stamp is not None and not np.array_equal(stamp, np.asarray(expect_stamp, dtype="<i4")):
        names = ("Number of processors", "Number of equations", "Number of meshes", "Number of nodes",
                 "Number of cplBC%x", "Number of dof", "dFlag specification")
        bad = [nm for nm, a, b in zip(names, stamp, expect_stamp) if a != b]
        raise ValueError(f"{path}: simulation stamp does not match: {', '.join(bad)}")
    return dict(stamp=stamp.copy(), cTS=cTS, time=time, timeP=timeP, iNorm=iNorm, xo=xo, Yo=Yo, Ao=Ao, Do=Do)


if sys.byteorder != "little":  # pragma: no cover
    raise ImportError("vtkio assumes a little-endian host")
