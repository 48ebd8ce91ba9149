"""VTU import / export: the package's own parser against files encoded the ways VTK and meshio write
them -- ascii, inline base64 (byte count encoded on its own, or in one stream with the data), appended raw and appended
base64, with and without vtkZLibDataCompressor, UInt32 / UInt64 headers -- built here byte by byte from the VTK XML file
format description, independently of the package's writer; then the data_io functions of the reference on top
(data_io.py:423-654)."""
import base64
import os
import struct
import zlib

import numpy as np
import pytest

from glimslib_b200.backend import vtu
from glimslib_b200.utils import data_io as dio
from glimslib_b200 import fenics_local as fenics

PTS = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [9, 9, 9]], dtype=np.float64)   # vertex 5 is orphaned
TETS = np.array([[0, 1, 2, 3], [1, 2, 3, 4]], dtype=np.int64)
LABEL = np.array([3, 7], dtype=np.int32)
TEMP = np.linspace(0.0, 1.0, len(PTS))


def _arrays():
    return [("Points", None, PTS.astype("<f8"), 3), ("Cells", "connectivity", TETS.ravel().astype("<i8"), 1),
            ("Cells", "offsets", np.array([4, 8], dtype="<i8"), 1), ("Cells", "types", np.array([10, 10], dtype="u1"), 1),
            ("CellData", "ElementBlockIds", LABEL.astype("<i4"), 1), ("PointData", "temp", TEMP.astype("<f8"), 1)]


_T = {"f8": "Float64", "i8": "Int64", "i4": "Int32", "u1": "UInt8"}


def _payload(a, hfmt, compress):
    raw = a.tobytes()
    if not compress:
        return struct.pack(hfmt, len(raw)), raw
    bs = 32          # tiny blocks so that several blocks occur
    blocks = [raw[i:i + bs] for i in range(0, len(raw), bs)]
    comp = [zlib.compress(b) for b in blocks]
    head = struct.pack(hfmt[0] + hfmt[1] * (3 + len(comp)), len(blocks), bs, len(blocks[-1]) if len(raw) % bs else bs if blocks else 0,
                       *[len(c) for c in comp])
    return head, b"".join(comp)


def _file(tmp_path, mode, header, compress, split_header=True):
    hfmt = "<I" if header == "UInt32" else "<Q"
    sections = {"Points": [], "Cells": [], "CellData": [], "PointData": []}
    appended = b""
    for sec, name, a, ncomp in _arrays():
        attrs = 'type="%s"' % _T[a.dtype.str[1:]]
        if name:
            attrs += ' Name="%s"' % name
        if ncomp > 1:
            attrs += ' NumberOfComponents="%d"' % ncomp
        head, data = _payload(a, hfmt, compress)
        if mode == "ascii":
            body = " ".join(repr(float(v)) if a.dtype.kind == "f" else str(int(v)) for v in a.ravel())
            sections[sec].append('<DataArray %s format="ascii">%s</DataArray>' % (attrs, body))
        elif mode == "binary":
            if compress or split_header:
                body = base64.b64encode(head) + base64.b64encode(data)
            else:
                body = base64.b64encode(head + data)
            sections[sec].append('<DataArray %s format="binary">%s</DataArray>' % (attrs, body.decode()))
        else:
            sections[sec].append('<DataArray %s format="appended" offset="%d"/>' % (attrs, len(appended)))
            if mode == "appended_raw":
                appended += head + data
            else:
                appended += (base64.b64encode(head) + base64.b64encode(data)) if (compress or split_header) else base64.b64encode(head + data)
    comp_attr = ' compressor="vtkZLibDataCompressor"' if compress else ''
    xml = ('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="%s"%s>\n'
           '<UnstructuredGrid><Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (header, comp_attr, len(PTS), len(TETS)))
    for sec in ("PointData", "CellData", "Points", "Cells"):
        xml += "<%s>%s</%s>\n" % (sec, "\n".join(sections[sec]), sec)
    xml += "</Piece></UnstructuredGrid>\n"
    blob = xml.encode()
    if mode.startswith("appended"):
        enc = "raw" if mode == "appended_raw" else "base64"
        blob += b'<AppendedData encoding="' + enc.encode() + b'">\n  _' + appended + b'\n</AppendedData>\n'
    blob += b"</VTKFile>\n"
    p = os.path.join(str(tmp_path), "m_%s_%s_%d_%d.vtu" % (mode, header, compress, split_header))
    with open(p, "wb") as f:
        f.write(blob)
    return p


@pytest.mark.parametrize("mode,header,compress,split", [
    ("ascii", "UInt32", False, True), ("binary", "UInt32", False, True), ("binary", "UInt64", False, False),
    ("binary", "UInt64", True, True), ("appended_raw", "UInt32", False, True), ("appended_raw", "UInt64", True, True),
    ("appended_b64", "UInt32", False, True), ("appended_b64", "UInt32", True, True), ("appended_b64", "UInt64", False, False)])
def test_reader_handles_every_encoding(tmp_path, mode, header, compress, split):
    m = vtu.read_vtu(_file(tmp_path, mode, header, compress, split))
    assert np.array_equal(m.points, PTS)
    assert np.array_equal(m.cells["tetrahedron"], TETS)
    assert np.array_equal(m.cell_data["tetrahedron"]["ElementBlockIds"], LABEL)
    assert np.allclose(m.point_data["temp"], TEMP, rtol=0, atol=0)


@pytest.mark.parametrize("binary", [False, True])
def test_writer_round_trip_and_fenics_conversion(tmp_path, binary):
    src = vtu.VtuMesh(PTS, {"tetrahedron": TETS}, {"temp": TEMP, "vec": np.arange(18.0).reshape(6, 3)},
                      {"tetrahedron": {"ElementBlockIds": LABEL}})
    p = os.path.join(str(tmp_path), "w.vtu")
    vtu.write_vtu(p, src, binary=binary)
    m = vtu.read_vtu(p)
    assert np.array_equal(m.points, PTS) and np.array_equal(m.cells["tetrahedron"], TETS)
    assert np.array_equal(m.point_data["vec"], src.point_data["vec"])
    # read_vtk_convert_to_fenics: orphaned vertex 5 removed, labels carried (data_io.py:470-524, 577-581)
    mesh, sub = dio.read_vtk_convert_to_fenics(p)
    assert mesh.num_vertices() == 5 and mesh.num_cells() == 2 and mesh.geometry().dim() == 3
    assert list(sub.array()) == [3, 7]
    assert dio.identify_orphaned_vertices(mesh) == []
    back = dio.convert_fenics_mesh_to_meshio(mesh, subdomains=sub)
    assert np.array_equal(back.cells["tetrahedron"], TETS) and list(back.cell_data["tetrahedron"]["ElementBlockIds"]) == [3, 7]
    sel_mesh, sel_sub = dio.remove_mesh_subdomain(mesh, sub, 7, 7)
    assert sel_mesh.num_cells() == 1 and sel_mesh.num_vertices() == 4 and list(sel_sub.array()) == [7]


def test_2d_vtu_drops_zero_third_coordinate(tmp_path):
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=float)
    tri = np.array([[0, 1, 2], [1, 3, 2]])
    p = os.path.join(str(tmp_path), "t.vtu")
    vtu.write_vtu(p, vtu.VtuMesh(pts, {"triangle": tri}, cell_data={"triangle": {"ElementBlockIds": np.array([1, 2])}}))
    mesh, sub = dio.read_vtk_convert_to_fenics(p)
    assert mesh.geometry().dim() == 2 and mesh.coordinates().shape == (4, 2) and list(sub.array()) == [1, 2]


def test_mesh_editor_and_merge_vtus(tmp_path):
    """fenics.MeshEditor as data_io.py:458-468 uses it, and the post-run merge of per-field VTUs with the label map
    (data_io.py:606-654) on files written by the backend's own `File(...) << (function, t)`."""
    mesh = fenics.Mesh()
    ed = fenics.MeshEditor()
    ed.open(mesh, "triangle", 2, 2)
    ed.init_vertices(4)
    ed.init_cells(2)
    for i, x in enumerate([(0, 0), (1, 0), (0, 1), (1, 1)]):
        ed.add_vertex(i, np.array(x, dtype=float))
    ed.add_cell(0, np.array([0, 1, 2], dtype=np.uintp))
    ed.add_cell(1, np.array([1, 3, 2], dtype=np.uintp))
    ed.close()
    assert mesh.num_cells() == 2 and mesh.geometry().dim() == 2
    base = str(tmp_path)
    V = fenics.FunctionSpace(mesh, "Lagrange", 1)
    labels = fenics.MeshFunction("size_t", mesh, 2)
    labels.array()[:] = [4, 5]
    fenics.File(os.path.join(base, "label_map", "label_map_00000.pvd")) << labels
    for step in range(3):
        c = fenics.Function(V)
        c.vector()[:] = np.arange(4.0) + step
        c.rename("concentration", "label")
        fenics.File(os.path.join(base, "concentration", "concentration_%05d.pvd" % step)) << (c, float(step))
    dio.merge_VTUs(base, 1, 2, remove=True, reference=None)
    for step in range(3):
        m = vtu.read_vtu(os.path.join(base, "merged", dio.create_file_name("all", step)))
        assert np.allclose(np.asarray(m.point_data["concentration"]).ravel(), np.arange(4.0) + step)
        assert len(m.cells["triangle"]) == 2
        assert not os.path.exists(os.path.join(base, "concentration", dio.create_file_name("concentration", step)))
