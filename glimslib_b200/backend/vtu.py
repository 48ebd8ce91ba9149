"""Reader / writer of VTK XML unstructured grids (``.vtu``) -- the mesh exchange format on either side of the hot path
(SURVEY.md 8f row N3): atlas and patient meshes arrive as VTU (``glimslib/utils/data_io.py:469-524``,
``vtk_utils.read_vtk_data``), simulation output leaves as VTU (``helper_classes.py:1376-1380``, ``data_io.py:606-654``).

The reference goes through ``vtk`` and ``meshio``; neither is a dependency here, so this module parses the format itself:
``ascii``, inline ``binary`` (base64) and ``appended`` (``raw`` or ``base64``) data arrays, ``UInt32`` / ``UInt64`` block
headers, optional ``vtkZLibDataCompressor``, little or big endian.  Only what the path needs is interpreted: points, cells
(connectivity / offsets / types), point data and cell data.
"""
import base64
import os
import re
import struct
import xml.etree.ElementTree as ET
import zlib

import numpy as np

VTK_TRIANGLE, VTK_TETRA = 5, 10
_DTYPES = {"Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2", "Int32": "i4", "UInt32": "u4",
           "Int64": "i8", "UInt64": "u8", "Float32": "f4", "Float64": "f8"}
_NAMES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


class VtuMesh:
    """meshio-shaped container: ``points [n, 3]``, ``cells {'triangle' | 'tetrahedron': [nc, k]}``,
    ``point_data {name: array}``, ``cell_data {cell_type: {name: array}}``."""

    def __init__(self, points, cells, point_data=None, cell_data=None):
        self.points = np.asarray(points)
        self.cells = dict(cells)
        self.point_data = dict(point_data or {})
        self.cell_data = {k: dict(v) for k, v in (cell_data or {}).items()}


def _decode_blocks(raw, header_dtype, compressed, b64, offset=0):
    """One data array's bytes from a (possibly base64, possibly zlib-blocked) payload starting at ``offset``.
    Returns (bytes, bytes consumed from ``raw``)."""
    hsize = np.dtype(header_dtype).itemsize

    def take(n_bytes, at):
        """n_bytes decoded bytes starting at decoded position 0 of the stream that begins at raw[at]"""
        if not b64:
            return raw[at:at + n_bytes], n_bytes
        n_chars = (n_bytes + 2) // 3 * 4
        return base64.b64decode(raw[at:at + n_chars])[:n_bytes], n_chars

    if not compressed:
        if not b64:
            n = int(np.frombuffer(raw[offset:offset + hsize], dtype=header_dtype)[0])
            return raw[offset + hsize:offset + hsize + n], hsize + n
        hchars = (hsize + 2) // 3 * 4
        hchunk = raw[offset:offset + hchars]
        n = int(np.frombuffer(base64.b64decode(hchunk)[:hsize], dtype=header_dtype)[0])
        if hchunk.endswith(b"="):
            # VTK encodes the byte count on its own (padded), the data follow as a second base64 stream
            dchars = (n + 2) // 3 * 4
            return base64.b64decode(raw[offset + hchars:offset + hchars + dchars])[:n], hchars + dchars
        total = hsize + n           # one stream: header and data encoded together
        n_chars = (total + 2) // 3 * 4
        buf = base64.b64decode(raw[offset:offset + n_chars])
        return buf[hsize:hsize + n], n_chars
    # compressed: [n_blocks, block_size, last_block_size, csize_0 .. csize_{n-1}] then the compressed blocks
    head, used = take(3 * hsize, offset)
    nb, bs, last = (int(v) for v in np.frombuffer(head, dtype=header_dtype))
    if nb == 0:
        return b"", used
    if b64:
        hbytes = (3 + nb) * hsize
        hchars = (hbytes + 2) // 3 * 4
        hdr = np.frombuffer(base64.b64decode(raw[offset:offset + hchars])[:hbytes], dtype=header_dtype)
        csizes = [int(v) for v in hdr[3:]]
        total = sum(csizes)
        dchars = (total + 2) // 3 * 4
        data = base64.b64decode(raw[offset + hchars:offset + hchars + dchars])
        consumed = hchars + dchars
    else:
        hdr = np.frombuffer(raw[offset:offset + (3 + nb) * hsize], dtype=header_dtype)
        csizes = [int(v) for v in hdr[3:]]
        total = sum(csizes)
        data = raw[offset + (3 + nb) * hsize:offset + (3 + nb) * hsize + total]
        consumed = (3 + nb) * hsize + total
    out, pos = [], 0
    for cs in csizes:
        out.append(zlib.decompress(data[pos:pos + cs]))
        pos += cs
    return b"".join(out), consumed


def read_vtu(path):
    """Parse ``path`` into a :class:`VtuMesh` (triangles or tetrahedra; mixed grids keep the most frequent of the two)."""
    with open(path, "rb") as f:
        blob = f.read()
    appended_raw = None
    m = re.search(rb"<AppendedData[^>]*encoding\s*=\s*\"raw\"[^>]*>", blob)
    if m:
        # raw appended data is not XML: cut it out before parsing ('_' marks its first byte)
        start = blob.index(b"_", m.end()) + 1
        end = blob.rindex(b"</AppendedData>")
        appended_raw = blob[start:end]
        blob = blob[:m.end()] + b"_" + blob[end:]
    root = ET.fromstring(blob)
    if root.tag != "VTKFile" or root.get("type") != "UnstructuredGrid":
        raise ValueError("%s: not a VTK UnstructuredGrid file" % path)
    order = "<" if root.get("byte_order", "LittleEndian") == "LittleEndian" else ">"
    header_dtype = np.dtype(order + _DTYPES[root.get("header_type", "UInt32")])
    compressed = root.get("compressor") is not None
    app = root.find("AppendedData")
    app_b64 = None
    if app is not None and appended_raw is None:
        txt = (app.text or "").strip()
        app_b64 = txt[txt.index("_") + 1:].encode() if "_" in txt else b""

    def array(da):
        dt = np.dtype(order + _DTYPES[da.get("type")])
        ncomp = int(da.get("NumberOfComponents", "1"))
        fmt = da.get("format", "ascii")
        if fmt == "ascii":
            a = np.array((da.text or "").split(), dtype=np.float64 if dt.kind == "f" else np.int64).astype(dt.newbyteorder("="))
        elif fmt == "binary":
            raw = "".join((da.text or "").split()).encode()
            buf, _ = _decode_blocks(raw, header_dtype, compressed, True)
            a = np.frombuffer(buf, dtype=dt).astype(dt.newbyteorder("="))
        elif fmt == "appended":
            off = int(da.get("offset", "0"))
            if appended_raw is not None:
                buf, _ = _decode_blocks(appended_raw, header_dtype, compressed, False, off)
            else:
                buf, _ = _decode_blocks(app_b64, header_dtype, compressed, True, off)
            a = np.frombuffer(buf, dtype=dt).astype(dt.newbyteorder("="))
        else:
            raise ValueError("unknown DataArray format %r" % fmt)
        return a.reshape(-1, ncomp) if ncomp > 1 else a

    piece = root.find("UnstructuredGrid/Piece")
    if piece is None:
        raise ValueError("%s: no <Piece>" % path)
    points = array(piece.find("Points/DataArray")).reshape(-1, 3).astype(np.float64)
    cell_arrays = {da.get("Name"): array(da) for da in piece.findall("Cells/DataArray")}
    conn, offs, types = cell_arrays["connectivity"], cell_arrays["offsets"], cell_arrays["types"]
    counts = {t: int((types == t).sum()) for t in (VTK_TRIANGLE, VTK_TETRA)}
    vt = VTK_TETRA if counts[VTK_TETRA] >= counts[VTK_TRIANGLE] and counts[VTK_TETRA] > 0 else VTK_TRIANGLE
    if counts[vt] == 0:
        raise ValueError("%s: no triangle or tetrahedron cells" % path)
    k = 3 if vt == VTK_TRIANGLE else 4
    keep = np.nonzero(types == vt)[0]
    starts = np.concatenate([[0], offs[:-1]]).astype(np.int64)[keep]
    cells = conn.astype(np.int64)[starts[:, None] + np.arange(k)[None, :]]
    name = "triangle" if vt == VTK_TRIANGLE else "tetrahedron"
    point_data = {da.get("Name"): array(da) for da in piece.findall("PointData/DataArray")}
    cell_data = {da.get("Name"): array(da)[keep] for da in piece.findall("CellData/DataArray")}
    return VtuMesh(points, {name: cells}, point_data, {name: cell_data} if cell_data else {})


def write_vtu(path, mesh, binary=False):
    """Write a :class:`VtuMesh` (one cell type) as ASCII or inline-binary (base64, UInt64 headers, uncompressed) VTU."""
    (name, cells), = mesh.cells.items()
    vt = VTK_TRIANGLE if name == "triangle" else VTK_TETRA
    cells = np.asarray(cells)
    nc, k = cells.shape
    pts = np.zeros((len(mesh.points), 3))
    pts[:, :mesh.points.shape[1]] = mesh.points

    def da(a, nm=None, ncomp=None):
        a = np.ascontiguousarray(a)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        tname = _NAMES[a.dtype.str[1:]]
        attrs = 'type="%s"' % tname
        if nm:
            attrs += ' Name="%s"' % nm
        nco = ncomp if ncomp is not None else (a.shape[1] if a.ndim == 2 else 1)
        if nco > 1:
            attrs += ' NumberOfComponents="%d"' % nco
        if binary:
            raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
            body = base64.b64encode(struct.pack("<Q", len(raw)) + raw).decode()
            return '<DataArray %s format="binary">%s</DataArray>\n' % (attrs, body)
        fmt = "%.17g" if a.dtype.kind == "f" else "%d"
        rows = a.reshape(len(a), -1)
        body = "\n".join(" ".join(fmt % v for v in r) for r in rows)
        return '<DataArray %s format="ascii">\n%s\n</DataArray>\n' % (attrs, body)

    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" '
                'header_type="UInt64">\n<UnstructuredGrid>\n')
        f.write('<Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (len(pts), nc))
        f.write("<Points>\n" + da(pts, ncomp=3) + "</Points>\n<Cells>\n")
        f.write(da(cells.astype(np.int64).ravel(), "connectivity"))
        f.write(da(((np.arange(nc) + 1) * k).astype(np.int64), "offsets"))
        f.write(da(np.full(nc, vt, dtype=np.uint8), "types"))
        f.write("</Cells>\n")
        if mesh.point_data:
            f.write("<PointData>\n")
            for nm, a in mesh.point_data.items():
                f.write(da(np.asarray(a), nm))
            f.write("</PointData>\n")
        cd = mesh.cell_data.get(name, {})
        if cd:
            f.write("<CellData>\n")
            for nm, a in cd.items():
                f.write(da(np.asarray(a), nm))
            f.write("</CellData>\n")
        f.write("</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
