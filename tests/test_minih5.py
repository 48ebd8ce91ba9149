"""CPU tests of the from-scratch HDF5 writer/reader (no libhdf5 / h5py exists here, so the check is structural:
signature, version-0 superblock fields, and a full round trip through the independent parsing code path)."""
import struct

import numpy as np

from glimslib_b200.backend import minih5
from glimslib_b200 import fenics_local as fenics


def _tree():
    root = minih5.Group()
    root.create_dataset("solution/vector_0", np.arange(12.0)).attrs["timestamp"] = 0.0
    root.create_dataset("solution/vector_1", np.arange(12.0) * 2).attrs["timestamp"] = 1.5
    root.get("solution").attrs["count"] = np.uint64(2)
    root.create_dataset("Mesh/mesh/topology", np.arange(24, dtype=np.int32).reshape(6, 4))
    root.create_dataset("Mesh/mesh/geometry", np.linspace(0, 1, 15).reshape(5, 3))
    root.attrs["title"] = "glims"
    for k in range(40):                      # more entries than the default leaf K: node sizes must adapt
        root.create_dataset("many/v_%02d" % k, np.full(3, k, dtype=np.int64))
    return root


def test_superblock_and_signature(tmp_path):
    p = str(tmp_path / "a.h5")
    minih5.write_file(p, _tree())
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8] == 0 and b[13] == 8 and b[14] == 8                 # superblock v0, 8-byte offsets and lengths
    leaf_k, internal_k = struct.unpack_from("<HH", b, 16)
    assert leaf_k >= 20 and internal_k == 16
    base, _, eof, _ = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and eof == len(b)
    root_oh = struct.unpack_from("<Q", b, 64)[0]
    assert b[root_oh] == 1                                          # version-1 object header
    assert b.count(b"TREE") >= 4 and b.count(b"SNOD") >= 4 and b.count(b"HEAP") >= 4


def test_round_trip(tmp_path):
    p = str(tmp_path / "a.h5")
    minih5.write_file(p, _tree())
    r = minih5.read_file(p)
    assert np.array_equal(r.get("solution/vector_1").data, np.arange(12.0) * 2)
    assert r.get("solution").attrs["count"] == 2
    assert r.get("solution/vector_1").attrs["timestamp"] == 1.5
    t = r.get("Mesh/mesh/topology").data
    assert t.shape == (6, 4) and t.dtype == np.int32 and t[5, 3] == 23
    assert r.attrs["title"] == "glims"
    assert sorted(r.get("many").children) == ["v_%02d" % k for k in range(40)]
    assert np.array_equal(r.get("many/v_17").data, [17, 17, 17])


def test_hdf5file_time_series_layout(tmp_path):
    """DOLFIN convention used by the reference (helper_classes.py:1256-1308): /<name>/vector_<k> + timestamp, count."""
    mesh = fenics.UnitSquareMesh(3, 3)
    V = fenics.FunctionSpace(mesh, "CG", 1)
    f = fenics.Function(V)
    p = str(tmp_path / "ts.h5")
    h = fenics.HDF5File(None, p, "w")
    for k in range(3):
        f.vector()[:] = float(k)
        h.write(f, "solution", float(k) * 0.5)
    h.close()
    raw = minih5.read_file(p)
    g = raw.get("solution")
    assert int(g.attrs["count"]) == 3
    assert {"vector_0", "vector_1", "vector_2", "cells", "cell_dofs", "x_cell_dofs"} <= set(g.children)
    assert g.children["vector_2"].attrs["timestamp"] == 1.0
    h2 = fenics.HDF5File(None, p, "r")
    assert h2.attributes("solution")["count"] == 3
    g2 = fenics.Function(V)
    h2.read(g2, "solution/vector_1")
    assert np.all(g2.vector().get_local() == 1.0)
