"""The HDF5 writer against the file-format specification, through a verifier that shares no code with it (tests/h5spec.py):
every structure from the superblock to the raw data is re-parsed from the published layout and checked for what the
specification requires; the datatype messages are compared with the byte strings libhdf5 emits for the same native types
(H5T_IEEE_F64LE, H5T_STD_I32LE, ... -- known answers from the specification's field tables).  SURVEY.md 8f row N1; the
reference writes these files through DOLFIN's HDF5File (helper_classes.py:1256-1308)."""
import struct

import numpy as np
import pytest

import h5spec
from glimslib_b200.backend import minih5
from glimslib_b200 import fenics_local as fenics


def _tree(n_many=40):
    root = minih5.Group()
    root.create_dataset("solution/vector_0", np.arange(12.0)).attrs["timestamp"] = 0.0
    root.create_dataset("solution/vector_1", np.arange(12.0) * 2).attrs["timestamp"] = 1.5
    root.get("solution").attrs["count"] = np.uint64(2)
    root.create_dataset("Mesh/mesh/topology", np.arange(24, dtype=np.int32).reshape(6, 4))
    root.create_dataset("Mesh/mesh/geometry", np.linspace(0, 1, 15).reshape(5, 3))
    root.attrs["title"] = "glims"
    for k in range(n_many):
        root.create_dataset("many/v_%03d" % k, np.full(3, k, dtype=np.int64))
    return root


def _bytes(tmp_path, root, name="a.h5"):
    p = str(tmp_path / name)
    minih5.write_file(p, root)
    return open(p, "rb").read()


# class/version, class bit field (3 bytes), size (4), properties -- IV.A.2.d of the specification
KNOWN_DATATYPES = {
    "<f8": bytes([0x11, 0x20, 0x3F, 0x00, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52]) + struct.pack("<I", 1023),
    "<f4": bytes([0x11, 0x20, 0x1F, 0x00, 4, 0, 0, 0, 0, 0, 32, 0, 23, 8, 0, 23]) + struct.pack("<I", 127),
    "<i4": bytes([0x10, 0x08, 0x00, 0x00, 4, 0, 0, 0, 0, 0, 32, 0]),
    "<i8": bytes([0x10, 0x08, 0x00, 0x00, 8, 0, 0, 0, 0, 0, 64, 0]),
    "<u8": bytes([0x10, 0x00, 0x00, 0x00, 8, 0, 0, 0, 0, 0, 64, 0]),
}


@pytest.mark.parametrize("dt", sorted(KNOWN_DATATYPES))
def test_datatype_messages_are_the_known_answers(dt):
    got = minih5._datatype_message(np.dtype(dt))
    want = KNOWN_DATATYPES[dt]
    assert got[:len(want)] == want and not any(got[len(want):])          # zero padding to the 8-byte message granule only
    assert h5spec.Walker.datatype(got) == np.dtype(dt)


def test_every_structure_of_a_written_file_satisfies_the_specification(tmp_path):
    data, groups = h5spec.verify(_bytes(tmp_path, _tree()))
    assert np.array_equal(data["/solution/vector_1"][0], np.arange(12.0) * 2)
    assert data["/solution/vector_1"][1] == {"timestamp": 1.5}
    t = data["/Mesh/mesh/topology"][0]
    assert t.dtype == np.int32 and t.shape == (6, 4) and t[5, 3] == 23
    assert np.allclose(data["/Mesh/mesh/geometry"][0], np.linspace(0, 1, 15).reshape(5, 3))
    assert groups["/"] == {"title": "glims"} and groups["/solution"]["count"] == 2
    assert sorted(k for k in data if k.startswith("/many/")) == ["/many/v_%03d" % k for k in range(40)]
    assert np.array_equal(data["/many/v_017"][0], [17, 17, 17])


@pytest.mark.parametrize("n", [1, 7, 33, 300])
def test_groups_of_any_size(tmp_path, n):
    """Symbol-table nodes hold at most 2 * leaf K entries and B-tree nodes 2 * internal K children: large groups need more of
    both, and the keys must still bracket the children."""
    data, _ = h5spec.verify(_bytes(tmp_path, _tree(n)))
    assert sum(k.startswith("/many/") for k in data) == n


def test_dolfin_style_time_series(tmp_path):
    mesh = fenics.UnitSquareMesh(3, 3)
    V = fenics.FunctionSpace(mesh, "CG", 1)
    f = fenics.Function(V)
    p = str(tmp_path / "ts.h5")
    h = fenics.HDF5File(None, p, "w")
    for k in range(3):
        f.vector()[:] = float(k)
        h.write(f, "solution", float(k) * 0.5)
    h.close()
    data, groups = h5spec.verify(open(p, "rb").read())
    assert int(groups["/solution"]["count"]) == 3
    assert data["/solution/vector_2"][1]["timestamp"] == 1.0
    assert np.all(data["/solution/vector_1"][0] == 1.0)


def _corrupt(b, where, new):
    return b[:where] + new + b[where + len(new):]


def test_the_verifier_rejects_broken_files(tmp_path):
    b = _bytes(tmp_path, _tree())
    with pytest.raises(h5spec.SpecError):                                  # end-of-file address
        h5spec.verify(b + b"\0")
    with pytest.raises(h5spec.SpecError):                                  # a symbol-table node signature
        h5spec.verify(_corrupt(b, b.index(b"SNOD"), b"SNOX"))
    with pytest.raises(h5spec.SpecError):                                  # reserved byte of the superblock
        h5spec.verify(_corrupt(b, 11, b"\x01"))
    i = b.index(b"v_003\0")
    with pytest.raises(h5spec.SpecError):                                  # a member name out of order
        h5spec.verify(_corrupt(b, i, b"v_9"))
    oh = struct.unpack_from("<Q", b, 64)[0]
    with pytest.raises(h5spec.SpecError):                                  # message count of the root object header
        h5spec.verify(_corrupt(b, oh + 2, struct.pack("<H", struct.unpack_from("<H", b, oh + 2)[0] + 1)))


def test_mesh_and_label_files(tmp_path):
    """N3: data_io.save_mesh_hdf5 / save_functions_hdf5 (reference data_io.py:663-713, 751-760) through the same verifier."""
    from glimslib_b200.utils import data_io as dio
    mesh = fenics.UnitSquareMesh(4, 3)
    labels = fenics.MeshFunction("size_t", mesh, 2)
    labels.array()[:] = np.arange(mesh.num_cells()) % 3
    p = str(tmp_path / "mesh.h5")
    dio.save_mesh_hdf5(mesh, p, subdomains=labels)
    data, groups = h5spec.verify(open(p, "rb").read())
    topo = [k for k in data if k.startswith("/mesh/") and data[k][0].ndim == 2 and data[k][0].dtype.kind in "iu"]
    geom = [k for k in data if k.startswith("/mesh/") and data[k][0].dtype.kind == "f"]
    assert topo and geom
    assert np.array_equal(np.sort(data[topo[0]][0], axis=1), np.sort(mesh.cells, axis=1))
    assert np.allclose(data[geom[0]][0], mesh.coords)
    vals = [v for k, (v, _) in data.items() if k.startswith("/subdomains/") and v.size == mesh.num_cells() and v.ndim == 1]
    assert any(np.array_equal(np.asarray(v, dtype=np.int64), labels.array()) for v in vals)
    V = fenics.FunctionSpace(mesh, "CG", 1)
    f = fenics.project(fenics.Expression("x[0]+2*x[1]", degree=1), V)
    dio.save_functions_hdf5({"f": f}, str(tmp_path / "f.h5"), time_step=0)
    data, _ = h5spec.verify(open(str(tmp_path / "f.h5"), "rb").read())
    assert np.array_equal(data["/f/vector_0"][0], f.vector().get_local())
