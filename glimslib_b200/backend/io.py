"""Result writers of the seam: ``File`` (VTK .pvd/.vtu), ``XDMFFile`` and ``HDF5File``.

No HDF5 library exists in this environment (no h5py / libhdf5), so the ``.h5`` files are produced and parsed by
:mod:`glimslib_b200.backend.minih5`, a from-scratch writer/reader of genuine HDF5 (version-0 superblock, old-style
groups, contiguous datasets, v1 attributes).  The layout follows DOLFIN's conventions: ``HDF5File.write(f, name, t)``
appends ``/<name>/vector_<k>`` with a ``timestamp`` attribute and keeps ``count`` on ``/<name>``
(helper_classes.py:1256-1308); the first write also stores ``cells``, ``cell_dofs`` and ``x_cell_dofs``;
``XDMFFile`` keeps light data in XML and heavy data in ``<name>.h5`` (helper_classes.py:1360-1375,1436-1437).
"""
import os

import numpy as np

from . import core, minih5


class _Container:
    """Path-addressed view of a minih5 tree: ``data[path]`` datasets, ``attrs[path][name]`` attributes."""

    def __init__(self, path, mode):
        self.path, self.mode = path, mode
        if mode in ("r", "a") and os.path.exists(path):
            self.root = minih5.read_file(path)
        elif mode == "r":
            raise IOError("cannot open %s" % path)
        else:
            self.root = minih5.Group()
        self.data, self.attrs = _DataView(self.root), _AttrView(self.root)

    def flush(self):
        if self.mode == "r":
            return
        os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
        minih5.write_file(self.path, self.root)


class _DataView:
    def __init__(self, root): self._r = root

    def __setitem__(self, path, arr):
        self._r.create_dataset(path, np.asarray(arr))

    def __getitem__(self, path):
        d = self._r.get(path)
        if not isinstance(d, minih5.Dataset):
            raise KeyError(path)
        return d.data

    def __contains__(self, path):
        return isinstance(self._r.get(path), minih5.Dataset)


class _AttrView:
    """``attrs.setdefault(path, {})[k] = v`` / ``attrs[path][k]`` / ``attrs.get(path, {})`` on groups and datasets."""

    def __init__(self, root): self._r = root

    def _node(self, path, create=False):
        n = self._r.get(path)
        if n is None and create:
            n = self._r.require_group(path)
        return n

    def setdefault(self, path, default=None):
        return self._node(path, create=True).attrs

    def __getitem__(self, path):
        n = self._node(path)
        if n is None:
            raise KeyError(path)
        return n.attrs

    def get(self, path, default=None):
        n = self._node(path)
        return default if n is None else n.attrs


class _Attributes:
    def __init__(self, store, path): self._s, self._p = store, path.strip("/")

    def __getitem__(self, k):
        v = self._s.attrs[self._p][k]
        return int(v) if isinstance(v, (np.integer,)) else v

    def __setitem__(self, k, v): self._s.attrs.setdefault(self._p, {})[k] = v
    def __contains__(self, k): return k in self._s.attrs.get(self._p, {})
    def to_dict(self): return dict(self._s.attrs.get(self._p, {}))


class HDF5File:
    """``HDF5File(comm, path, 'w'|'r'|'a')`` with DOLFIN's time-series convention: ``write(f, name, t)`` appends
    ``/<name>/vector_<count>`` (+ ``timestamp``) and bumps ``count`` on ``/<name>``."""

    def __init__(self, comm, path, mode):
        self._c = _Container(path, mode)

    def write(self, obj, name, timestamp=None):
        name = name.strip("/")
        c = self._c
        if isinstance(obj, core.Function):
            if timestamp is None:
                self._write_function_layout(obj, name)
                c.data[name + "/vector_0"] = obj._x.copy()
                c.attrs.setdefault(name, {})["count"] = np.uint64(1)
            else:
                k = int(c.attrs.setdefault(name, {}).get("count", 0))
                if k == 0:
                    self._write_function_layout(obj, name)
                c.data["%s/vector_%d" % (name, k)] = obj._x.copy()
                c.attrs.setdefault("%s/vector_%d" % (name, k), {})["timestamp"] = float(timestamp)
                c.attrs[name]["count"] = np.uint64(k + 1)
        elif isinstance(obj, core.MeshFunction):
            c.data[name + "/values"] = np.asarray(obj.array()).copy()
            c.attrs.setdefault(name, {})["dim"] = np.int64(obj.dim())
        elif hasattr(obj, "coords") and hasattr(obj, "cells"):
            c.data[name + "/coordinates"] = obj.coords.copy()
            c.data[name + "/topology"] = obj.cells.copy()
        else:
            raise TypeError("HDF5File.write: unsupported object %r" % type(obj))

    def _write_function_layout(self, f, name):
        """``cells`` / ``cell_dofs`` / ``x_cell_dofs`` next to the vectors, as DOLFIN's HDF5File::write(Function) does."""
        V = f.function_space()
        m = V.mesh()
        if V._element.family != "CG":
            return
        nc, nvc = m.cells.shape
        dofs = (m.cells[:, :, None].astype(np.int64) * V.ncomp + np.arange(V.ncomp)[None, None, :]).reshape(nc, -1)
        c = self._c
        c.data[name + "/cells"] = np.arange(nc, dtype=np.uint64)
        c.data[name + "/cell_dofs"] = dofs.ravel()
        c.data[name + "/x_cell_dofs"] = (np.arange(nc + 1, dtype=np.int64) * dofs.shape[1])

    def read(self, obj, name, use_partition_from_file=False):
        name = name.strip("/")
        c = self._c
        if isinstance(obj, core.Function):
            key = name if name in c.data else name + "/vector_0"
            obj._x[:] = c.data[key]
            obj._touch()
        elif isinstance(obj, core.MeshFunction):
            obj.array()[:] = c.data[name + "/values"]
        elif isinstance(obj, core.Mesh):
            obj._assign(core.Mesh(c.data[name + "/coordinates"], c.data[name + "/topology"]))
        else:
            raise TypeError("HDF5File.read: unsupported object %r" % type(obj))

    def attributes(self, name): return _Attributes(self._c, name)
    def has_dataset(self, name): return self._c.root.get(name.strip("/")) is not None
    def flush(self): self._c.flush()
    def close(self): self._c.flush()


class XDMFFile:
    """``solution.xdmf`` (XML light data) + ``solution.h5`` heavy data (HDF5 via minih5).
    ``write(mesh)`` and ``write_checkpoint(function, name, t)`` as used at helper_classes.py:1360-1375,1436-1437."""

    class Encoding:
        HDF5, ASCII = 0, 1

    def __init__(self, comm, path):
        self.path = path
        self.h5path = os.path.splitext(path)[0] + ".h5"
        self._h5 = _Container(self.h5path, "w")
        self._mesh, self._steps = None, []
        self.parameters = {}

    def write(self, mesh, *a, **k):
        self._mesh = mesh
        self._h5.data["Mesh/mesh/geometry"] = mesh.coords.copy()
        self._h5.data["Mesh/mesh/topology"] = mesh.cells.copy()

    def write_checkpoint(self, function, name, time_step=0.0, encoding=None, append=True):
        V = function.function_space()
        if self._mesh is None:
            self.write(V.mesh())
        k = sum(1 for s in self._steps if s[0] == name)
        key = "%s/%s_%d/vector" % (name, name, k)
        self._h5.data[key] = function._x.copy()
        self._steps.append((name, float(time_step), key, V.ncomp, function._x.size // V.ncomp))

    def close(self):
        self._h5.flush()
        m = self._mesh
        if m is None:
            return
        ttype = "Triangle" if m.dim == 2 else "Tetrahedron"
        h5 = os.path.basename(self.h5path)
        lines = ['<?xml version="1.0"?>', '<Xdmf Version="3.0"><Domain>']
        names = sorted({s[0] for s in self._steps})
        for nm in names:
            lines.append('<Grid Name="%s" GridType="Collection" CollectionType="Temporal">' % nm)
            for (n2, t, key, ncomp, nn) in self._steps:
                if n2 != nm:
                    continue
                lines += ['<Grid Name="%s" GridType="Uniform">' % nm, '<Time Value="%.16g"/>' % t,
                          '<Topology TopologyType="%s" NumberOfElements="%d"><DataItem Dimensions="%d %d" Format="HDF">%s:/Mesh/mesh/topology</DataItem></Topology>'
                          % (ttype, m.num_cells(), m.num_cells(), m.dim + 1, h5),
                          '<Geometry GeometryType="%s"><DataItem Dimensions="%d %d" Format="HDF">%s:/Mesh/mesh/geometry</DataItem></Geometry>'
                          % ("XY" if m.dim == 2 else "XYZ", m.num_vertices(), m.dim, h5),
                          '<Attribute Name="%s" AttributeType="%s" Center="Node"><DataItem Dimensions="%d %d" Format="HDF">%s:/%s</DataItem></Attribute>'
                          % (nm, "Scalar" if ncomp == 1 else "Vector", nn, ncomp, h5, key), '</Grid>']
            lines.append('</Grid>')
        lines.append('</Domain></Xdmf>')
        os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
        with open(self.path, "w") as f:
            f.write("\n".join(lines) + "\n")


class File:
    """``File('x.pvd') << (function, t)``: ASCII VTU per record + a .pvd index (helper_classes.py:1376-1380)."""

    def __init__(self, path):
        self.path = path
        self._records = []

    def __lshift__(self, item):
        obj, t = item if isinstance(item, tuple) else (item, 0.0)
        base = os.path.splitext(self.path)[0]
        vtu = "%s%06d.vtu" % (base, len(self._records))
        _write_vtu(vtu, obj)
        self._records.append((float(t), os.path.basename(vtu)))
        with open(self.path, "w") as f:
            f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1"><Collection>\n')
            for tt, name in self._records:
                f.write('<DataSet timestep="%.16g" part="0" file="%s"/>\n' % (tt, name))
            f.write('</Collection></VTKFile>\n')
        return self


def _write_vtu(path, obj):
    if isinstance(obj, core.MeshFunction):
        mesh, cell_data, point_data, name = obj.mesh(), np.asarray(obj.array(), float), None, "f"
    else:
        mesh = obj.function_space().mesh()
        name, cell_data = obj.name(), None
        point_data = obj.node_values()
        if obj.function_space()._element.family != "CG":
            point_data = None
    d, nc, nv = mesh.dim, mesh.num_cells(), mesh.num_vertices()
    pts = np.zeros((nv, 3))
    pts[:, :d] = mesh.coords
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1"><UnstructuredGrid>\n')
        f.write('<Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (nv, nc))
        f.write('<Points><DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
        np.savetxt(f, pts, fmt="%.16g")
        f.write('</DataArray></Points>\n<Cells><DataArray type="Int32" Name="connectivity" format="ascii">\n')
        np.savetxt(f, mesh.cells, fmt="%d")
        f.write('</DataArray><DataArray type="Int32" Name="offsets" format="ascii">\n')
        np.savetxt(f, (np.arange(nc) + 1) * (d + 1), fmt="%d")
        f.write('</DataArray><DataArray type="UInt8" Name="types" format="ascii">\n')
        np.savetxt(f, np.full(nc, 5 if d == 2 else 10), fmt="%d")
        f.write('</DataArray></Cells>\n')
        if point_data is not None:
            ncomp = point_data.shape[1]
            out = point_data
            if ncomp == 2:
                out = np.concatenate([point_data, np.zeros((nv, 1))], axis=1)
            f.write('<PointData><DataArray type="Float64" Name="%s" NumberOfComponents="%d" format="ascii">\n' % (name, out.shape[1]))
            np.savetxt(f, out, fmt="%.16g")
            f.write('</DataArray></PointData>\n')
        if cell_data is not None and len(cell_data) == nc:
            f.write('<CellData><DataArray type="Float64" Name="%s" format="ascii">\n' % name)
            np.savetxt(f, cell_data, fmt="%.16g")
            f.write('</DataArray></CellData>\n')
        f.write('</Piece></UnstructuredGrid></VTKFile>\n')
