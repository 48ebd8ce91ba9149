"""DOLFIN-shaped host objects for the drop-in seam (what ``glimslib/fenics_local.py:3-10`` re-exports and the
reference's hot-path modules actually touch -- SURVEY.md section 8b).  These hold *problem description* on the
host (meshes, labels, nodal arrays); every number the solve produces comes from the CUDA engine.

Dof layouts: Lagrange-1 spaces number ``dof(v, k) = v * ncomp + k`` (vertex-blocked; for the mixed
[P1^d, P1] space this is the engine's layout); DG-1 ``(cell * (d+1) + a) * ncomp + k``; DG-0 ``cell * ncomp + k``.
"""
import logging

import numpy as np

from .. import mesh as _mesh
from . import cexpr

__version__ = "2018.1.0"          # what fenics_local.is_version() parses (fenics_local.py:12-25)
DOLFIN_EPS = 3.0e-16
CRITICAL, ERROR, WARNING, INFO, PROGRESS, TRACE, DBG = 50, 40, 30, 20, 16, 13, 10


class LogLevel:
    CRITICAL, ERROR, WARNING, INFO, PROGRESS, TRACE, DEBUG = 50, 40, 30, 20, 16, 13, 10


def set_log_level(level):
    logging.getLogger("glimslib_b200").setLevel(int(level))


# ------------------------------------------------------------------------------------------------ geometry
class Point:
    def __init__(self, *xyz):
        if len(xyz) == 1 and np.ndim(xyz[0]) == 1:
            xyz = tuple(xyz[0])
        self._x = np.zeros(3)
        self._x[:len(xyz)] = xyz
        self._n = len(xyz)

    def x(self): return self._x[0]
    def y(self): return self._x[1]
    def z(self): return self._x[2]
    def array(self): return self._x.copy()
    def __getitem__(self, i): return self._x[i]


class Mesh(_mesh.SimplexMesh):
    """``fenics.Mesh()``: empty shell to be filled (e.g. by HDF5File.read) or built from arrays."""

    def __init__(self, coords=None, cells=None):
        if coords is None:
            coords, cells = np.zeros((0, 2)), np.zeros((0, 3), dtype=np.int32)
        super().__init__(coords, cells)

    def _assign(self, other):
        self.__dict__.update(_mesh.SimplexMesh(other.coords, other.cells).__dict__)


class MeshEditor:
    """``fenics.MeshEditor``: build a mesh vertex by vertex / cell by cell (data_io.py:458-468, 495-505)."""

    def open(self, mesh, cell_type, tdim, gdim, degree=1):
        if isinstance(cell_type, int):          # 2017.2 also accepts (mesh, tdim, gdim)
            cell_type, tdim, gdim = ("triangle" if cell_type == 2 else "tetrahedron"), cell_type, tdim
        self._mesh, self._tdim, self._gdim = mesh, int(tdim), int(gdim)
        if cell_type not in ("triangle", "tetrahedron"):
            raise ValueError("MeshEditor: only simplex cells are supported")

    def init_vertices(self, n):
        self._pts = np.zeros((int(n), self._gdim))

    init_vertices_global = init_vertices

    def init_cells(self, n):
        self._cells = np.zeros((int(n), self._tdim + 1), dtype=np.int32)

    init_cells_global = init_cells

    def add_vertex(self, i, x):
        self._pts[int(i)] = np.asarray(x.array() if hasattr(x, "array") else x, dtype=np.float64)[:self._gdim]

    def add_cell(self, i, verts):
        self._cells[int(i)] = np.asarray(verts, dtype=np.int64)

    def close(self, order=True):
        self._mesh._assign(_mesh.SimplexMesh(self._pts, self._cells))


def _as_mesh(sm):
    m = Mesh(sm.coords, sm.cells)
    if getattr(sm, "_boundary_vertices", None) is not None:
        m._boundary_vertices = sm._boundary_vertices
    return m


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    return _as_mesh(_mesh.rectangle_mesh((p0[0], p0[1]), (p1[0], p1[1]), nx, ny, diagonal))


def BoxMesh(p0, p1, nx, ny, nz):
    return _as_mesh(_mesh.box_mesh((p0[0], p0[1], p0[2]), (p1[0], p1[1], p1[2]), nx, ny, nz))


def UnitSquareMesh(nx, ny, diagonal="right"):
    return RectangleMesh(Point(0, 0), Point(1, 1), nx, ny, diagonal)


def UnitCubeMesh(nx, ny, nz):
    return BoxMesh(Point(0, 0, 0), Point(1, 1, 1), nx, ny, nz)


class _Entity:
    def __init__(self, mesh, dim, index):
        self._mesh, self._dim, self._index = mesh, dim, index

    def index(self): return self._index

    def midpoint(self):
        if self._dim == self._mesh.dim:
            return Point(self._mesh.coords[self._mesh.cells[self._index]].mean(axis=0))
        f = self._mesh.facets()[0][self._index]
        return Point(self._mesh.coords[f].mean(axis=0))

    def exterior(self):
        return self._mesh.facets()[2][self._index, 1] < 0


def cells(obj):
    if isinstance(obj, _Entity):       # cells(facet)
        fc = obj._mesh.facets()[2][obj._index]
        return (_Entity(obj._mesh, obj._mesh.dim, int(c)) for c in fc if c >= 0)
    return (_Entity(obj, obj.dim, i) for i in range(obj.num_cells()))


def facets(mesh):
    return (_Entity(mesh, mesh.dim - 1, i) for i in range(len(mesh.facets()[0])))


class MeshFunction:
    """``fenics.MeshFunction('size_t', mesh, dim)`` over cells (dim = d) or facets (dim = d-1)."""

    def __init__(self, dtype, mesh, dim, value=0):
        self._mesh, self._dim = mesh, dim
        n = mesh.num_cells() if dim == mesh.dim else len(mesh.facets()[0])
        self._a = np.full(n, value, dtype=np.int64 if dtype in ("size_t", "int", "uint") else np.float64)

    def mesh(self): return self._mesh
    def dim(self): return self._dim
    def array(self): return self._a
    def set_all(self, v): self._a[:] = v
    def size(self): return len(self._a)

    def _i(self, k):
        return k._index if isinstance(k, _Entity) else k

    def __getitem__(self, k): return self._a[self._i(k)]
    def __setitem__(self, k, v): self._a[self._i(k)] = v
    def rename(self, *a): pass


class SubDomain:
    """Subclass and override ``inside(x, on_boundary)`` (reference: helper_classes.py:61-63)."""

    def inside(self, x, on_boundary):
        return False

    def _inside_many(self, X, on_boundary):
        n = X.shape[0]
        try:   # vectorised attempt: x[i] is an array over points
            r = self.inside(X.T, on_boundary)
            r = np.asarray(r)
            if r.shape == (n,) and r.dtype == bool:
                return r
            if r.shape == () and n > 0 and isinstance(on_boundary, (bool, np.bool_)):
                pass
        except Exception:
            pass
        ob = np.broadcast_to(np.asarray(on_boundary), (n,))
        return np.array([bool(self.inside(X[i], bool(ob[i]))) for i in range(n)], dtype=bool)

    def mark(self, meshfunction, value):
        """DOLFIN semantics: an entity is marked when all its vertices and its midpoint are inside;
        ``on_boundary`` is true for entities on the exterior boundary."""
        mesh = meshfunction.mesh()
        if meshfunction.dim() == mesh.dim:
            ent = mesh.cells
            onb = np.zeros(len(ent), dtype=bool)
        else:
            f, _, fc = mesh.facets()
            ent, onb = f, fc[:, 1] < 0
        ok = np.ones(len(ent), dtype=bool)
        for k in range(ent.shape[1]):
            ok &= self._inside_many(mesh.coords[ent[:, k]], onb)
        ok &= self._inside_many(mesh.coords[ent].mean(axis=1), onb)
        meshfunction.array()[ok] = value


class _Coefficient:
    """Scalar coefficients multiply into integrands: ``Constant(2.0) * c * ds(3)`` (the boundary terms of the reference,
    helper_classes.py:896-908, and its unit tests build such expressions and hand them to ``assemble``)."""
    __array_priority__ = 1000          # numpy scalars defer to these operators

    def __mul__(self, other): return Integrand([self]) * other
    def __rmul__(self, other): return Integrand(_factors(other) + [self])


def _factors(obj):
    if isinstance(obj, Integrand):
        return list(obj.factors)
    if isinstance(obj, (int, float, np.integer, np.floating)):
        return [Constant(float(obj))]
    if isinstance(obj, (_Coefficient, ComponentView)):
        return [obj]
    raise TypeError("cannot multiply an integrand by %r" % type(obj).__name__)


class Integrand:
    """Product of scalar coefficients, each affine per cell (Constant, P1 function or sub-function, degree-1 expression)."""

    def __init__(self, factors): self.factors = list(factors)

    def __mul__(self, other):
        if isinstance(other, Measure):
            return Form([(self, other)])
        return Integrand(self.factors + _factors(other))

    def __rmul__(self, other): return Integrand(_factors(other) + self.factors)


class Form:
    """Sum of integrand x measure terms; ``assemble(form)`` integrates it exactly (scalar functionals only -- the residual
    and Jacobian of the hot path are not built from forms, they are the kernels of DESIGN.md section 5)."""

    def __init__(self, terms): self.terms = list(terms)

    def __add__(self, other):
        if isinstance(other, (int, float)) and other == 0:
            return self
        return Form(self.terms + other.terms)

    __radd__ = __add__

    def __neg__(self): return Form([(Integrand([Constant(-1.0)] + i.factors), m) for i, m in self.terms])
    def __sub__(self, other): return self + (-other)


class Measure:
    def __init__(self, kind, subdomain_data=None, subdomain_id=None):
        self.kind, self._data, self.subdomain_id = kind, subdomain_data, subdomain_id

    def subdomain_data(self): return self._data

    def __call__(self, subdomain_id=None, subdomain_data=None):
        return Measure(self.kind, subdomain_data if subdomain_data is not None else self._data,
                       subdomain_id if subdomain_id is not None else self.subdomain_id)

    def __rmul__(self, other): return Integrand(_factors(other)) * self


dx, ds, dS = Measure("dx"), Measure("ds"), Measure("dS")


# ------------------------------------------------------------------------------------------------ coefficients
class Constant(_Coefficient):
    def __init__(self, value, **kw):
        self._v = np.atleast_1d(np.asarray(value, dtype=np.float64)).ravel()
        self._scalar = np.ndim(value) == 0

    def value_size(self): return len(self._v)
    def values(self): return self._v.copy()
    def __float__(self): return float(self._v[0])
    def assign(self, v): self._v[:] = np.atleast_1d(np.asarray(float(v) if np.ndim(v) == 0 else v)).ravel()
    def eval_points(self, X): return np.broadcast_to(self._v, (X.shape[0], len(self._v))).copy()
    def ufl_shape(self): return () if self._scalar else (len(self._v),)


class Expression(_Coefficient):
    """``fenics.Expression(cpp_code | (cpp, ...), degree=..., **params)`` or a Python subclass overriding
    ``eval(values, x)`` (and optionally ``value_shape``)."""

    def __init__(self, code=None, degree=1, element=None, **params):
        object.__setattr__(self, "_params", dict(params))
        self.degree = degree
        if code is None:
            self._trees = None
        else:
            codes = (code,) if isinstance(code, str) else tuple(code)
            self._trees = [cexpr.parse(c) for c in codes]
            self._vector = not isinstance(code, str)

    def __setattr__(self, k, v):
        p = self.__dict__.get("_params")
        if p is not None and k in p:
            p[k] = v
        object.__setattr__(self, k, v)

    def __getattr__(self, k):
        p = self.__dict__.get("_params")
        if p is not None and k in p:
            return p[k]
        raise AttributeError(k)

    def value_shape(self):
        if getattr(self, "_trees", None) is None:
            return ()
        return (len(self._trees),) if self._vector else ()

    def value_size(self):
        s = self.value_shape()
        return int(np.prod(s)) if s else 1

    def eval_points(self, X):
        n = X.shape[0]
        if getattr(self, "_trees", None) is not None:
            return np.stack([cexpr.evaluate(t, X, self._params) for t in self._trees], axis=1)
        out = np.zeros((n, self.value_size()))
        for i in range(n):
            self.eval(out[i], X[i])
        return out

    def __call__(self, *x):
        p = np.asarray(x[0].array()[:len(x)] if isinstance(x[0], Point) else (x[0] if len(x) == 1 else x), float)
        v = self.eval_points(np.atleast_2d(p))[0]
        return float(v[0]) if len(v) == 1 else v


UserExpression = Expression


# ------------------------------------------------------------------------------------------------ elements / spaces
class FiniteElement:
    def __init__(self, family, cell=None, degree=1):
        self.family = {"Lagrange": "CG", "P": "CG", "CG": "CG", "DG": "DG", "Discontinuous Lagrange": "DG"}[family]
        self.cell, self.degree, self.ncomp = cell, degree, 1

    def value_size(self): return self.ncomp
    def sub_elements(self): return []
    def num_sub_elements(self): return 0


class VectorElement(FiniteElement):
    def __init__(self, family, cell=None, degree=1, dim=None):
        super().__init__(family, cell, degree)
        self.ncomp = dim if dim is not None else {"triangle": 2, "tetrahedron": 3, "interval": 1}[cell]

    def sub_elements(self):
        return [FiniteElement(self.family, self.cell, self.degree) for _ in range(self.ncomp)]

    def num_sub_elements(self): return self.ncomp


class TensorElement(FiniteElement):
    def __init__(self, family, cell=None, degree=1, shape=None):
        super().__init__(family, cell, degree)
        d = {"triangle": 2, "tetrahedron": 3, "interval": 1}[cell]
        self.shape = shape or (d, d)
        self.ncomp = int(np.prod(self.shape))

    def sub_elements(self):
        return [FiniteElement(self.family, self.cell, self.degree) for _ in range(self.ncomp)]

    def num_sub_elements(self): return self.ncomp


class MixedElement:
    def __init__(self, *elements):
        self._subs = list(elements[0]) if len(elements) == 1 and isinstance(elements[0], (list, tuple)) else list(elements)
        fam = {e.family for e in self._subs}
        deg = {e.degree for e in self._subs}
        if len(fam) != 1 or len(deg) != 1:
            raise NotImplementedError("mixed elements of different family/degree are outside the P1 hot path")
        self.family, self.degree = fam.pop(), deg.pop()
        self.ncomp = sum(e.ncomp for e in self._subs)

    def value_size(self): return self.ncomp
    def sub_elements(self): return list(self._subs)
    def num_sub_elements(self): return len(self._subs)


class FunctionSpace:
    def __init__(self, mesh, element, degree=None, _parent=None, _comps=None):
        if isinstance(element, str):
            element = FiniteElement(element, mesh.ufl_cell(), 1 if degree is None else degree)
        if element.degree not in (0, 1) or (element.family == "CG" and element.degree != 1):
            raise NotImplementedError("only P1 / DG0 / DG1 spaces are on the hot path (north star: P1)")
        self._mesh, self._element = mesh, element
        self.ncomp = element.ncomp
        self._parent, self._comps = _parent, _comps

    # DOLFIN API
    def mesh(self): return self._mesh
    def ufl_element(self): return self._element
    def element(self): return self._element
    def num_sub_spaces(self): return self._element.num_sub_elements()

    def n_nodes(self):
        m, e = self._mesh, self._element
        if e.family == "CG":
            return m.num_vertices()
        return m.num_cells() * (m.dim + 1 if e.degree == 1 else 1)

    def dim(self): return self.n_nodes() * self.ncomp

    def node_coordinates(self):
        m, e = self._mesh, self._element
        if e.family == "CG":
            return m.coords
        if e.degree == 1:
            return m.coords[m.cells].reshape(-1, m.dim)
        return m.cell_midpoints()

    def tabulate_dof_coordinates(self):
        return np.repeat(self.node_coordinates(), self.ncomp, axis=0)

    def sub(self, i):
        subs = self._element.sub_elements()
        off = sum(e.ncomp for e in subs[:i])
        base = self._comps[0] if self._comps is not None else 0
        root = self._parent if self._parent is not None else self
        return FunctionSpace(self._mesh, subs[i], _parent=root, _comps=(base + off, base + off + subs[i].ncomp))

    def collapse(self):
        return FunctionSpace(self._mesh, self._element)

    def dofmap(self): return self

    def dofs(self):
        if self._parent is None:
            return np.arange(self.dim())
        n = self._parent.n_nodes()
        return (np.arange(n)[:, None] * self._parent.ncomp + np.arange(*self._comps)[None, :]).ravel()


def TensorFunctionSpace(mesh, family, degree, shape=None):
    return FunctionSpace(mesh, TensorElement(family, mesh.ufl_cell(), degree, shape))


def VectorFunctionSpace(mesh, family, degree, dim=None):
    return FunctionSpace(mesh, VectorElement(family, mesh.ufl_cell(), degree, dim))


class Vector:
    """numpy-backed stand-in for ``GenericVector`` (``f.vector()``)."""

    def __init__(self, owner): self._o = owner
    def get_local(self): return self._o._x.copy()
    def array(self): return self._o._x.copy()
    def set_local(self, a): self._o._x[:] = np.asarray(a, dtype=np.float64).ravel(); self._o._touch()
    def apply(self, mode=""): pass
    def size(self): return self._o._x.size
    def __len__(self): return self._o._x.size
    def __getitem__(self, k): return self._o._x[k]

    def __setitem__(self, k, v):
        self._o._x[k] = v._o._x if isinstance(v, Vector) else v
        self._o._touch()

    def norm(self, kind="l2"):
        return float(np.linalg.norm(self._o._x, {"l2": 2, "l1": 1, "linf": np.inf}[kind]))

    def max(self): return float(self._o._x.max())
    def min(self): return float(self._o._x.min())
    def axpy(self, a, other): self._o._x += a * other._o._x; self._o._touch()
    def __array__(self, dtype=None, copy=None): return self._o._x if dtype is None else self._o._x.astype(dtype)


class ComponentView:
    """Result of ``split(u)[i]`` / ``u.sub(i)``: components [c0, c1) of a function on a blocked space."""

    def __init__(self, function, c0, c1):
        self.function, self.c0, self.c1 = function, c0, c1

    def values(self):
        f = self.function
        return f._x.reshape(-1, f.function_space().ncomp)[:, self.c0:self.c1]

    def function_space(self):
        W = self.function.function_space()
        sizes = np.cumsum([0] + [e.ncomp for e in W._element.sub_elements()])
        return W.sub(int(np.nonzero(sizes == self.c0)[0][0]))

    def geometric_dimension(self): return self.function.geometric_dimension()
    def __len__(self): return self.c1 - self.c0
    def __mul__(self, other): return Integrand([self]) * other
    def __rmul__(self, other): return Integrand(_factors(other) + [self])


class Function(_Coefficient):
    def __init__(self, V, name=None, **kw):
        if isinstance(V, Function):
            self._V, self._x = V._V, V._x.copy()
        else:
            self._V, self._x = V, np.zeros(V.dim())
        self._name, self.version = name or "f", 0

    def _touch(self): self.version += 1
    def function_space(self): return self._V
    def vector(self): return Vector(self)
    def geometric_dimension(self): return self._V.mesh().dim
    def name(self): return self._name
    def rename(self, name, label=""): self._name = name
    def value_size(self): return self._V.ncomp

    def copy(self, deepcopy=True):
        g = Function(self._V, name=self._name)
        g._x[:] = self._x
        return g

    def assign(self, other):
        if isinstance(other, Function):
            self._x[:] = other._x
        elif isinstance(other, (Constant, Expression)):
            self._x[:] = other.eval_points(self._V.node_coordinates()).ravel()
        else:
            raise TypeError("cannot assign %r" % type(other))
        self._touch()

    def interpolate(self, expr): self.assign(expr)

    def _combine(self, other, sign):
        if not isinstance(other, Function) or other._x.shape != self._x.shape:
            raise TypeError("only functions of the same space can be added / subtracted")
        out = self.copy(deepcopy=True)
        out._x[:] = self._x + sign * other._x
        out._touch()
        return out

    def __sub__(self, other): return self._combine(other, -1.0)      # a Function (the reference gets a UFL sum it then stores)
    def __add__(self, other): return self._combine(other, 1.0)

    def sub(self, i, deepcopy=False):
        S = self._V.sub(i)
        if deepcopy:
            g = Function(S.collapse())
            g._x[:] = ComponentView(self, *S._comps).values().ravel()
            return g
        return ComponentView(self, *S._comps)

    def split(self, deepcopy=False):
        return tuple(self.sub(i, deepcopy) for i in range(self._V.num_sub_spaces()))

    def compute_vertex_values(self, mesh=None):
        """DOLFIN order: component-major [ncomp][n_vertices] (CG1 only)."""
        assert self._V._element.family == "CG"
        return self._x.reshape(-1, self._V.ncomp).T.ravel().copy()

    def node_values(self):
        return self._x.reshape(-1, self._V.ncomp)

    def cell_midpoint_values(self):
        """Value of the function at every cell midpoint (what ``f(cell.midpoint())`` returns)."""
        m, e = self._V.mesh(), self._V._element
        v = self.node_values()
        if e.family == "CG":
            return v[m.cells].mean(axis=1)
        if e.degree == 1:
            return v.reshape(m.num_cells(), m.dim + 1, -1).mean(axis=1)
        return v

    def __call__(self, *x):
        p = np.asarray(x[0].array()[:self._V.mesh().dim] if isinstance(x[0], Point) else (x[0] if len(x) == 1 else x), float)
        m = self._V.mesh()
        X = m.coords[m.cells]
        T = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))
        lam = np.linalg.solve(T, np.broadcast_to(p - X[:, 0], (len(X), m.dim))[..., None])[..., 0]
        bary = np.concatenate([1 - lam.sum(axis=1, keepdims=True), lam], axis=1)
        c = int(np.argmax(bary.min(axis=1)))
        if bary[c].min() < -1e-10:
            raise RuntimeError("point outside the mesh")
        e = self._V._element
        v = self.node_values()
        if e.family == "CG":
            val = bary[c] @ v[m.cells[c]]
        elif e.degree == 1:
            val = bary[c] @ v.reshape(m.num_cells(), m.dim + 1, -1)[c]
        else:
            val = v[c]
        return float(val[0]) if len(val) == 1 else val


def split(f):
    return f.split()


def TrialFunction(V): return ("trial", V)
def TestFunction(V): return ("test", V)
def TestFunctions(V): return tuple(("test", V.sub(i)) for i in range(V.num_sub_spaces()))
def TrialFunctions(V): return tuple(("trial", V.sub(i)) for i in range(V.num_sub_spaces()))


def _values_on(V, src):
    """Nodal values (n_nodes, ncomp) of ``src`` on space V -- exact for degree<=1 data (the P1 nodal
    interpolant is its own L2 projection; helper_classes.py:983-986 / SURVEY.md a11)."""
    if isinstance(src, (int, float)):
        src = Constant(float(src))
    if isinstance(src, (tuple, list)):
        return np.concatenate([_values_on(FunctionSpace(V.mesh(), FiniteElement(V._element.family, None, V._element.degree)), s)
                               for s in src], axis=1)
    if isinstance(src, (Constant, Expression)):
        v = src.eval_points(V.node_coordinates())
    elif isinstance(src, (Function, ComponentView)):
        f = src if isinstance(src, Function) else src.function
        vals = f.node_values() if isinstance(src, Function) else src.values()
        Vs, m = f.function_space(), V.mesh()
        if Vs._element.family == V._element.family and Vs._element.degree == V._element.degree:
            v = vals
        elif Vs._element.family == "CG" and V._element.family == "DG" and V._element.degree == 1:
            v = vals[m.cells].reshape(-1, vals.shape[1])
        elif Vs._element.family == "CG" and V._element.degree == 0:
            v = vals[m.cells].mean(axis=1)
        else:
            raise NotImplementedError("projection %s%d -> %s%d is outside the hot path"
                                      % (Vs._element.family, Vs._element.degree, V._element.family, V._element.degree))
    else:
        raise TypeError("cannot project %r" % type(src))
    if v.shape[1] != V.ncomp:
        raise ValueError("value size %d does not match the space (%d)" % (v.shape[1], V.ncomp))
    return v


# ------------------------------------------------------------------------------------------------ scalar functionals
def _scalar_vertex_values(f, mesh):
    """Values at the mesh vertices of a scalar coefficient that is affine per cell."""
    nv = mesh.num_vertices()
    if isinstance(f, Constant):
        if f.value_size() != 1:
            raise TypeError("assemble: only scalar integrands are supported (use inner/dot on the host side)")
        return np.full(nv, float(f))
    if isinstance(f, Function):
        V = f.function_space()
        if V.ncomp != 1 or V._element.family != "CG":
            raise TypeError("assemble: factors must be scalar P1 functions")
        return f.node_values()[:, 0]
    if isinstance(f, ComponentView):
        v = f.values()
        if v.shape[1] != 1:
            raise TypeError("assemble: factors must be scalar sub-functions")
        return v[:, 0]
    if isinstance(f, Expression):
        v = np.asarray(f.eval_points(mesh.coords)).reshape(nv, -1)
        if v.shape[1] != 1:
            raise TypeError("assemble: factors must be scalar expressions")
        return v[:, 0]
    raise TypeError("assemble: unsupported factor %r" % type(f).__name__)


def _integrate(integrand, measure):
    """Exact integral of a product of per-cell affine scalars over the cells (dx) or exterior facets (ds) the measure
    selects: the product is expanded in barycentric monomials and int lambda^alpha = |e| k! alpha! / (k + |alpha|)!."""
    from math import factorial
    data = measure.subdomain_data()
    mesh = data.mesh() if data is not None else None
    for f in integrand.factors:
        if mesh is not None:
            break
        if isinstance(f, Function):
            mesh = f.function_space().mesh()
        elif isinstance(f, ComponentView):
            mesh = f.function.function_space().mesh()
    if mesh is None:
        raise ValueError("assemble: the form does not name a mesh (no subdomain data, no function)")
    if measure.kind == "dx":
        ent = np.asarray(mesh.cells)
        sel = np.ones(len(ent), dtype=bool)
        if measure.subdomain_id is not None:
            sel = np.asarray(data.array()) == measure.subdomain_id
    elif measure.kind == "ds":
        ids, ent, _ = mesh.exterior_facets()
        sel = np.ones(len(ent), dtype=bool)
        if measure.subdomain_id is not None:
            sel = np.asarray(data.array())[ids] == measure.subdomain_id
    else:
        raise NotImplementedError("assemble: interior-facet measures are not part of the drop-in")
    ent = ent[sel]
    if len(ent) == 0:
        return 0.0
    X = np.asarray(mesh.coords)[ent]                                    # (n, k+1, dim)
    k = ent.shape[1] - 1
    E = X[:, 1:, :] - X[:, :1, :]
    vol = np.sqrt(np.abs(np.linalg.det(E @ E.transpose(0, 2, 1)))) / factorial(k) if k > 0 else np.ones(len(ent))
    poly = {(0,) * (k + 1): np.ones(len(ent))}
    for f in integrand.factors:
        V = _scalar_vertex_values(f, mesh)[ent]                         # (n, k+1)
        nxt = {}
        for alpha, coef in poly.items():
            for a in range(k + 1):
                b = alpha[:a] + (alpha[a] + 1,) + alpha[a + 1:]
                nxt[b] = nxt.get(b, 0.0) + coef * V[:, a]
        poly = nxt
    total = 0.0
    for alpha, coef in poly.items():
        w = factorial(k) * np.prod([factorial(a) for a in alpha]) / factorial(k + sum(alpha))
        total += w * float(np.dot(coef, vol))
    return total


def assemble(form):
    """``fenics.assemble`` for scalar functionals (sums of coefficient products times dx / ds measures)."""
    if isinstance(form, (int, float)):
        return float(form)
    if not isinstance(form, Form):
        raise TypeError("assemble: expected a Form (integrand * measure), got %r" % type(form).__name__)
    return float(sum(_integrate(i, m) for i, m in form.terms))


def project(v, V=None, **kw):
    if V is None:
        raise ValueError("project(v, V): V is required")
    tgt = V.collapse() if V._parent is not None else V
    f = Function(tgt)
    f._x[:] = _values_on(tgt, v).ravel()
    return f


def interpolate(v, V):
    return project(v, V)


class FunctionAssigner:
    """``FunctionAssigner(W.sub(i), V_i).assign(U.sub(i), f)`` (helper_classes.py:358-359) and the reverse."""

    def __init__(self, to_space, from_space):
        self._to, self._from = to_space, from_space

    def assign(self, to_f, from_f):
        if isinstance(to_f, ComponentView):
            src = from_f.node_values() if isinstance(from_f, Function) else from_f.values()
            to_f.values()[:] = src
            to_f.function._touch()
        else:
            src = from_f.values() if isinstance(from_f, ComponentView) else from_f.node_values()
            to_f.node_values()[:] = src
            to_f._touch()


def assign(to_f, from_f):
    FunctionAssigner(None, None).assign(to_f, from_f)


class DirichletBC:
    """Topological vertex-dof Dirichlet set on a (sub)space: ``DirichletBC(V, g, SubDomain)`` or
    ``DirichletBC(V, g, facet_function, id)`` (helper_classes.py:705,712,717)."""

    def __init__(self, V, value, where, marker=None, method="topological"):
        self._V, self._value = V, value
        mesh = V.mesh()
        if isinstance(where, SubDomain):
            mf = MeshFunction("size_t", mesh, mesh.dim - 1, 0)
            where.mark(mf, 1)
            sel = mf.array() == 1
        elif isinstance(where, MeshFunction):
            sel = where.array() == marker
        else:
            raise TypeError("DirichletBC: third argument must be a SubDomain or a facet MeshFunction")
        verts = np.unique(mesh.facets()[0][sel].ravel()) if sel.any() else np.zeros(0, dtype=np.int64)
        root = V._parent if V._parent is not None else V
        c0, c1 = V._comps if V._comps is not None else (0, V.ncomp)
        if root._element.family != "CG":
            raise NotImplementedError("Dirichlet conditions on DG spaces")
        self.vertices = verts
        self.dofs = (verts[:, None] * root.ncomp + np.arange(c0, c1)[None, :]).ravel().astype(np.int64)
        self._ncomp = c1 - c0
        self.update()

    def update(self):
        v = self._value if not isinstance(self._value, (int, float)) else Constant(float(self._value))
        vals = v.eval_points(self._V.mesh().coords[self.vertices])
        if vals.shape[1] != self._ncomp:
            raise ValueError("DirichletBC value size %d does not match the sub-space (%d)" % (vals.shape[1], self._ncomp))
        self.values = vals.ravel()

    def function_space(self): return self._V
    def get_boundary_values(self): return dict(zip(self.dofs.tolist(), self.values.tolist()))

    def apply(self, x):
        tgt = x._o._x if isinstance(x, Vector) else x
        tgt[self.dofs] = self.values


def _mass_apply(mesh, ncomp, e):
    """(M e) for P1 nodal differences e[n_vertices, ncomp]: exact P1 mass matrix, matrix-free."""
    X = mesh.coords[mesh.cells]
    d = mesh.dim
    T = X[:, 1:] - X[:, :1]
    vol = np.abs(np.linalg.det(T)) / {2: 2.0, 3: 6.0}[d]
    ec = e[mesh.cells]                                   # (nc, d+1, ncomp)
    loc = (ec + ec.sum(axis=1, keepdims=True)) * (vol / ((d + 1) * (d + 2)))[:, None, None]
    return float((ec * loc).sum())


def errornorm(u, uh, norm_type="L2", degree_rise=0, mesh=None):
    """L2 norm of ``u - uh`` for P1 data (helper_classes.py:2001-2013 uses it as the comparison metric)."""
    if norm_type.lower() != "l2":
        raise NotImplementedError("only the L2 errornorm is provided")
    a = u.node_values() if isinstance(u, Function) else u.values()
    b = uh.node_values() if isinstance(uh, Function) else uh.values()
    m = (u if isinstance(u, Function) else u.function).function_space().mesh()
    return float(np.sqrt(max(_mass_apply(m, a.shape[1], a - b), 0.0)))


def norm(f, norm_type="L2"):
    if isinstance(f, Vector):
        return f.norm("l2")
    v = f.node_values() if isinstance(f, Function) else f.values()
    m = (f if isinstance(f, Function) else f.function).function_space().mesh()
    return float(np.sqrt(max(_mass_apply(m, v.shape[1], v), 0.0)))


# ------------------------------------------------------------------------------------------------ UFL algebra
def _no_ufl(name):
    def f(*a, **k):
        raise NotImplementedError(
            "fenics.%s: UFL form algebra is not part of the B200 backend -- the coupled RD-mechanics weak form "
            "(simulation_tumor_growth.py:110-124) is evaluated in closed form by the CUDA element kernels and "
            "described to the solver by CoupledRDMechanicsForm" % name)
    f.__name__ = name
    return f


inner, grad, sym, tr, det, dot, div, sqrt, derivative, Identity, solve = (
    _no_ufl(n) for n in ("inner", "grad", "sym", "tr", "det", "dot", "div", "sqrt", "derivative", "Identity", "solve"))
