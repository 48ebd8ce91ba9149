"""Host-side simplex meshes for the drop-in backend (``fenics.RectangleMesh`` / ``BoxMesh`` / ``Mesh``).

Numbering follows DOLFIN 2017.2 [MEM, SURVEY.md 8c.3]: vertices x-fastest, two 'right'-diagonal
triangles per quad, six tetrahedra per hexahedron around the (v0, v7) diagonal.  Reference call sites:
``test_cases/test_simulation_tumor_growth/test_case_simulation_tumor_growth_2D_subdomains.py:34-35``.
"""
import numpy as np


class _Dim(int):
    """Geometric dimension that is both the int the backend computes with (``mesh.dim``, ``mesh.geometry().dim - 1``)
    and DOLFIN's method (``mesh.geometry().dim()``, used 16x in the reference, e.g. simulation_base.py:98)."""

    def __call__(self):
        return int(self)


class SimplexMesh:
    """Vertices + cells of a 2D/3D simplex mesh with lazily built facet topology."""

    def __init__(self, coords, cells):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.dim = _Dim(self.coords.shape[1])
        assert self.cells.shape[1] == self.dim + 1
        self._facets = None

    # -- DOLFIN-like accessors used by the reference (simulation_base.py:98, helper_classes.py:247-250)
    def geometry(self):
        return self

    def geometric_dimension(self):
        return self.dim

    def num_vertices(self):
        return len(self.coords)

    def num_cells(self):
        return len(self.cells)

    def coordinates(self):
        return self.coords

    def ufl_cell(self):
        return "triangle" if self.dim == 2 else "tetrahedron"

    def mpi_comm(self):
        return None

    def cell_midpoints(self):
        return self.coords[self.cells].mean(axis=1)

    # -- topology ---------------------------------------------------------------------------------
    def facets(self):
        """(facets[nf, d] sorted vertex ids, cell_facet[nc, d+1], facet_cells[nf, 2] with -1 padding)."""
        if self._facets is None:
            nc, nv = self.cells.shape
            parts = [np.delete(self.cells, k, axis=1) for k in range(nv)]
            f = np.sort(np.concatenate(parts, axis=0), axis=1).astype(np.int64)
            # 1-D integer keys (np.unique(axis=0) is ~20x slower at 10M tets)
            N = len(self.coords)
            key = f[:, 0] * N + f[:, 1]
            if self.dim == 3:
                _, id1 = np.unique(key, return_inverse=True)
                key = id1.reshape(-1) * N + f[:, 2]
            _, first_idx, inv = np.unique(key, return_index=True, return_inverse=True)
            inv = inv.reshape(-1)
            uniq = f[first_idx].astype(np.int32)
            cell_facet = inv.reshape(nv, nc).T
            fc = np.full((len(uniq), 2), -1, dtype=np.int64)
            cell_ids = np.tile(np.arange(nc), nv)
            order = np.argsort(inv, kind="stable")
            si, sc = inv[order], cell_ids[order]
            first = np.r_[True, si[1:] != si[:-1]]
            fc[si[first], 0] = sc[first]
            fc[si[~first], 1] = sc[~first]
            self._facets = (uniq, cell_facet, fc)
        return self._facets

    def exterior_facets(self):
        f, _, fc = self.facets()
        ext = np.nonzero(fc[:, 1] < 0)[0]
        return ext, f[ext], fc[ext, 0]

    def boundary_vertices(self):
        if getattr(self, "_boundary_vertices", None) is not None:
            return self._boundary_vertices
        _, f, _ = self.exterior_facets()
        m = np.zeros(len(self.coords), dtype=bool)
        m[f.ravel()] = True
        return np.nonzero(m)[0]


def rectangle_mesh(p0, p1, nx, ny, diagonal="right"):
    x = np.linspace(p0[0], p1[0], nx + 1)
    y = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(x, y, indexing="xy")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1, v2 = v0 + 1, v0 + nx + 1
    v3 = v2 + 1
    cells = np.empty((2 * nx * ny, 3), dtype=np.int32)
    if diagonal == "right":
        cells[0::2] = np.stack([v0, v1, v3], axis=1)
        cells[1::2] = np.stack([v0, v2, v3], axis=1)
    elif diagonal == "left":
        cells[0::2] = np.stack([v0, v1, v2], axis=1)
        cells[1::2] = np.stack([v1, v2, v3], axis=1)
    else:
        raise ValueError("diagonal must be 'right' or 'left'")
    return SimplexMesh(coords, cells)


def box_cells(nx, ny, nz, mask=None):
    """Six Kuhn tetrahedra per hexahedron; ``mask[nz, ny, nx]`` keeps a subset of hexahedra."""
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    v0 = (iz * sy + iy * sx + ix).ravel()
    if mask is not None:
        v0 = v0[np.asarray(mask).ravel()]
    v1, v2, v4 = v0 + 1, v0 + sx, v0 + sy
    v3, v5, v6 = v1 + sx, v1 + sy, v2 + sy
    v7 = v3 + sy
    tets = [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
            (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
    cells = np.empty((6 * len(v0), 4), dtype=np.int32)
    for k, t in enumerate(tets):
        cells[k::6] = np.stack(t, axis=1)
    return cells


def box_mesh(p0, p1, nx, ny, nz):
    x = np.linspace(p0[0], p1[0], nx + 1)
    y = np.linspace(p0[1], p1[1], ny + 1)
    z = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    return SimplexMesh(coords, box_cells(nx, ny, nz))


def voxel_ellipsoid_mesh(n, semi_axes):
    """Brain-like synthetic domain: the hexahedra of an n^3 voxel grid over the ellipsoid's bounding
    box whose centre lies inside the ellipsoid, each split into six tetrahedra -- the staircase
    boundary that segmentation-derived meshes have.  Unused vertices are dropped and the rest
    renumbered in grid order.  Returns (mesh, r) with r the normalised radius of every cell centroid."""
    a = np.asarray(semi_axes, dtype=np.float64)
    g = (np.arange(n) + 0.5) / n * 2.0 - 1.0
    Zc, Yc, Xc = np.meshgrid(g, g, g, indexing="ij")
    mask = (Xc * Xc + Yc * Yc + Zc * Zc) <= 1.0
    cells = box_cells(n, n, n, mask)
    used = np.zeros((n + 1) ** 3, dtype=bool)
    used[cells.ravel()] = True
    new_id = np.cumsum(used, dtype=np.int64) - 1
    cells = new_id[cells].astype(np.int32)
    lin = np.linspace(-1.0, 1.0, n + 1)
    Z, Y, X = np.meshgrid(lin, lin, lin, indexing="ij")
    unit = np.stack([X.ravel()[used], Y.ravel()[used], Z.ravel()[used]], axis=1)
    mesh = SimplexMesh(unit * a[None, :], cells)
    r = np.sqrt((unit[cells].mean(axis=1) ** 2).sum(axis=1))
    # exterior-boundary vertices, structurally: a used vertex with at least one adjacent voxel outside
    pad = np.zeros((n + 2, n + 2, n + 2), dtype=np.int8)
    pad[1:-1, 1:-1, 1:-1] = mask
    cnt = np.zeros((n + 1, n + 1, n + 1), dtype=np.int8)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                cnt += pad[dz:dz + n + 1, dy:dy + n + 1, dx:dx + n + 1]
    on_bnd = (cnt.ravel() > 0) & (cnt.ravel() < 8)
    mesh._boundary_vertices = new_id[np.nonzero(on_bnd)[0]]
    return mesh, r


# ---- locality renumbering (the library's internal reorder; exposed through the dof-permutation API) ----------------------
def _spread_bits(x, dim):
    x = x.astype(np.uint64)
    if dim == 3:        # 21 bits -> every third bit
        x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    else:               # 32 bits -> every second bit
        x = (x | (x << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        x = (x | (x << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        x = (x | (x << np.uint64(2))) & np.uint64(0x3333333333333333)
        x = (x | (x << np.uint64(1))) & np.uint64(0x5555555555555555)
    return x


def morton_keys(coords):
    d = coords.shape[1]
    lo = coords.min(axis=0)
    span = float((coords.max(axis=0) - lo).max()) or 1.0
    bits = 21 if d == 3 else 31
    q = np.minimum(((coords - lo) / span * (2 ** bits - 1)).astype(np.uint64), np.uint64(2 ** bits - 1))
    key = np.zeros(len(coords), dtype=np.uint64)
    for k in range(d):
        key |= _spread_bits(q[:, k], d) << np.uint64(k)
    return key


def locality_order(coords, cells, window=1024):
    """Vertex and cell order for meshes that arrive without locality: vertices along a Morton (Z-order) curve of their
    coordinates, then -- inside windows of ``window`` consecutive vertices -- by descending vertex degree, so that the 32
    rows of a SELL slice have similar lengths (SELL-C-sigma by renumbering); cells sorted by their smallest new vertex.
    Returns (new_of_old[n_vertices], cell_order[n_cells])."""
    coords, cells = np.asarray(coords), np.asarray(cells)
    nv = len(coords)
    order = np.argsort(morton_keys(coords), kind="stable")            # old id of new position
    if window > 1:
        deg = np.bincount(cells.ravel(), minlength=nv)                 # incident cells ~ row length
        pos = np.arange(nv)
        dmax = int(deg.max()) if nv else 0
        key = (pos // window).astype(np.int64) * (dmax + 1) + (dmax - deg[order])
        order = order[np.argsort(key, kind="stable")]
    new_of_old = np.empty(nv, dtype=np.int64)
    new_of_old[order] = np.arange(nv)
    cmin = new_of_old[cells].min(axis=1)
    return new_of_old, np.argsort(cmin, kind="stable")
