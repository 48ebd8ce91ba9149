"""Structured simplex meshes with DOLFIN's vertex / cell numbering.  (oracle: test infrastructure)

The reference builds its self-contained cases with ``fenics.RectangleMesh`` /
``fenics.BoxMesh`` (``test_cases/test_simulation_tumor_growth/
test_case_simulation_tumor_growth_2D_subdomains.py:34-35``,
``..._2D_uniform.py:34``).  Numbering below follows DOLFIN 2017.2's
``RectangleMesh.cpp`` ("right" diagonal) and ``BoxMesh.cpp`` (6 tets per
hexahedron) as recalled [MEM] -- SURVEY.md section 8c item 3: reproducible and
asserted here, not verifiable against a DOLFIN install in this container.
"""
import numpy as np


def rectangle_mesh(p0, p1, nx, ny, diagonal="right"):
    """Vertices row-major in (iy, ix); two triangles per quad, 'right' diagonal
    = the diagonal from lower-left to upper-right: (v0,v1,v3),(v0,v2,v3)."""
    x = np.linspace(p0[0], p1[0], nx + 1)
    y = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(x, y, indexing="xy")          # shape (ny+1, nx+1)
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v1 + (nx + 1)
    if diagonal == "right":
        t0 = np.stack([v0, v1, v3], axis=1)
        t1 = np.stack([v0, v2, v3], axis=1)
    elif diagonal == "left":
        t0 = np.stack([v0, v1, v2], axis=1)
        t1 = np.stack([v1, v2, v3], axis=1)
    else:
        raise ValueError("diagonal must be 'right' or 'left'")
    cells = np.empty((2 * nx * ny, 3), dtype=np.int32)
    cells[0::2] = t0
    cells[1::2] = t1
    return np.ascontiguousarray(coords, dtype=np.float64), cells


def box_mesh(p0, p1, nx, ny, nz):
    """Vertices numbered iz-major, then iy, then ix; 6 tets per hexahedron all
    sharing the v0-v7 diagonal."""
    x = np.linspace(p0[0], p1[0], nx + 1)
    y = np.linspace(p0[1], p1[1], ny + 1)
    z = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")    # (nz+1, ny+1, nx+1)
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    v0 = (iz * sy + iy * sx + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + sx
    v3 = v1 + sx
    v4 = v0 + sy
    v5 = v1 + sy
    v6 = v2 + sy
    v7 = v3 + sy
    tets = [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
            (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
    cells = np.empty((6 * nx * ny * nz, 4), dtype=np.int32)
    for k, t in enumerate(tets):
        cells[k::6] = np.stack(t, axis=1)
    return np.ascontiguousarray(coords, dtype=np.float64), cells


def ellipsoid_map(coords, semi_axes):
    """Smooth bijection of the cube [-1,1]^3 onto the ellipsoid with the given
    semi-axes: x -> x * |x|_inf / |x|_2."""
    c = np.asarray(coords, dtype=np.float64)
    ninf = np.abs(c).max(axis=1)
    n2 = np.sqrt((c * c).sum(axis=1))
    s = np.where(n2 > 0, ninf / np.where(n2 > 0, n2, 1.0), 1.0)
    return c * s[:, None] * np.asarray(semi_axes, dtype=np.float64)[None, :]


def boundary_vertices(cells, n_vertices):
    """Vertices on exterior facets (facets that belong to exactly one cell) --
    what ``DirichletBC(V, g, on_boundary)`` touches topologically."""
    f, _ = exterior_facets(cells)
    mask = np.zeros(n_vertices, dtype=bool)
    mask[f.ravel()] = True
    return np.nonzero(mask)[0]


def all_facets(cells):
    """(facets[nf, d] sorted vertex tuples unique, cell_facet[nc, d+1] ids,
    facet_ncells[nf])."""
    nc, nv = cells.shape
    fl = []
    for k in range(nv):
        fl.append(np.delete(cells, k, axis=1))
    f = np.sort(np.concatenate(fl, axis=0), axis=1)
    uniq, inv, cnt = np.unique(f, axis=0, return_inverse=True, return_counts=True)
    cell_facet = inv.reshape(nv, nc).T
    return uniq, cell_facet, cnt


def exterior_facets(cells):
    """(facets[nf_ext, d], owning cell index)."""
    nc, nv = cells.shape
    uniq, cell_facet, cnt = all_facets(cells)
    ext = np.nonzero(cnt == 1)[0]
    owner = np.full(len(uniq), -1, dtype=np.int64)
    for k in range(nv):
        owner[cell_facet[:, k]] = np.arange(nc)
    return uniq[ext], owner[ext]
