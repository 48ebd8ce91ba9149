"""Synthetic configurations of BASELINE.json as engine inputs.

Each builder returns a dict: mesh, cell_mat, table (mu, lambda, D, rho, gamma per material),
bc_dofs / bc_vals (displacement clamped on the whole exterior boundary), x0 (nodal interpolant of the
initial condition, helper_classes.py:983-986), dt.  Deterministic, no RNG.
"""
import numpy as np

from . import mesh as M


def lame(E, nu):
    """math_linear_elasticity.py:6-10"""
    E, nu = np.asarray(E, dtype=np.float64), np.asarray(nu, dtype=np.float64)
    return E / (2.0 * (1.0 + nu)), E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))


def material_table(E, nu, D, rho, gamma):
    mu, lam = lame(E, nu)
    return np.ascontiguousarray(np.stack([mu, lam, np.asarray(D, float), np.asarray(rho, float),
                                          np.asarray(gamma, float)], axis=1))


def _clamp_displacement(mesh):
    d = mesh.dim
    bv = mesh.boundary_vertices()
    dofs = np.sort((bv[:, None] * (d + 1) + np.arange(d)[None, :]).ravel()).astype(np.int64)
    return dofs, np.zeros(len(dofs))


def _pack(name, mesh, cell_mat, table, x0, dt, **extra):
    dofs, vals = _clamp_displacement(mesh)
    out = dict(name=name, mesh=mesh, cell_mat=np.ascontiguousarray(cell_mat, dtype=np.int32), table=table,
               bc_dofs=dofs, bc_vals=vals, x0=x0, dt=float(dt))
    out.update(extra)
    return out


def c1_2d_subdomains(nx=50):
    """test_case_simulation_tumor_growth_2D_subdomains.py:34-89"""
    mesh = M.rectangle_mesh((-5, -5), (5, 5), nx, nx)
    lab_v = np.where(mesh.coords[:, 0] >= 0, 1.0, 2.0)
    lab = lab_v[mesh.cells].mean(axis=1).astype(int)     # helper_classes.py:441-442 on a DG1 label function
    table = material_table([1e-3, 1e-3], [0.4, 0.1], [0.1, 0.0], [0.1, 0.0], [0.2, 0.0])
    x0 = np.zeros(mesh.num_vertices() * 3)
    x0[2::3] = (np.hypot(mesh.coords[:, 0] - 2.5, mesh.coords[:, 1] - 2.5) < 0.4).astype(float)
    return _pack("C1 2D 50x50 two subdomains coupled", mesh, lab - 1, table, x0, 1.0)


def c2_2d_1m(n=707):
    """2D 1M-cell single-label mesh; the reference has no scalar-only class, so coupling=0 (u == 0)."""
    mesh = M.rectangle_mesh((0, 0), (1, 1), n, n)
    table = material_table([1e-3], [0.4], [1e-4], [0.1], [0.0])
    x0 = np.zeros(mesh.num_vertices() * 3)
    x0[2::3] = np.exp(-200.0 * ((mesh.coords - 0.5) ** 2).sum(axis=1))
    return _pack("C2 2D %dx%d concentration-driven (coupling=0)" % (n, n), mesh, np.zeros(mesh.num_cells()), table, x0, 1.0)


def c3_box(n=55):
    """3D unit box, two tissues split at x=0.5, brain-like values scaled to the box
    (test_case_comparison_2D_atlas.py:84-122)."""
    mesh = M.box_mesh((0, 0, 0), (1, 1, 1), n, n, n)
    cm = (mesh.cell_midpoints()[:, 0] >= 0.5).astype(np.int32)
    table = material_table([3e-3, 1e-3], [0.45, 0.45], [2e-4, 0.0], [0.05, 0.0], [0.1, 0.0])
    x0 = np.zeros(mesh.num_vertices() * 4)
    x0[3::4] = np.exp(-60.0 * ((mesh.coords - np.array([0.3, 0.5, 0.5])) ** 2).sum(axis=1))
    return _pack("C3 3D box %d^3 two tissues coupled" % n, mesh, cm, table, x0, 1.0)


def c4_ellipsoid(n=148):
    """Brain-like ellipsoid, semi-axes (70, 85, 60) mm, three tissues in radial shells with the id-sorted
    map {1: CSF, 2: GM, 3: WM} and the atlas values of test_case_simulation_tumor_growth_3D_atlas.py:97-117.
    n=148 gives 10.19M tetrahedra / 1.75M vertices."""
    mesh, r = M.voxel_ellipsoid_mesh(n, (70.0, 85.0, 60.0))
    label = np.where(r < 0.6, 3, np.where(r < 0.85, 2, 1))                 # WM / GM / CSF
    #                 CSF      GM      WM
    table = material_table([1e-3, 3e-3, 3e-3], [0.47, 0.4, 0.4], [0.0, 0.02, 0.1], [0.0, 0.05, 0.05],
                           [0.0, 0.1, 0.1])
    x0 = np.zeros(mesh.num_vertices() * 4)
    seed = np.array([20.0, -15.0, 10.0])
    x0[3::4] = np.exp(-0.02 * ((mesh.coords - seed) ** 2).sum(axis=1))
    return _pack("C4 3D voxel ellipsoid n=%d three tissues (CSF/GM/WM) coupled" % n, mesh, label - 1, table, x0, 1.0,
                 labels=label)


def c5_ellipsoid(n=252):
    """Config C5's mesh size on the C4 geometry: n=252 gives 50.3M tetrahedra / 8.6M vertices."""
    w = c4_ellipsoid(n)
    w["name"] = "C5 3D voxel ellipsoid n=%d (50M-tet class) three tissues coupled" % n
    return w


# ---- vertex / cell renumbering (unstructured-ordering studies and the library's locality reorder) ------------------
def renumber(w, new_of_old, cell_order=None):
    """Workload ``w`` with vertex ``v`` renamed ``new_of_old[v]`` (and cells listed in ``cell_order``)."""
    mesh = w["mesh"]
    nb = mesh.dim + 1
    new_of_old = np.asarray(new_of_old, dtype=np.int64)
    old_of_new = np.empty_like(new_of_old)
    old_of_new[new_of_old] = np.arange(len(new_of_old))
    cells = new_of_old[mesh.cells].astype(np.int32)
    cell_mat = w["cell_mat"]
    if cell_order is not None:
        cells, cell_mat = cells[cell_order], cell_mat[cell_order]
    m2 = M.SimplexMesh(mesh.coords[old_of_new], cells)
    bv = getattr(mesh, "_boundary_vertices", None)
    if bv is not None:
        m2._boundary_vertices = np.sort(new_of_old[bv])
    v, k = w["bc_dofs"] // nb, w["bc_dofs"] % nb
    dofs = new_of_old[v] * nb + k
    o = np.argsort(dofs, kind="stable")
    out = dict(w)
    out.update(mesh=m2, cell_mat=np.ascontiguousarray(cell_mat, dtype=np.int32), bc_dofs=dofs[o].astype(np.int64),
               bc_vals=np.asarray(w["bc_vals"])[o], x0=np.ascontiguousarray(w["x0"].reshape(-1, nb)[old_of_new].ravel()),
               new_of_old=new_of_old)
    return out


def random_numbering(w, seed=0):
    """Random vertex and cell numbering: what an unstructured mesh generator without any locality would hand over."""
    rng = np.random.default_rng(seed)
    nv, nc = w["mesh"].num_vertices(), w["mesh"].num_cells()
    out = renumber(w, rng.permutation(nv), rng.permutation(nc))
    out["name"] = w["name"] + " [random vertex/cell numbering, seed %d]" % seed
    return out


def locality_numbering(w, window=1024):
    """The library's internal reorder for meshes that arrive without locality (``mesh.locality_order``: Morton curve +
    degree-sorted windows + cells by smallest vertex), applied to a whole workload.  ``new_of_old`` holds the vertex
    permutation: it is what the dof-permutation API reports."""
    new_of_old, cell_order = M.locality_order(w["mesh"].coords, w["mesh"].cells, window)
    out = renumber(w, new_of_old, cell_order)
    out["name"] = w["name"] + " [Morton + degree-windowed renumbering]"
    return out


def build_engine(w, device=0):
    from .engine import Engine
    eng = Engine(w["mesh"].coords, w["mesh"].cells, w["cell_mat"], device=device)
    eng.set_materials(w["table"])
    eng.set_dt(w["dt"])
    eng.set_dirichlet(w["bc_dofs"], w["bc_vals"])
    return eng
