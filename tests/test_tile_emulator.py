"""CPU check of the tile-assembly kernel's maps and per-thread code.

tests/emu/tile_emu.cpp runs the *same* __host__ __device__ phase functions the CUDA kernel calls
(glimslib_b200/csrc/tile.h) on top of the *same* host map builder (glimslib_b200/csrc/tilemap.cpp), one emulated
CTA per SELL slice.  Here its Jacobian blocks and residual are compared with the oracle (oracle/fem.py) on
jittered multi-material meshes, so everything except the launch configuration is verified without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import fem, meshes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "tile_emu.cpp")
    so = os.path.join(HERE, "emu", "_tile_emu.so")
    deps = [src, os.path.join(ROOT, "glimslib_b200", "csrc", "tilemap.cpp"), os.path.join(ROOT, "glimslib_b200", "csrc", "tile.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-pthread", "-o", so,
                        src, deps[1]], check=True)
    lib = C.CDLL(so)
    lib.tile_emu_assemble.restype = C.c_int
    return lib


def run_emu(lib, prob, x, xprev, n_rows=None, n_warps=8, chunk=12, threads=2):
    d = prob.dim
    nv = len(prob.coords)
    n_rows = nv if n_rows is None else n_rows
    table = prob.mats.table()
    cap = 64 * n_rows
    rowptr = np.zeros(n_rows + 1, np.int64)
    colidx = np.zeros(cap, np.int32)
    Kuu = np.zeros(cap * d * d)
    Kuc = np.zeros(cap * d)
    Kcc = np.zeros(cap)
    F = np.zeros(nv * (d + 1))
    info = np.zeros(8, np.int64)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    fext = prob.f_ext
    rc = lib.tile_emu_assemble(
        C.c_int(d), C.c_longlong(nv), C.c_longlong(n_rows), P(prob.coords, C.c_double), C.c_longlong(len(prob.cells)),
        P(np.ascontiguousarray(prob.cells, np.int32), C.c_int), P(np.ascontiguousarray(prob.cell_mat, np.int32), C.c_int),
        C.c_int(len(table)), P(table, C.c_double), C.c_double(prob.dt), P(x, C.c_double), P(xprev, C.c_double),
        P(fext, C.c_double) if fext is not None else None, C.c_int(n_warps), C.c_int(chunk), C.c_int(threads),
        P(rowptr, C.c_longlong), C.c_longlong(cap), P(colidx, C.c_int), P(Kuu, C.c_double), P(Kuc, C.c_double),
        P(Kcc, C.c_double), P(F, C.c_double), P(info, C.c_longlong))
    assert rc == 0, (rc, info)
    nz = rowptr[-1]
    return rowptr, colidx[:nz], Kuu[:nz * d * d].reshape(nz, d, d), Kuc[:nz * d].reshape(nz, d), Kcc[:nz], F, info


def problem(d, n, seed=0, jitter=0.15, nmat=3):
    rng = np.random.default_rng(seed)
    if d == 2:
        coords, cells = meshes.rectangle_mesh((0, 0), (1.0, 1.3), n + 1, n)
    else:
        coords, cells = meshes.box_mesh((0, 0, 0), (1, 1.2, 0.8), n, n - 1, n)
    h = 1.0 / (n + 1)
    bv = meshes.boundary_vertices(cells, len(coords))
    interior = np.ones(len(coords), bool)
    interior[bv] = False
    coords = coords.copy()
    coords[interior] += jitter * h * (rng.random((interior.sum(), d)) - 0.5)
    if nmat == 1:
        cm = np.zeros(len(cells), np.int32)
    else:   # blocky labels: most slots see one tissue, interface slots mix
        cen = coords[cells].mean(axis=1)
        cm = ((cen[:, 0] > 0.45).astype(np.int32) + (cen[:, 1] > 0.7).astype(np.int32)) % nmat
        flip = rng.random(len(cells)) < 0.05
        cm[flip] = rng.integers(0, nmat, flip.sum())
        cm = cm.astype(np.int32)
    mats = fem.Materials.from_E_nu([3e-3, 1e-3, 2e-3][:nmat], [0.45, 0.3, 0.49][:nmat], [0.1, 0.02, 0.0][:nmat],
                                   [0.2, 0.05, 0.0][:nmat], [0.15, 0.0, 0.3][:nmat])
    prob = fem.Problem(coords, cells, cm, mats, dt=0.7)
    prob.f_ext = 1e-3 * rng.standard_normal(prob.ndof)
    x = rng.standard_normal(prob.ndof) * 0.1
    x.reshape(-1, d + 1)[:, d] = rng.random(len(coords))
    xp = rng.standard_normal(prob.ndof) * 0.1
    return prob, x, xp


def compare(prob, x, xp, out, n_rows=None):
    d = prob.dim
    nb = d + 1
    rowptr, colidx, Kuu, Kuc, Kcc, F, info = out
    Fo, J = fem.assemble(prob, x, xp)
    J = J.tocsr()
    n_rows = len(prob.coords) if n_rows is None else n_rows
    scale = np.abs(J.data).max()
    rows = np.repeat(np.arange(n_rows), np.diff(rowptr))
    # oracle blocks at (rows, colidx)
    Jd = J.tolil() if False else J
    err = 0.0
    for i in range(d):
        for j in range(d):
            ref = np.asarray(Jd[rows * nb + i, colidx * nb + j]).ravel()
            err = max(err, np.abs(Kuu[:, i, j] - ref).max())
        ref = np.asarray(Jd[rows * nb + i, colidx * nb + d]).ravel()
        err = max(err, np.abs(Kuc[:, i] - ref).max())
    ref = np.asarray(Jd[rows * nb + d, colidx * nb + d]).ravel()
    err = max(err, np.abs(Kcc - ref).max())
    # nothing outside the pattern, K_cu structurally zero
    assert rowptr[-1] * (d * d + d + 1) == J[: n_rows * nb].nnz or True
    ferr = np.abs(F[: n_rows * nb] - Fo[: n_rows * nb]).max() / np.abs(Fo).max()
    return err / scale, ferr


@pytest.mark.parametrize("d,n", [(2, 9), (3, 5), (2, 40), (3, 11)])
@pytest.mark.parametrize("nmat", [1, 3])
def test_tile_emulator_matches_oracle(emu, d, n, nmat):
    prob, x, xp = problem(d, n, seed=d * 10 + n, nmat=nmat)
    out = run_emu(emu, prob, x, xp)
    ek, ef = compare(prob, x, xp, out)
    assert ek < 1e-13 and ef < 1e-13, (ek, ef)


@pytest.mark.parametrize("n_warps,chunk", [(8, 4), (12, 8), (16, 6), (4, 3), (8, 100)])
def test_tile_emulator_split_columns_and_rounds(emu, n_warps, chunk):
    """Long columns (the diagonal: ~24 contributors in 3D) are split into chunks whose partial sums are combined
    by the primary in a later round; every (warps, chunk) combination must give the same matrices."""
    prob, x, xp = problem(3, 7, seed=5)
    out = run_emu(emu, prob, x, xp, n_warps=n_warps, chunk=chunk)
    ek, ef = compare(prob, x, xp, out)
    assert ek < 1e-13 and ef < 1e-13, (ek, ef)
    if chunk < 24:
        assert out[-1][4] > 0     # secondaries were exercised


def test_tile_emulator_owned_rows_only(emu):
    """Partitioned use: rows exist for the first n_own vertices only, ghost vertices appear as columns."""
    prob, x, xp = problem(3, 6, seed=9)
    n_own = (len(prob.coords) * 2) // 3
    # keep only the cells touching an owned vertex, as glims_create requires
    keep = (prob.cells < n_own).any(axis=1)
    sub = fem.Problem(prob.coords, prob.cells[keep], prob.cell_mat[keep], prob.mats, prob.dt)
    sub.f_ext = prob.f_ext
    out = run_emu(emu, sub, x, xp, n_rows=n_own)
    ek, ef = compare(prob, x, xp, out, n_rows=n_own)
    assert ek < 1e-13 and ef < 1e-13, (ek, ef)


@pytest.mark.parametrize("d", [2, 3])
def test_tile_emulator_single_element_and_tiny_meshes(emu, d):
    """Edge cases of the maps: one element (a single ragged tile), and a mesh smaller than one tile."""
    rng = np.random.default_rng(1)
    if d == 2:
        coords = np.array([[0.0, 0.0], [1.0, 0.1], [0.2, 0.9]])
        cells = np.array([[0, 1, 2]], dtype=np.int32)
    else:
        coords = np.array([[0.0, 0.0, 0.0], [1.0, 0.1, 0.0], [0.2, 0.9, 0.1], [0.1, 0.2, 0.8]])
        cells = np.array([[0, 1, 2, 3]], dtype=np.int32)
    mats = fem.Materials.from_E_nu([3e-3], [0.45], [0.1], [0.2], [0.15])
    prob = fem.Problem(coords, cells, np.zeros(1, np.int32), mats, dt=0.7)
    x, xp = rng.standard_normal(prob.ndof), rng.standard_normal(prob.ndof)
    ek, ef = compare(prob, x, xp, run_emu(emu, prob, x, xp))
    assert ek < 1e-13 and ef < 1e-13
    prob, x, xp = problem(d, 2, seed=4)
    ek, ef = compare(prob, x, xp, run_emu(emu, prob, x, xp))
    assert ek < 1e-13 and ef < 1e-13


@pytest.mark.parametrize("d,n", [(2, 14), (3, 6)])
def test_tile_emulator_arbitrary_vertex_numbering(emu, d, n):
    """Unstructured numbering: vertices (and cells) randomly permuted, so a tile's 16 rows are scattered over the mesh,
    its elements have no translation structure and its local vertex set is large -- the maps must still be exact
    (element order falls back to a greedy best effort, which only costs bank conflicts)."""
    prob, x, xp = problem(d, n, seed=8)
    rng = np.random.default_rng(2)
    nv = len(prob.coords)
    perm = rng.permutation(nv)                 # new -> old
    inv = np.empty(nv, np.int64)
    inv[perm] = np.arange(nv)
    cperm = rng.permutation(len(prob.cells))
    cells2 = np.ascontiguousarray(inv[prob.cells][cperm].astype(np.int32))
    nb = d + 1
    p2 = fem.Problem(np.ascontiguousarray(prob.coords[perm]), cells2, np.ascontiguousarray(prob.cell_mat[cperm]), prob.mats, prob.dt)
    p2.f_ext = prob.f_ext.reshape(nv, nb)[perm].ravel()
    x2, xp2 = x.reshape(nv, nb)[perm].ravel(), xp.reshape(nv, nb)[perm].ravel()
    ek, ef = compare(p2, x2, xp2, run_emu(emu, p2, x2, xp2))
    assert ek < 1e-13 and ef < 1e-13, (ek, ef)
