"""GPU parity tests: every number the CUDA path produces is compared with the CPU oracle on the
same seeded inputs, through the C ABI (ctypes).  FP64 path; tolerances are stated per test."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fem, meshes, solver as osolver

pytestmark = pytest.mark.gpu


def _engine(*a, **k):
    from glimslib_b200.engine import Engine
    return Engine(*a, **k)


def small_problem(d, seed=0, n=4, jitter=0.15, with_bc=True):
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
    cm = rng.integers(0, 3, len(cells)).astype(np.int32)
    mats = fem.Materials.from_E_nu([3e-3, 1e-3, 2e-3], [0.45, 0.3, 0.49], [0.1, 0.02, 0.0], [0.2, 0.05, 0.0],
                                   [0.15, 0.0, 0.3])
    nb = d + 1
    if with_bc:
        dofs = (bv[:, None] * nb + np.arange(d)[None, :]).ravel()
        vals = 0.01 * rng.standard_normal(len(dofs))
        # a few concentration Dirichlet dofs too (sub-space 1 BCs are legal, helper_classes.py:673-723)
        cd = bv[::7] * nb + d
        dofs = np.concatenate([dofs, cd])
        vals = np.concatenate([vals, 0.3 * np.ones(len(cd))])
    else:
        dofs, vals = np.zeros(0, np.int64), np.zeros(0)
    prob = fem.Problem(coords, cells, cm, mats, dt=0.7, bc_dofs=dofs.astype(np.int64), bc_vals=vals)
    prob.f_ext = 1e-3 * rng.standard_normal(prob.ndof)
    return prob, rng


def make_engine(prob):
    eng = _engine(prob.coords, prob.cells, prob.cell_mat)
    eng.set_materials(prob.mats.table())
    eng.set_dt(prob.dt)
    eng.set_dirichlet(prob.bc_dofs, prob.bc_vals)
    eng.set_load(prob.f_ext)
    return eng


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("kernel", [0, 1, 2, 3, 4])
def test_assembly_matches_oracle(d, kernel):
    """K1/K2: residual and Jacobian values, raw and with DOLFIN-style Dirichlet rows; <= 1e-13 relative."""
    prob, rng = small_problem(d)
    x, xp = rng.standard_normal(prob.ndof), rng.standard_normal(prob.ndof)
    eng = make_engine(prob)
    eng.set_state(x)
    eng.set_prev(xp)
    F0, J0 = fem.assemble(prob, x, xp)
    eng.assemble(kernel=kernel, apply_bc=0)
    assert relerr(eng.residual(), F0) < 1e-13
    J = eng.export_jacobian()
    assert abs(J - J0).max() / abs(J0).max() < 1e-13
    F1, J1 = fem.apply_dirichlet(prob, F0, J0, x)
    eng.assemble(kernel=kernel, apply_bc=1)
    assert relerr(eng.residual(), F1) < 1e-13
    J = eng.export_jacobian()
    assert abs(J - J1).max() / abs(J1).max() < 1e-13
    # rows + columns (symmetric elimination) keeps the symmetric blocks symmetric
    eng.assemble(kernel=kernel, apply_bc=2)
    J = eng.export_jacobian()
    nb = d + 1
    keep = np.ones(prob.ndof)
    keep[prob.bc_dofs] = 0
    J2 = (sp.diags(keep) @ J0 @ sp.diags(keep) + sp.diags(1 - keep)).tocsr()
    assert abs(J - J2).max() / abs(J2).max() < 1e-13
    eng.close()


@pytest.mark.parametrize("d", [2, 3])
def test_atomic_and_gather_kernels_agree(d):
    prob, rng = small_problem(d, seed=3, n=6)
    eng = make_engine(prob)
    eng.set_state(rng.standard_normal(prob.ndof))
    eng.assemble(what=6, kernel=0)
    a = eng.export_blocks()
    for kernel in (1, 2, 3, 4):
        eng.assemble(what=6, kernel=kernel)
        b = eng.export_blocks()
        for p, q in zip(a[2:], b[2:]):
            assert relerr(p, q) < 1e-13
    eng.close()


@pytest.mark.parametrize("d,n", [(2, 40), (3, 12)])
@pytest.mark.parametrize("nt,chunk", [(128, 12), (256, 8), (192, 5), (192, 100)])
def test_tile_kernel_all_masks_and_configs(d, n, nt, chunk):
    """Fused tile kernel (GLIMS_ASMK_TILE): every `what` mask leaves exactly the requested outputs equal to the oracle,
    for every CTA size / column-split setting (split columns combine partial sums across warps), on a jittered
    three-tissue mesh with several slices and a ragged last slice; <= 1e-13 relative."""
    prob, rng = small_problem(d, seed=7, n=n, with_bc=False)
    if chunk != 12:    # blocky tissue labels: most slots see one material (fast path), interface slots mix
        cen = prob.coords[prob.cells].mean(axis=1)
        prob.cell_mat = ((cen[:, 0] > 0.45).astype(np.int32) + (cen[:, 1] > 0.7).astype(np.int32)).astype(np.int32)
    x, xp = 0.1 * rng.standard_normal(prob.ndof), 0.1 * rng.standard_normal(prob.ndof)
    eng = make_engine(prob)
    eng.tile_config(nt, chunk)
    eng.set_state(x)
    eng.set_prev(xp)
    F0, J0 = fem.assemble(prob, x, xp)
    eng.assemble(what=7, kernel=3)
    info = eng.tile_info()
    assert info["threads_per_cta"] == nt and info["smem_bytes_per_cta"] <= 227 * 1024
    assert relerr(eng.residual(), F0) < 1e-13
    assert abs(eng.export_jacobian() - J0).max() / abs(J0).max() < 1e-13
    # a different state: residual-only and residual+K_cc must refresh F (and K_cc) but leave K_uu/K_uc alone
    x2 = 0.1 * rng.standard_normal(prob.ndof)
    eng.set_state(x2)
    F2, J2 = fem.assemble(prob, x2, xp)
    eng.assemble(what=1, kernel=3)
    assert relerr(eng.residual(), F2) < 1e-13
    assert abs(eng.export_jacobian() - J0).max() / abs(J0).max() < 1e-13
    eng.assemble(what=5, kernel=3)
    assert relerr(eng.residual(), F2) < 1e-13
    assert abs(eng.export_jacobian() - J2).max() / abs(J2).max() < 1e-13
    # deterministic: bitwise identical on repetition
    Fa = eng.residual().copy()
    Ka = eng.export_blocks()
    eng.assemble(what=7, kernel=3)
    assert np.array_equal(eng.residual(), Fa)
    for p, q in zip(Ka[2:], eng.export_blocks()[2:]):
        assert np.array_equal(p, q)
    eng.close()


@pytest.mark.parametrize("d,n", [(2, 40), (3, 12)])
def test_rows_kernel_masks_determinism_and_numbering(d, n):
    """Row-walk assembly (GLIMS_ASMK_ROWS, the default of glims_step): F_c / K_cc from per-slot constants + (row, element)
    pair lists, F_u by SpMV over the stored K_uu / K_uc.  Every `what` mask leaves exactly the requested outputs equal to
    the oracle (<= 1e-13), results are bitwise repeatable, a change of dt or of the materials refreshes the constants, and a
    random vertex / cell numbering (ragged rows, scattered columns) gives the same answer."""
    prob, rng = small_problem(d, seed=17, n=n, with_bc=False)
    x, xp = 0.1 * rng.standard_normal(prob.ndof), 0.1 * rng.standard_normal(prob.ndof)
    eng = make_engine(prob)
    eng.set_state(x)
    eng.set_prev(xp)
    F0, J0 = fem.assemble(prob, x, xp)
    eng.assemble(what=7, kernel=4)
    assert relerr(eng.residual(), F0) < 1e-13
    assert abs(eng.export_jacobian() - J0).max() / abs(J0).max() < 1e-13
    x2 = 0.1 * rng.standard_normal(prob.ndof)
    eng.set_state(x2)
    F2, J2 = fem.assemble(prob, x2, xp)
    eng.assemble(what=1, kernel=4)                       # residual only: K_cc untouched
    assert relerr(eng.residual(), F2) < 1e-13
    assert abs(eng.export_jacobian() - J0).max() / abs(J0).max() < 1e-13
    Fa = eng.residual().copy()
    eng.assemble(what=4, kernel=4)                       # K_cc only: F untouched
    assert np.array_equal(Fa, eng.residual())
    assert abs(eng.export_jacobian() - J2).max() / abs(J2).max() < 1e-13
    eng.assemble(what=5, kernel=4)
    Ka = eng.export_blocks()
    Fa = eng.residual().copy()
    eng.assemble(what=5, kernel=4)                       # bitwise repeatable
    assert np.array_equal(eng.residual(), Fa)
    for p, q in zip(Ka[2:], eng.export_blocks()[2:]):
        assert np.array_equal(p, q)
    # new dt and new materials: the per-slot constants (Klin, M, rho|K|) follow
    prob.dt = 0.31
    eng.set_dt(prob.dt)
    F3, J3 = fem.assemble(prob, x2, xp)
    eng.assemble(what=7, kernel=4)
    assert relerr(eng.residual(), F3) < 1e-13
    assert abs(eng.export_jacobian() - J3).max() / abs(J3).max() < 1e-13
    prob.mats = fem.Materials.from_E_nu([1e-3, 2e-3, 5e-3], [0.4, 0.2, 0.3], [0.03, 0.0, 0.2], [0.0, 0.3, 0.1], [0.1, 0.2, 0.0])
    eng.set_materials(prob.mats.table())
    F4, J4 = fem.assemble(prob, x2, xp)
    eng.assemble(what=7, kernel=4)
    assert relerr(eng.residual(), F4) < 1e-13
    assert abs(eng.export_jacobian() - J4).max() / abs(J4).max() < 1e-13
    eng.close()
    # random numbering
    nv, nb = len(prob.coords), d + 1
    perm = rng.permutation(nv)
    inv = np.empty(nv, np.int64)
    inv[perm] = np.arange(nv)
    cperm = rng.permutation(len(prob.cells))
    p2 = fem.Problem(np.ascontiguousarray(prob.coords[perm]), np.ascontiguousarray(inv[prob.cells][cperm].astype(np.int32)),
                     np.ascontiguousarray(prob.cell_mat[cperm]), prob.mats, prob.dt)
    p2.f_ext = prob.f_ext.reshape(nv, nb)[perm].ravel()
    x, xp = 0.1 * rng.standard_normal(p2.ndof), 0.1 * rng.standard_normal(p2.ndof)
    eng = make_engine(p2)
    eng.set_state(x)
    eng.set_prev(xp)
    F0, J0 = fem.assemble(p2, x, xp)
    eng.assemble(what=7, kernel=4)
    assert relerr(eng.residual(), F0) < 1e-13
    assert abs(eng.export_jacobian() - J0).max() / abs(J0).max() < 1e-13
    eng.close()


@pytest.mark.parametrize("d", [2, 3])
def test_time_dependent_dirichlet_values_keep_the_hierarchy(d):
    """Dirichlet VALUES that change from step to step on an unchanged dof set (the reference's time_update_bcs,
    simulation_base.py:277-300): the library re-uploads the values and rebuilds only the lift of the eliminated columns
    (K_uu, the AMG hierarchy and the captured graphs stay), and every step still matches the oracle <= 1e-8."""
    prob, rng = small_problem(d, seed=23, n=8 if d == 3 else 20, jitter=0.2)
    prob.f_ext = None
    nb = d + 1
    x0 = np.zeros(prob.ndof)
    x0[nb - 1::nb] = np.exp(-8 * ((prob.coords - prob.coords.mean(axis=0)) ** 2).sum(axis=1))
    base_vals = prob.bc_vals.copy()
    eng = make_engine(prob)
    eng.set_load(None)
    eng.set_prev(x0)
    eng.set_state(np.zeros(prob.ndof))
    x_prev, x_start = x0.copy(), np.zeros(prob.ndof)
    for k in range(1, 4):
        is_u = (prob.bc_dofs % nb) < d
        prob.bc_vals = np.where(is_u, base_vals * (1.0 + 0.5 * k), base_vals)    # u data grows in time, c data fixed
        eng.set_dirichlet(prob.bc_dofs, prob.bc_vals)
        l0 = eng.launch_count
        st = eng.step(1, snes_rtol=1e-11, snes_atol=1e-14, ksp_rtol=1e-12)[0]
        assert st["converged"] == 1
        x_ref, _ = osolver.newton(prob, x_start.copy(), x_prev, linear="lu", rtol=1e-12, atol=1e-15)
        x = eng.get_state().reshape(-1, nb)
        ref = x_ref.reshape(-1, nb)
        assert np.linalg.norm(x[:, d] - ref[:, d]) / np.linalg.norm(ref[:, d]) < 1e-8, k
        assert np.linalg.norm(x[:, :d] - ref[:, :d]) / np.linalg.norm(ref[:, :d]) < 1e-8, k
        x_prev, x_start = x_ref.copy(), x_ref.copy()
        eng.set_prev(x_ref)
        eng.set_state(x_ref)
    eng.close()


@pytest.mark.parametrize("d", [2, 3])
def test_dof_permutation_and_internal_reorder(d):
    """glims_set_dof_permutation (SURVEY 8b `dof_permutation`): host vectors and Dirichlet dofs in an arbitrary caller
    numbering -- here the UFC 'unreordered' mixed numbering DOLFIN uses with reorder_dofs_serial=False (component-major
    blocks of vertex indices, data_io.py:242-252) -- give the same fields; and Engine(reorder='morton') renumbers the mesh
    internally for locality while every result stays in the caller's numbering."""
    from glimslib_b200.engine import Engine
    prob, rng = small_problem(d, seed=31, n=8 if d == 3 else 20, jitter=0.2)
    nb = d + 1
    nv = len(prob.coords)
    x0 = np.zeros(prob.ndof)
    x0[nb - 1::nb] = np.exp(-8 * ((prob.coords - prob.coords.mean(axis=0)) ** 2).sum(axis=1))
    kw = dict(snes_rtol=1e-11, snes_atol=1e-14, ksp_rtol=1e-12)
    ref = make_engine(prob)
    ref.set_prev(x0)
    ref.set_state(np.zeros(prob.ndof))
    ref.step(2, **kw)
    x_ref = ref.get_state()
    f_ref = ref.cell_fields()
    fv_ref = ref.cell_fields(vertex=True)
    ref.close()
    # (a) component-major caller numbering: caller dof k*nv + v  <->  vertex-blocked v*nb + k
    perm = (np.arange(nv)[None, :] * nb + np.arange(nb)[:, None]).ravel()
    to_caller = lambda x: x.reshape(nv, nb).T.ravel()
    eng = Engine(prob.coords, prob.cells, prob.cell_mat)
    eng.set_dof_permutation(perm)
    assert np.array_equal(eng.get_dof_permutation(), perm)
    eng.set_materials(prob.mats.table())
    eng.set_dt(prob.dt)
    caller_bc = (prob.bc_dofs % nb) * nv + prob.bc_dofs // nb
    eng.set_dirichlet(caller_bc, prob.bc_vals)
    eng.set_load(to_caller(prob.f_ext))
    eng.set_prev(to_caller(x0))
    eng.set_state(np.zeros(prob.ndof))
    eng.step(2, **kw)
    assert np.abs(eng.get_state() - to_caller(x_ref)).max() <= 1e-12 * np.abs(x_ref).max()
    eng.close()
    # (b) internal Morton reorder on a randomly numbered copy of the mesh: results come back in the caller's numbering
    pv = rng.permutation(nv)                  # new id of old vertex
    inv = np.empty(nv, np.int64)
    inv[pv] = np.arange(nv)
    cperm = rng.permutation(len(prob.cells))
    coords2 = np.ascontiguousarray(prob.coords[inv])
    cells2 = np.ascontiguousarray(pv[prob.cells][cperm].astype(np.int32))
    vec2 = lambda x: np.ascontiguousarray(x.reshape(nv, nb)[inv].ravel())
    eng = Engine(coords2, cells2, prob.cell_mat[cperm], reorder="morton")
    eng.set_materials(prob.mats.table())
    eng.set_dt(prob.dt)
    eng.set_dirichlet(pv[prob.bc_dofs // nb] * nb + prob.bc_dofs % nb, prob.bc_vals)
    eng.set_load(vec2(prob.f_ext))
    eng.set_prev(vec2(x0))
    eng.set_state(np.zeros(prob.ndof))
    eng.step(2, **kw)
    x2 = eng.get_state()
    assert np.abs(x2 - vec2(x_ref)).max() <= 1e-9 * np.abs(x_ref).max()
    f2 = eng.cell_fields()
    assert np.abs(f2["von_mises"] - f_ref["von_mises"][cperm]).max() <= 1e-8 * np.abs(f_ref["von_mises"]).max()
    fv2 = eng.cell_fields(vertex=True)
    assert np.abs(fv2["pressure"] - fv_ref["pressure"][inv]).max() <= 1e-8 * np.abs(fv_ref["pressure"]).max()
    eng.close()


@pytest.mark.parametrize("config", ["c1", "c3_small"])
def test_default_tolerances_meet_the_parity_bar(config):
    """The tolerances bench.py times (glims_default_opts: SNES rtol 1e-9 / atol 1e-10, KSP rtol 1e-10) against the oracle
    solved to 1e-12: <= 1e-8 relative L2 per field and step (the timed configuration must be the
    parity-tested one)."""
    if config == "c1":
        prob, x0 = c1_problem()
        nb, steps = 3, 10
    else:
        coords, cells = meshes.box_mesh((0, 0, 0), (1, 1, 1), 10, 10, 10)
        cm = (coords[cells].mean(axis=1)[:, 0] >= 0.5).astype(np.int32)
        mats = fem.Materials.from_E_nu([3e-3, 1e-3], [0.45, 0.45], [2e-4, 0.0], [0.05, 0.0], [0.1, 0.0])
        bv = meshes.boundary_vertices(cells, len(coords))
        dofs = np.sort((bv[:, None] * 4 + np.arange(3)[None, :]).ravel())
        prob = fem.Problem(coords, cells, cm, mats, dt=1.0, bc_dofs=dofs, bc_vals=np.zeros(len(dofs)))
        x0 = np.zeros(prob.ndof)
        x0[3::4] = np.exp(-60.0 * ((coords - np.array([0.3, 0.5, 0.5])) ** 2).sum(axis=1))
        nb, steps = 4, 5
    d = nb - 1
    recs, _ = osolver.run(prob, x0, steps, linear="lu", rtol=1e-12, atol=1e-15)
    eng = make_engine(prob)
    eng.set_prev(x0)
    eng.set_state(np.zeros(prob.ndof))
    worst = 0.0
    for k in range(1, steps + 1):
        st = eng.step(1)[0]                      # library defaults
        assert st["converged"] == 1
        x = eng.get_state().reshape(-1, nb)
        ref = recs[k][2].reshape(-1, nb)
        ec = np.linalg.norm(x[:, d] - ref[:, d]) / np.linalg.norm(ref[:, d])
        eu = np.linalg.norm(x[:, :d] - ref[:, :d]) / np.linalg.norm(ref[:, :d])
        worst = max(worst, ec, eu)
        assert ec < 1e-8 and eu < 1e-8, (config, k, ec, eu)
    print("default tolerances, %s: worst relative L2 error over %d steps %.2e" % (config, steps, worst))
    eng.close()


@pytest.mark.parametrize("d", [2, 3])
def test_tile_kernel_arbitrary_vertex_numbering(d):
    """Unstructured numbering (vertices and cells randomly permuted): a tile's 16 rows are scattered over the mesh, its
    local vertex set is large and its elements have no translation structure; the fused tile kernel must still
    reproduce the oracle (<= 1e-13) -- only its bank-conflict-free element order degrades."""
    prob, rng = small_problem(d, seed=13, n=10, with_bc=False)
    nv, nb = len(prob.coords), d + 1
    perm = rng.permutation(nv)
    inv = np.empty(nv, np.int64)
    inv[perm] = np.arange(nv)
    cperm = rng.permutation(len(prob.cells))
    p2 = fem.Problem(np.ascontiguousarray(prob.coords[perm]), np.ascontiguousarray(inv[prob.cells][cperm].astype(np.int32)),
                     np.ascontiguousarray(prob.cell_mat[cperm]), prob.mats, prob.dt)
    p2.f_ext = prob.f_ext.reshape(nv, nb)[perm].ravel()
    x, xp = 0.1 * rng.standard_normal(p2.ndof), 0.1 * rng.standard_normal(p2.ndof)
    eng = make_engine(p2)
    eng.set_state(x)
    eng.set_prev(xp)
    F0, J0 = fem.assemble(p2, x, xp)
    eng.assemble(what=7, kernel=3)
    assert relerr(eng.residual(), F0) < 1e-13
    assert abs(eng.export_jacobian() - J0).max() / abs(J0).max() < 1e-13
    eng.close()


@pytest.mark.parametrize("d", [2, 3])
def test_spmv_matches_oracle(d):
    """K4: SELL-32 SpMV of the monolithic Jacobian and of each block; <= 1e-13 relative."""
    prob, rng = small_problem(d, seed=1, n=5)
    x = rng.standard_normal(prob.ndof)
    eng = make_engine(prob)
    eng.set_state(x)
    eng.set_prev(x)
    eng.assemble(apply_bc=1)
    _, J = fem.assemble(prob, x, x)
    _, J = fem.apply_dirichlet(prob, np.zeros(prob.ndof), J, x)
    v = rng.standard_normal(prob.ndof)
    assert relerr(eng.spmv(0, v), J @ v) < 1e-13
    nb = d + 1
    iu = (np.arange(prob.ndof) % nb) < d
    vu, vc = rng.standard_normal(iu.sum()), rng.standard_normal((~iu).sum())
    assert relerr(eng.spmv(1, vu), J[iu][:, iu] @ vu) < 1e-13
    assert relerr(eng.spmv(2, vc), J[~iu][:, ~iu] @ vc) < 1e-13
    eng.close()


def c1_problem(nx=50):
    """test_case_simulation_tumor_growth_2D_subdomains.py:34-89 (config C1)."""
    coords, cells = meshes.rectangle_mesh((-5, -5), (5, 5), nx, nx)
    lab_v = np.where(coords[:, 0] >= 0, 1.0, 2.0)
    lab = lab_v[cells].mean(axis=1).astype(int)          # helper_classes.py:441-442 on the DG1 label function
    mats = fem.Materials.from_E_nu([1e-3, 1e-3], [0.4, 0.1], [0.1, 0.0], [0.1, 0.0], [0.2, 0.0])
    bv = meshes.boundary_vertices(cells, len(coords))
    dofs = np.sort(np.concatenate([bv * 3, bv * 3 + 1]))
    prob = fem.Problem(coords, cells, (lab - 1).astype(np.int32), mats, dt=1.0, bc_dofs=dofs,
                       bc_vals=np.zeros(len(dofs)))
    x0 = np.zeros(prob.ndof)
    r = np.hypot(coords[:, 0] - 2.5, coords[:, 1] - 2.5)
    x0[2::3] = (r < 0.4).astype(float)
    return prob, x0


@pytest.mark.parametrize("mode", ["block_tri_jacobi", "block_tri_amg", "block_tri_amg_fp64", "block_tri_nolag", "mono_gmres"])
def test_c1_time_loop_matches_oracle(mode):
    """Config C1, 10 backward-Euler steps: each step's field within 1e-8 relative L2 of the oracle
    (oracle: Newton + sparse LU to rtol 1e-12).  The device solve runs SNES rtol 1e-10 / KSP rtol 1e-12."""
    prob, x0 = c1_problem()
    recs, _ = osolver.run(prob, x0, 10, linear="lu", rtol=1e-12, atol=1e-14)
    eng = make_engine(prob)
    eng.set_prev(x0)
    eng.set_state(np.zeros(prob.ndof))
    kw = dict(snes_rtol=1e-10, snes_atol=1e-13, ksp_rtol=1e-12)
    if mode == "mono_gmres":
        kw.update(solver=1)
    elif mode == "block_tri_jacobi":
        kw.update(solver=0, pc=0)
    elif mode == "block_tri_amg":
        kw.update(solver=0, pc=1)          # FP32-storage V-cycle inside FP64 PCG
    elif mode == "block_tri_amg_fp64":
        kw.update(solver=0, pc=2)
    else:
        kw.update(solver=0, pc=0, lag_mechanics=0)
    nb = 3
    for k in range(1, 11):
        st = eng.step(1, **kw)[0]
        assert st["converged"] == 1
        x = eng.get_state()
        ref = recs[k][2]
        for comp, name in ((slice(2, None, nb), "concentration"),):
            e = np.linalg.norm(x[comp] - ref[comp]) / np.linalg.norm(ref[comp])
            assert e < 1e-8, (k, name, e)
        u = x.reshape(-1, nb)[:, :2]
        ur = ref.reshape(-1, nb)[:, :2]
        e = np.linalg.norm(u - ur) / np.linalg.norm(ur)
        assert e < 1e-8, (k, "displacement", e)
    eng.close()


@pytest.mark.parametrize("env", [{}, {"GLIMS_NO_COND_GRAPH": "1"}, {"GLIMS_NO_GRAPH": "1"}, {"GLIMS_AMG_FP16": "0"},
                                 {"GLIMS_AMG_FP16": "2"}, {"GLIMS_PCG_UNROLL": "1"}])
def test_3d_coupled_steps_match_oracle(env, monkeypatch):
    """3D two-tissue coupled box (config C3 shrunk to 8^3): 3 steps within 1e-8 relative L2 -- with the default solver
    plumbing (conditional CUDA graphs with the convergence test on the device, FP16 fine-level matrix in the V-cycle,
    four K_cc iterations per graph body) and with each of those measures switched off / widened."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    coords, cells = meshes.box_mesh((0, 0, 0), (1, 1, 1), 8, 8, 8)
    cm = (coords[cells].mean(axis=1)[:, 0] >= 0.5).astype(np.int32)
    mats = fem.Materials.from_E_nu([3e-3, 1e-3], [0.45, 0.45], [0.002, 0.0], [0.05, 0.0], [0.1, 0.0])
    bv = meshes.boundary_vertices(cells, len(coords))
    dofs = np.sort((bv[:, None] * 4 + np.arange(3)[None, :]).ravel())
    prob = fem.Problem(coords, cells, cm, mats, dt=1.0, bc_dofs=dofs, bc_vals=np.zeros(len(dofs)))
    x0 = np.zeros(prob.ndof)
    x0[3::4] = np.exp(-40 * ((coords - np.array([0.35, 0.5, 0.5])) ** 2).sum(axis=1))
    recs, _ = osolver.run(prob, x0, 3, linear="lu", rtol=1e-12, atol=1e-15)
    eng = make_engine(prob)
    eng.set_prev(x0)
    eng.set_state(np.zeros(prob.ndof))
    for k in range(1, 4):
        st = eng.step(1, snes_rtol=1e-10, snes_atol=1e-14, ksp_rtol=1e-12)[0]
        assert st["converged"] == 1
        x = eng.get_state().reshape(-1, 4)
        ref = recs[k][2].reshape(-1, 4)
        assert np.linalg.norm(x[:, 3] - ref[:, 3]) / np.linalg.norm(ref[:, 3]) < 1e-8
        assert np.linalg.norm(x[:, :3] - ref[:, :3]) / np.linalg.norm(ref[:, :3]) < 1e-8
    eng.close()


def test_not_converged_is_reported():
    """GLIMS_ERR_NOT_CONVERGED surfaces as an exception (simulation_base.py:301-305 catches it)."""
    from glimslib_b200.engine import SolverNotConverged
    prob, x0 = c1_problem(nx=20)
    eng = make_engine(prob)
    eng.set_prev(x0)
    eng.set_state(np.zeros(prob.ndof))
    with pytest.raises(SolverNotConverged):
        eng.step(1, max_newton=1, snes_rtol=1e-14, snes_atol=1e-300)
    eng.close()


@pytest.mark.parametrize("d", [2, 3])
def test_heterogeneous_jittered_mesh_steps_match_oracle(d):
    """Irregular (jittered) mesh, three materials with a 100x stiffness contrast and nu up to 0.49, non-zero
    Dirichlet data on u and c, a load vector: two steps through the default solver stack (AMG with FP32 V-cycle,
    projection, graph replay) within 1e-8 relative L2 of the oracle (Newton + sparse LU)."""
    prob, rng = small_problem(d, seed=5, n=9 if d == 3 else 24, jitter=0.25)
    prob.mats = fem.Materials.from_E_nu([3e-3, 3e-1, 2e-2], [0.45, 0.3, 0.49], [0.05, 0.02, 0.0], [0.2, 0.05, 0.0],
                                        [0.15, 0.0, 0.3])
    prob.f_ext *= 1e-2
    nb = d + 1
    x0 = np.zeros(prob.ndof)
    x0[nb - 1::nb] = np.exp(-8 * ((prob.coords - prob.coords.mean(axis=0)) ** 2).sum(axis=1))
    recs, _ = osolver.run(prob, x0, 2 * prob.dt, linear="lu", rtol=1e-12, atol=1e-15)
    eng = make_engine(prob)
    eng.set_prev(x0)
    eng.set_state(np.zeros(prob.ndof))
    for k in (1, 2):
        st = eng.step(1, snes_rtol=1e-11, snes_atol=1e-14, ksp_rtol=1e-12)[0]
        assert st["converged"] == 1
        x = eng.get_state().reshape(-1, nb)
        ref = recs[k][2].reshape(-1, nb)
        assert np.linalg.norm(x[:, d] - ref[:, d]) / np.linalg.norm(ref[:, d]) < 1e-8
        assert np.linalg.norm(x[:, :d] - ref[:, :d]) / np.linalg.norm(ref[:, :d]) < 1e-8
    eng.close()


def test_pure_neumann_mechanics_is_reported_not_hung():
    """Quirk Q1: the shipped 3D atlas cases end up with no Dirichlet condition, so K_uu is singular (rigid modes).
    The load of this model is self-equilibrated, so PCG still converges on the consistent system; the run must
    terminate with a converged flag or a clean SolverNotConverged -- never hang or return NaN."""
    from glimslib_b200.engine import SolverNotConverged
    prob, rng = small_problem(3, seed=7, n=6, with_bc=False)
    prob.f_ext = None
    eng = make_engine(prob)
    x0 = np.zeros(prob.ndof)
    x0[3::4] = np.exp(-8 * ((prob.coords - prob.coords.mean(axis=0)) ** 2).sum(axis=1))
    eng.set_prev(x0)
    eng.set_state(np.zeros(prob.ndof))
    try:
        st = eng.step(1, max_krylov=2000)[0]
        assert st["converged"] == 1 and np.isfinite(eng.get_state()).all()
    except SolverNotConverged:
        pass
    eng.close()


def test_indefinite_concentration_block_falls_back_to_gmres():
    """dt*rho > 1 makes K_cc = (1 - dt*rho) M + ... indefinite: PCG is not applicable, the reference's GMRES is.
    The block solver must fall back to the monolithic GMRES update and still match the oracle."""
    prob, rng = small_problem(2, seed=11, n=10, with_bc=True)
    prob.mats = fem.Materials.from_E_nu([3e-3, 3e-3, 3e-3], [0.3, 0.3, 0.3], [1e-3, 1e-3, 1e-3], [1.0, 1.0, 1.0], [0.1, 0.1, 0.1])
    prob.dt = 1.5          # K_cc spectrum at the start: [-2.4e-3, 1.2e-2] (checked with the oracle)
    prob.f_ext = None
    bc_u = prob.bc_dofs % 3 != 2
    prob.bc_dofs, prob.bc_vals = prob.bc_dofs[bc_u], 0.0 * prob.bc_vals[bc_u]
    x0 = np.zeros(prob.ndof)
    x0[2::3] = 0.1 + 0.02 * np.cos(3 * prob.coords[:, 0])
    x_ref, _ = osolver.newton(prob, x0.copy(), x0, linear="lu", rtol=1e-12, atol=1e-14)
    eng = make_engine(prob)
    eng.set_prev(x0)
    eng.set_state(x0)
    st = eng.step(1, snes_rtol=1e-11, snes_atol=1e-14, ksp_rtol=1e-12, max_krylov=3000)[0]
    assert st["converged"] == 1
    x = eng.get_state()
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-7
    eng.close()
