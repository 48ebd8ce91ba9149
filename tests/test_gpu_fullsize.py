"""Full-size GPU checks (BASELINE.json configs C2 / C3 at their named sizes) through size-independent properties,
because the CPU oracle cannot finish these sizes in seconds: closed-form logistic step,
mass conservation, symmetry of the eliminated blocks, atomic == gather == slice assembly, monolithic |F| after a
step, and idempotence of a converged state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _wl(name):
    from glimslib_b200 import workloads as W
    return W.c3_box(55) if name == "C3" else W.c2_2d_1m(707)


def _engine(w):
    from glimslib_b200 import workloads as W
    return W.build_engine(w)


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_uniform_concentration_takes_the_closed_form_logistic_step(name):
    """Spatially uniform c0 and no c-BC: each backward-Euler step solves c - c0 - dt*rho*c*(1-c) = 0 in every
    vertex of the proliferating tissue; with D > 0 only where rho > 0 is uniform we use one material."""
    w = _wl(name)
    nb = w["mesh"].dim + 1
    t = w["table"].copy()
    t[:, 2], t[:, 3], t[:, 4] = 0.01, 0.3, 0.0          # same D, rho everywhere; no coupling
    eng = _engine(w)
    eng.set_materials(t)
    x0 = np.zeros(eng.ndof)
    x0[nb - 1::nb] = 0.2
    eng.set_prev(x0)
    eng.set_state(x0)
    c = 0.2
    for _ in range(2):
        st = eng.step(1, snes_rtol=1e-12, snes_atol=1e-13, ksp_rtol=1e-13)[0]
        assert st["converged"] == 1
        a, b, cc = w["dt"] * 0.3, 1 - w["dt"] * 0.3, -c
        c = (-b + np.sqrt(b * b - 4 * a * cc)) / (2 * a)
        x = eng.get_state()
        assert np.abs(x[nb - 1::nb] - c).max() < 1e-11
        assert np.abs(x.reshape(-1, nb)[:, :nb - 1]).max() < 1e-14      # gamma = 0 => u == 0
    eng.close()


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_mass_is_conserved_without_proliferation(name):
    w = _wl(name)
    mesh = w["mesh"]
    nb = mesh.dim + 1
    t = w["table"].copy()
    t[:, 3] = 0.0
    eng = _engine(w)
    eng.set_materials(t)
    X = mesh.coords[mesh.cells]
    vol = np.abs(np.linalg.det(X[:, 1:] - X[:, :1])) / {2: 2.0, 3: 6.0}[mesh.dim]
    lumped = np.bincount(mesh.cells.ravel(), weights=np.repeat(vol / nb, nb), minlength=mesh.num_vertices())
    eng.set_prev(w["x0"])
    eng.set_state(w["x0"])
    m0 = lumped @ w["x0"][nb - 1::nb]
    for _ in range(2):
        eng.step(1, snes_rtol=1e-11, snes_atol=1e-13, ksp_rtol=1e-12)
        assert abs(lumped @ eng.get_state()[nb - 1::nb] - m0) < 1e-10 * abs(m0)
    eng.close()


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_assembly_variants_agree_and_blocks_are_symmetric(name):
    """atomic == gather == slice == tile at full size (<= 1e-12: summation order differs), and the Dirichlet-eliminated
    K_uu, K_cc satisfy x.(A y) == y.(A x) (needed by PCG)."""
    from glimslib_b200 import _native as N
    w = _wl(name)
    d = w["mesh"].dim
    eng = _engine(w)
    rng = np.random.default_rng(0)
    eng.set_state(0.1 * rng.standard_normal(eng.ndof))
    ref = None
    xu, yu = rng.standard_normal(eng.n_vertices * d), rng.standard_normal(eng.n_vertices * d)
    xc, yc = rng.standard_normal(eng.n_vertices), rng.standard_normal(eng.n_vertices)
    for kernel in (N.ASMK_ATOMIC, N.ASMK_GATHER, N.ASMK_SLICE, N.ASMK_TILE):
        eng.assemble(what=N.ASM_JACOBIAN, kernel=kernel, apply_bc=2)
        res = (eng.spmv(1, xu), eng.spmv(2, xc), eng.spmv(0, np.ones(eng.ndof)))
        if ref is None:
            ref = res
            assert abs(yu @ res[0] - xu @ eng.spmv(1, yu)) < 1e-11 * np.linalg.norm(res[0]) * np.linalg.norm(yu)
            assert abs(yc @ res[1] - xc @ eng.spmv(2, yc)) < 1e-11 * np.linalg.norm(res[1]) * np.linalg.norm(yc)
        else:
            for a, b in zip(res, ref):
                assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
    eng.close()


def test_c3_step_meets_the_monolithic_tolerance_and_is_idempotent():
    """After a step the monolithic residual (recomputed from scratch by the element kernel) is below the SNES
    tolerance, and re-solving from the converged state with the same u_previous changes nothing."""
    from glimslib_b200 import _native as N
    w = _wl("C3")
    eng = _engine(w)
    eng.set_prev(w["x0"])
    eng.set_state(np.zeros(eng.ndof))
    st = eng.step(1)[0]
    assert st["converged"] == 1 and st["fnorm"] <= max(1e-9 * st["fnorm0"], 1e-10)
    x1 = eng.get_state()
    eng.set_prev(w["x0"])
    eng.set_state(x1)
    eng.assemble(what=N.ASM_RESIDUAL, apply_bc=1)
    assert np.linalg.norm(eng.residual()) <= max(1e-9 * st["fnorm0"], 1e-10) * 1.0001
    st2 = eng.step(1)[0]
    assert st2["newton_its"] == 0
    assert np.array_equal(eng.get_state(), x1)
    eng.close()
