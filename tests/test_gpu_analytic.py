"""The CUDA path against closed-form solutions (tests/analytic_cases.py) -- known answers that involve neither the oracle nor
the reference: free growth of a minimally supported body is the stress-free dilatation u = c*gamma*(x - x_fixed); boundary values
of a linear displacement field reproduce that field.  Through the C ABI; FP64; tolerance 1e-8 of the largest displacement
(Krylov tolerance 1e-13 on stiffness matrices that are nearly singular by construction in the free-growth case)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
TIGHT = dict(snes_rtol=1e-13, snes_atol=1e-16, ksp_rtol=1e-13, max_newton=30)


def _run(prob, x0):
    from glimslib_b200.engine import Engine
    eng = Engine(prob.coords, prob.cells, prob.cell_mat)
    eng.set_materials(prob.mats.table())
    eng.set_dt(prob.dt)
    eng.set_dirichlet(prob.bc_dofs, prob.bc_vals)
    eng.set_prev(x0)
    eng.set_state(np.zeros_like(x0))
    st = eng.step(1, **TIGHT)
    x = eng.get_state()
    eng.close()
    assert st[0]["converged"]
    return x


@pytest.mark.parametrize("d", [2, 3])
def test_free_growth_is_a_stress_free_dilatation(d):
    import analytic_cases as ac
    prob, x0, exact, c = ac.free_growth_case(d)
    X = _run(prob, x0).reshape(-1, d + 1)
    assert np.abs(X[:, d] - c).max() < 1e-12
    assert np.abs(X[:, :d] - exact).max() < 1e-8 * np.abs(exact).max()


@pytest.mark.parametrize("d", [2, 3])
def test_elasticity_patch_test(d):
    import analytic_cases as ac
    prob, exact = ac.patch_case(d)
    X = _run(prob, np.zeros(prob.ndof)).reshape(-1, d + 1)
    assert np.abs(X[:, :d] - exact).max() < 1e-9 * np.abs(exact).max()
