"""Committed fixtures (tests/golden/, made by tests/golden/make_golden.py from the oracle): the oracle still reproduces them
(CPU), and the CUDA path reproduces them through the C ABI without running the oracle (GPU)."""
import os

import numpy as np
import pytest

from glimslib_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"c1_nx20_3steps": (lambda: W.c1_2d_subdomains(nx=20), 3), "c3_box6_2steps": (lambda: W.c3_box(6), 2)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_committed_fixture(name):
    from oracle import fem, solver as osolver
    make, steps = CASES[name]
    w = make()
    t = w["table"]
    prob = fem.Problem(w["mesh"].coords, w["mesh"].cells, w["cell_mat"],
                       fem.Materials(t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]), w["dt"],
                       bc_dofs=w["bc_dofs"], bc_vals=w["bc_vals"])
    recs, _ = osolver.run(prob, w["x0"], steps, linear="lu", rtol=1e-12, atol=1e-14)
    gold = np.load(os.path.join(GOLD, name + ".npz"))["states"]
    assert gold.shape == (steps + 1, prob.ndof)
    for k in range(steps + 1):
        assert np.abs(recs[k][2] - gold[k]).max() <= 1e-11 * max(np.abs(gold[k]).max(), 1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_the_committed_fixture(name):
    """<= 1e-8 relative L2 per field and step (device solve: SNES rtol 1e-10, KSP rtol 1e-12)."""
    make, steps = CASES[name]
    w = make()
    d = w["mesh"].dim
    nb = d + 1
    gold = np.load(os.path.join(GOLD, name + ".npz"))["states"]
    eng = W.build_engine(w)
    eng.set_prev(w["x0"])
    eng.set_state(np.zeros_like(w["x0"]))
    for k in range(1, steps + 1):
        st = eng.step(1, snes_rtol=1e-10, snes_atol=1e-14, ksp_rtol=1e-12)[0]
        assert st["converged"] == 1
        x = eng.get_state().reshape(-1, nb)
        ref = gold[k].reshape(-1, nb)
        assert np.linalg.norm(x[:, d] - ref[:, d]) <= 1e-8 * np.linalg.norm(ref[:, d])
        assert np.linalg.norm(x[:, :d] - ref[:, :d]) <= 1e-8 * np.linalg.norm(ref[:, :d])
    eng.close()
