"""FEMBRAIN_B200_SPMV=sym (fb_sym.cu): the solver's products from the block-upper triangle of Keff, lower blocks as transposes
out of L2.  Keff is symmetric up to its own rounding (checked here on the oracle's matrix), so PCG must behave like the
full-matrix path: same iteration counts to within a few, displacements <= 1e-8 relative after convergence (BASELINE.json)."""
import numpy as np
import pytest

from tests import cases

def _built():
    from fembrain_b200 import api

    return api.experiments_built()


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _built(), reason="shelved experiment (csrc/experiments/fb_sym.cu): build with `python -m fembrain_b200.build --experiments`")]


def _pair(monkeypatch, v, t, fixed, **kw):
    import fembrain_b200 as fb

    monkeypatch.setenv("FEMBRAIN_B200_SPMV", "sym")
    sym = fb.Simulation(v, t, fixed, **kw)
    monkeypatch.delenv("FEMBRAIN_B200_SPMV")
    full = fb.Simulation(v, t, fixed, **kw)
    return sym, full


MESHES = {
    "cube7": lambda: cases.cube_case(7)[:3],
    "slab_5x3x9": lambda: cases.cube_case(5, 3, 9)[:3],
    "cube13": lambda: cases.cube_case(13)[:3],
    "eggshell": lambda: cases.golden_mesh("eggshell"),
}


@pytest.mark.parametrize("name", list(MESHES))
def test_sym_products_match_the_full_matrix_path(port_oracle, monkeypatch, name):
    v, t, fixed = MESHES[name]()
    sym, full = _pair(monkeypatch, v, t, fixed)
    ora = port_oracle.Oracle(v, t, fixed, kind="port")
    u = cases.perturbation(v, 0.5, 3)
    u[sym.constrained_dofs()] = 0.0
    f = cases.point_load(sym.r, int(np.argmax(v[:, 1] * 1000 + v[:, 0])))
    for s in (sym, full, ora):
        s.set_state(u, np.zeros_like(u))
        s.set_external_forces(f)
    for step in range(3):
        assert sym.do_timestep() == 0 and full.do_timestep() == 0 and ora.do_timestep() == 0
        # the linear system is the reference's, bit for bit, on both paths
        assert np.array_equal(sym.rhs(), ora.rhs()) and np.array_equal(sym.K_values(), ora.K_values())
        # Keff is symmetric up to rounding: that is all the sym path relies on
        if step == 0:
            import scipy.sparse as sp

            ia, ja, a = ora.sys_csr()
            A = sp.csr_matrix((a, ja, ia), shape=(len(ia) - 1,) * 2)
            assert abs((A - A.T)).max() <= 1e-14 * abs(A).max()
        # (beam3 with this state is left out on purpose: on the reference's own arithmetic rho/rho0 dips to 1.036e-12 at
        # iteration 645 of 720 in the third step, 3.6 % above the threshold — the sym path's rounding stops there, 76 iterations
        # early, and at eps = 1e-6 on a system conditioned 3e5 that is a 1e-2 different answer.  Nothing to assert on.)
        its, itf = sym.last_cg_iterations, full.last_cg_iterations
        assert its > 0 and abs(its - itf) <= max(3, itf // 50), (its, itf)
        assert sym.last_cg_residual_ratio <= 1e-12
        q, qv, _ = sym.get_state()
        oq, oqv, _ = ora.get_state()
        assert cases.rel_err(qv, oqv) <= 1e-4 and cases.rel_err(q, oq) <= 1e-4
        for s in (sym, full):
            s.set_state(oq, oqv)  # identical inputs for the next step: U is re-packed from the new Keff
    x, it = sym.solve(eps=1e-12, max_iter=20000)
    fx, fit = full.solve(eps=1e-12, max_iter=20000)
    ox, oit = ora.solve(eps=1e-12, max_iter=20000)
    assert it > 0 and abs(it - oit) <= max(3, oit // 50), (it, fit, oit)
    assert cases.rel_err(x, ox) <= 1e-8 and cases.rel_err(x, fx) <= 1e-8
    # true residual of the sym path's solution against the ORACLE's matrix
    b = ora.rhs()
    assert np.linalg.norm(b - ora.sys_spmv(x)) <= 1e-10 * np.linalg.norm(b)


def test_sym_after_changing_fixed_vertices(port_oracle, monkeypatch):
    v, t, fixed = cases.cube_case(8)[:3]
    sym, full = _pair(monkeypatch, v, t, fixed)
    f = cases.point_load(sym.r, len(v) - 1)
    for s in (sym, full):
        s.set_external_forces(f)
        assert s.do_timestep() == 0
    fixed2 = np.unique(np.concatenate([fixed[::2], [len(v) - 3]])).astype(np.int32)
    for s in (sym, full):
        s.set_fixed_vertices(fixed2)
        s.reset_to_rest()
        s.set_external_forces(f)
        assert s.do_timestep() == 0
    assert abs(sym.last_cg_iterations - full.last_cg_iterations) <= 3
    x, _ = sym.solve(eps=1e-12, max_iter=20000)
    fx, _ = full.solve(eps=1e-12, max_iter=20000)
    assert cases.rel_err(x, fx) <= 1e-8
    # run-to-run reproducible
    x2, _ = sym.solve(eps=1e-12, max_iter=20000)
    assert np.array_equal(x, x2)
