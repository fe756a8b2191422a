"""FEMBRAIN_B200_SPMV=tma (fb_tma.cu): the solver's products with the matrix stream staged through shared memory by 1-D bulk
copies (cp.async.bulk + mbarrier).  Same arithmetic per row as the default kernel, so PCG must behave the same.
EXPERIMENTAL path: these tests only run with FEMBRAIN_B200_TEST_EXPERIMENTAL=1 until the kernel has been measured."""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FEMBRAIN_B200_TEST_EXPERIMENTAL") != "1", reason="experimental kernel: set FEMBRAIN_B200_TEST_EXPERIMENTAL=1")]

MESHES = {
    "cube7": lambda: cases.cube_case(7)[:3],
    "cube13": lambda: cases.cube_case(13)[:3],
    "cube30": lambda: cases.cube_case(30)[:3],   # more tiles than CTAs: the stage ring wraps
    "eggshell": lambda: cases.golden_mesh("eggshell"),
}


@pytest.mark.parametrize("name", list(MESHES))
def test_tma_products_match_the_default_path(monkeypatch, name):
    import fembrain_b200 as fb

    v, t, fixed = MESHES[name]()
    monkeypatch.setenv("FEMBRAIN_B200_SPMV", "tma")
    tma = fb.Simulation(v, t, fixed)
    monkeypatch.delenv("FEMBRAIN_B200_SPMV")
    full = fb.Simulation(v, t, fixed)
    u = cases.perturbation(v, 0.5, 3)
    u[tma.constrained_dofs()] = 0.0
    f = cases.point_load(tma.r, int(np.argmax(v[:, 1] * 1000 + v[:, 0])))
    for s in (tma, full):
        s.set_state(u, np.zeros_like(u))
        s.set_external_forces(f)
    for step in range(2):
        assert tma.do_timestep() == 0 and full.do_timestep() == 0
        assert np.array_equal(tma.rhs(), full.rhs())
        # per-row arithmetic is identical; only the per-CTA grouping of the dot-product sums differs
        assert abs(tma.last_cg_iterations - full.last_cg_iterations) <= max(3, full.last_cg_iterations // 50)
        q, qv, _ = tma.get_state()
        fq, fqv, _ = full.get_state()
        assert cases.rel_err(qv, fqv) <= 1e-4 and cases.rel_err(q, fq) <= 1e-4
        tma.set_state(fq, fqv)
    x, it = tma.solve(eps=1e-12, max_iter=20000)
    fx, fit = full.solve(eps=1e-12, max_iter=20000)
    assert it > 0 and cases.rel_err(x, fx) <= 1e-8
