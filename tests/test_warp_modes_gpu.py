"""fb_set_warp: the `warp` argument of CorotationalLinearFEMForceModel(fem, warp).  0 = linear FEM,
2 = exact tangent (corotationalLinearFEM.cpp:296-428).  K and f through the C ABI must be BIT-IDENTICAL to
ComputeForceAndStiffnessMatrix(u, f, K, warp) of the unmodified reference: committed golden outputs
(tests/golden/warp_*.npz, made by `make_golden.py warp`), and live against oracle/_ref where that library is present."""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mesh(name):
    if name == "cube5":
        return cases.cube_case(5)[:3]
    m = np.load(os.path.join(GOLDEN, "mesh_egg_shell_sample.npz"))
    return m["verts"], m["tets"], m["fixed"]


@pytest.mark.parametrize("name", ["cube5", "egg_shell_sample"])
def test_warp_0_and_2_bit_exact_vs_reference_golden(name):
    import fembrain_b200 as fb

    v, t, fixed = _mesh(name)
    g = np.load(os.path.join(GOLDEN, f"warp_{name}.npz"))
    sim = fb.Simulation(v, t, fixed)
    assert sim.warp == 1
    f1, K1 = sim.force_and_matrix(cases.perturbation(v, 1.0, 1))
    for warp in (0, 2):
        sim.set_warp(warp)
        assert sim.warp == warp
        for seed, scale in ((1, 1.0), (2, 6.0)):
            f, K = sim.force_and_matrix(cases.perturbation(v, scale, seed))
            assert np.array_equal(f, g[f"f_w{warp}_s{seed}"]), (warp, seed)
            assert np.array_equal(K, g[f"K_w{warp}_s{seed}"]), (warp, seed)
    # back to the default: the gather path again, same bits as before
    sim.set_warp(1)
    f, K = sim.force_and_matrix(cases.perturbation(v, 1.0, 1))
    assert np.array_equal(f, f1) and np.array_equal(K, K1)


def test_warp_modes_live_vs_compiled_reference_and_a_step(ref_oracle):
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(6)
    o = ref_oracle.Oracle(v, t, fixed, kind="ref")
    sim = fb.Simulation(v, t, fixed)
    u = cases.perturbation(v, 3.0, 7)
    for warp in (0, 1, 2):
        sim.set_warp(warp)
        f, K = sim.force_and_matrix(u)
        fr, Kr = o.force_and_matrix_warp(u, warp)
        assert np.array_equal(f, fr) and np.array_equal(K, Kr), warp
    # a whole step with the exact tangent: Keff / rhs are formed from that K by the same epilogue and the solve converges
    sim.set_warp(2)
    sim.set_external_forces(cases.point_load(sim.r, load))
    sim.set_state(u * 0.1, np.zeros_like(u))
    assert sim.do_timestep() == 0 and sim.last_cg_iterations > 0
    h, dk = 0.0333, 0.01
    _, K2 = sim.force_and_matrix(u * 0.1)
    Ke = sim.K_values()
    # k != l entries of Keff = (h^2 + h dK) K exactly in the reference's order of operations: ((K dK) + ... is checked
    # bit-for-bit elsewhere for warp = 1; here: same epilogue, so the ratio is the constant h (h + dK) to rounding
    nz = np.abs(K2) > 1e-3 * np.abs(K2).max()
    ia, ja, _ = sim.K_csr()
    rows = np.repeat(np.arange(len(ia) - 1), np.diff(ia))
    off = nz & (rows % 3 != ja % 3)   # the consistent mass matrix adds to the k == l entries of EVERY 3x3 block
    assert np.allclose(Ke[off] / K2[off], h * (h + dk), rtol=1e-12)


def test_set_warp_errors():
    import fembrain_b200 as fb

    v, t, fixed, _ = cases.cube_case(4)
    sim = fb.Simulation(v, t, fixed)
    with pytest.raises(fb.FemBrainError):
        sim.set_warp(3)
    sim.set_grid(4)
    sim.set_solver("mg")
    with pytest.raises(fb.FemBrainError):
        sim.set_warp(2)
