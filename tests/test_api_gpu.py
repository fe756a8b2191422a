"""GPU tests of the C-ABI surface beyond the plain step: golden fixtures of the unmodified reference,
fixed-vertex changes, constrained-DOF lists, per-element materials, error codes, degenerate input."""
import numpy as np
import pytest

from tests import cases
from tests.test_oracle import GOLD, check, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLD)
def test_cuda_matches_reference_golden(name):
    """CUDA path against outputs of the UNMODIFIED reference (tests/golden/ref_*.npz): integers and every float
    that defines the linear system bit-exact; states to solver tolerance; tight solve to 1e-8."""
    import fembrain_b200 as fb

    z, prm = load_golden(name)
    sim = fb.Simulation(z["verts"], z["tets"], z["fixed"], youngs_modulus=prm["E"], poisson_ratio=prm["nu"], density=prm["rho"],
                        timestep=prm["h"], damping_mass=prm["damp_mass"], damping_stiffness=prm["damp_stiffness"])
    ia, ja, _ = sim.K_csr()
    check(z, "K_ia", ia), check(z, "K_ja", ja)
    mia, mja, ma = sim.M_csr()
    check(z, "M_ia", mia), check(z, "M_ja", mja), check(z, "M_a", ma)
    sia, sja, _ = sim.sys_csr(values=False)
    check(z, "S_ia", sia), check(z, "S_ja", sja)
    row, col = sim.element_maps()
    check(z, "el_row", row), check(z, "el_col", col)
    sr, si = sim.super_maps()
    check(z, "super_rows", sr), check(z, "super_idx", si)
    check(z, "sub_idx", sim.submatrix_map())
    f, K = sim.force_and_matrix(z["u"])
    assert np.array_equal(f, z["f"])
    check(z, "K_a", K)
    sim.set_state(z["q0"], z["qvel0"])
    sim.set_external_forces(z["fext"])
    for step in range(3):
        assert sim.do_timestep() == 0
        check(z, f"Keff_{step}", sim.K_values())
        assert np.array_equal(sim.internal_forces(), z[f"fint_{step}"])
        assert np.array_equal(sim.rhs(), z[f"rhs_{step}"])
        ref_it = int(z["cg_iterations"][step])
        assert abs(sim.last_cg_iterations - ref_it) <= max(3, ref_it // 40), (sim.last_cg_iterations, ref_it)
        q, qv, _ = sim.get_state()
        assert cases.rel_err(q, z[f"q_{step}"]) <= 2e-4 and cases.rel_err(qv, z[f"qvel_{step}"]) <= 2e-4
        sim.set_state(z[f"q_{step}"], z[f"qvel_{step}"], np.zeros(sim.r))
    x, it = sim.solve(eps=1e-12, max_iter=20000)
    assert it > 0 and cases.rel_err(x, z["x_tight"]) <= 1e-8, cases.rel_err(x, z["x_tight"])


def test_set_fixed_vertices_rebuilds_the_constrained_system(port_oracle):
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(6)
    sim = fb.Simulation(v, t, fixed)
    f = cases.point_load(sim.r, load)
    sim.set_external_forces(f)
    sim.do_timestep()
    new_fixed = np.concatenate([fixed[::2], [load]]).astype(np.int32)  # different set, unsorted
    sim.set_fixed_vertices(new_fixed)
    sim.reset_to_rest()
    ora = port_oracle.Oracle(v, t, new_fixed, kind="port")
    ora.set_external_forces(f)
    assert sim.num_constrained == 3 * len(new_fixed) and sim.rows_sys == ora.rows_sys
    sim.do_timestep(), ora.do_timestep()
    for a, b in zip(sim.sys_csr(), ora.sys_csr()):
        assert np.array_equal(a, b)
    assert np.array_equal(sim.rhs(), ora.rhs())
    q, oq = sim.get_state()[0], ora.get_state()[0]
    assert cases.rel_err(q, oq) <= 1e-4
    assert np.all(q[sim.constrained_dofs()] == 0)


def test_constrained_dof_list_constructor(port_oracle):
    """The integrator constructor's own argument: an arbitrary sorted DOF list, not only whole vertices."""
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(5)
    dofs = np.sort(np.concatenate([3 * fixed, 3 * fixed + 1, 3 * fixed + 2])).astype(np.int32)
    a = fb.Simulation(v, t, constrained_dofs=dofs)
    b = fb.Simulation(v, t, fixed)
    f = cases.point_load(a.r, load)
    for s in (a, b):
        s.set_external_forces(f)
        s.do_timestep()
    assert np.array_equal(a.get_state()[0], b.get_state()[0])  # same list => same arithmetic, bit for bit
    # single DOFs of a vertex (e.g. only y fixed): rows/cols of just those DOFs disappear
    partial = np.sort(np.concatenate([3 * fixed + 1, [3 * load + 2]])).astype(np.int32)
    c = fb.Simulation(v, t, constrained_dofs=partial)
    c.set_external_forces(f)
    c.do_timestep()
    q = c.get_state()[0]
    assert np.all(q[partial] == 0) and np.abs(q).max() > 0
    assert c.rows_sys == c.r - len(partial)
    y = c.sys_spmv(np.ones(c.rows_sys))
    ia, ja, a_ = c.sys_csr()
    assert np.allclose(y, np.add.reduceat(a_, ia[:-1]), rtol=1e-12, atol=1e-6 * np.abs(a_).max())
    with pytest.raises(fb.FemBrainError):
        fb.Simulation(v, t, constrained_dofs=np.array([5, 3], np.int32))  # unsorted => invalid argument


def test_per_element_materials_uniform_arrays_equal_scalars():
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(5)
    nT = len(t)
    a = fb.Simulation(v, t, fixed)
    b = fb.Simulation(v, t, fixed, materials=(np.full(nT, 1e7), np.full(nT, 0.46), np.full(nT, 1000.0)))
    u = cases.perturbation(v, 1.0, 2)
    for x, y in zip(a.force_and_matrix(u), b.force_and_matrix(u)):
        assert np.array_equal(x, y)
    assert np.array_equal(a.M_csr()[2], b.M_csr()[2])
    # two regions: stiffer top half changes K but not the pattern
    E = np.where(np.arange(nT) < nT // 2, 1e7, 3e7)
    c = fb.Simulation(v, t, fixed, materials=(E, None, None))
    Kc = c.force_and_matrix(u)[1]
    assert not np.array_equal(Kc, a.force_and_matrix(u)[1])
    assert np.array_equal(c.K_csr()[1], a.K_csr()[1])


def test_error_codes():
    import fembrain_b200 as fb
    from fembrain_b200 import api

    v, t, fixed, load = cases.cube_case(4)
    bad = t.copy()
    bad[3, 2] = len(v) + 5
    with pytest.raises(fb.FemBrainError) as e:
        fb.Simulation(v, bad, fixed)
    assert e.value.status == api.FB_ERR_BAD_MESH
    with pytest.raises(fb.FemBrainError) as e:
        fb.Simulation(np.vstack([v, [[9.0, 9.0, 9.0]]]), t, fixed)  # a vertex in no tet
    assert e.value.status == api.FB_ERR_BAD_MESH
    with pytest.raises(fb.FemBrainError) as e:
        fb.Simulation(v, t, np.array([0, 0], np.int32))  # duplicate fixed vertex
    assert e.value.status == api.FB_ERR_INVALID_ARGUMENT
    # PCG iteration cap: the reference printf+exit(-1)s; here a status code and an untouched state
    sim = fb.Simulation(v, t, fixed, cg_max_iterations=3)
    sim.set_external_forces(cases.point_load(sim.r, load))
    assert sim.step_raw() == api.FB_ERR_SOLVER_NOT_CONVERGED
    assert sim.last_cg_iterations == -3
    assert not sim.get_state()[0].any()


def test_degenerate_element_propagates_nan_like_the_reference(port_oracle):
    """blobtree/tumor.veg (the reference's default model) has an element with a repeated vertex: the reference's K
    holds NaNs there.  Same NaN pattern here, same finite values elsewhere."""
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(4)
    t = t.copy()
    t[7, 3] = t[7, 2]
    sim = fb.Simulation(v, t, fixed)
    ora = port_oracle.Oracle(v, t, fixed, kind="port")
    for a, b in zip(sim.K_csr()[:2] + sim.element_maps(), ora.K_csr(values=False)[:2] + ora.element_maps()):
        assert np.array_equal(a, b)
    u = cases.perturbation(v, 1.0, 1)
    (f, K), (of, oK) = sim.force_and_matrix(u), ora.force_and_matrix(u)
    assert np.isnan(oK).any()
    assert np.array_equal(K, oK, equal_nan=True) and np.array_equal(f, of, equal_nan=True)


def test_empty_and_tiny_inputs():
    import fembrain_b200 as fb
    from fembrain_b200 import meshes

    v, t = meshes.one_tetra()
    sim = fb.Simulation(v, t, [0, 1, 2])  # a single free vertex
    f = np.zeros(12)
    f[9:12] = (10.0, -5.0, 2.0)
    sim.set_external_forces(f)
    sim.do_timestep()
    q = sim.get_state()[0]
    assert np.all(q[:9] == 0) and np.all(q[9:] != 0)
    assert sim.last_cg_iterations >= 1
    everything_fixed = fb.Simulation(v, t, [0, 1, 2, 3])
    everything_fixed.set_external_forces(f)
    everything_fixed.do_timestep()  # 0 unknowns: rho0 = 0 => zero iterations
    assert everything_fixed.last_cg_iterations == 0 and not everything_fixed.get_state()[0].any()
    empty = fb.Simulation(np.zeros((0, 3)), np.zeros((0, 4), np.int32), [])
    assert empty.r == 0 and empty.nnz_K == 0
    empty.do_timestep()


def test_state_api_roundtrip_and_zero_force_rest():
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(4)
    sim = fb.Simulation(v, t, fixed)
    rng = np.random.default_rng(0)
    q, qv, qa = rng.standard_normal((3, sim.r))
    sim.set_state(q, qv, qa)
    for a, b in zip(sim.get_state(), (q, qv, qa)):
        assert np.array_equal(a, b)
    f = rng.standard_normal(sim.r)
    sim.set_external_forces(f)
    sim.add_external_forces(f)
    assert np.array_equal(sim.get_external_forces(), f + f)
    sim.set_external_forces_to_zero()
    assert not sim.get_external_forces().any()
    sim.reset_to_rest()
    sim.do_timestep()  # at rest with no load: f_int is rounding noise (K_el P - RK x0), so is the motion
    assert np.abs(sim.get_state()[0]).max() < 1e-12


def test_sync_force_model_rebuilds_in_place(port_oracle):
    """fb_sync_force_model = Deformable::syncForceModel after a cut (DEF/Deformable.cpp:127-220): same handle, new mesh, state at
    rest, Deformable-level options kept; everything the new mesh's oracle says must hold bit for bit."""
    import time

    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(8)
    sim = fb.Simulation(v, t, fixed)
    sim.set_gravity(True)
    sim.set_floor(True, -3.0)
    sim.set_external_forces(cases.point_load(sim.r, load))
    sim.do_timestep()
    # the "cut": another mesh altogether (fewer cells, other fixed set)
    v2, t2, fixed2, load2 = cases.cube_case(6, 5, 7)
    t0 = time.perf_counter()
    sim.sync_force_model(v2, t2, fixed2)
    dt = time.perf_counter() - t0
    assert sim.r == 3 * len(v2) and dt < 5.0
    q, qv, qa = sim.get_state()
    assert not q.any() and not qv.any() and not qa.any()           # a new integrator starts at rest
    ora = port_oracle.Oracle(v2, t2, fixed2, kind="port")
    ia, ja, _ = sim.K_csr()
    oia, oja, _ = ora.K_csr(values=False)
    assert np.array_equal(ia, oia) and np.array_equal(ja, oja)
    f = cases.point_load(sim.r, load2)
    for s in (sim, ora):
        s.set_external_forces(f)
        assert s.do_timestep() == 0
    assert np.array_equal(sim.K_values(), ora.K_values()) and np.array_equal(sim.rhs(), ora.rhs())
    # options of the Deformable survived: gravity shows up in the next deformable frame's force vector
    sim.deformable_timestep()
    fe = sim.get_external_forces()
    assert np.all(fe[1::3] == -10000.0)
