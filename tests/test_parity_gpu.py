"""GPU parity: CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): CSR structure and index maps bit-exact; FP64 stiffness values within
1e-12 relative (measured: bit-exact); displacements within 1e-8 relative after convergence to the same
residual.  The oracle is oracle/vega_port.c (pinned bit-for-bit to the compiled reference in
tests/test_oracle.py) and, where it was built, the compiled reference itself (oracle/_ref).
"""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _mk(port_oracle, v, t, fixed, **kw):
    import fembrain_b200 as fb

    sim = fb.Simulation(v, t, fixed, **kw)
    ora = port_oracle.Oracle(v, t, fixed, kind="port")
    return sim, ora


CASES = {
    "two_tetra": lambda: (*__import__("fembrain_b200").meshes.two_tetra(), np.array([0], np.int32)),
    "one_tetra": lambda: (*__import__("fembrain_b200").meshes.one_tetra(), np.array([1], np.int32)),
    "cube3": lambda: cases.cube_case(3)[:3],
    "cube7": lambda: cases.cube_case(7)[:3],
    "slab_5x3x9": lambda: cases.cube_case(5, 3, 9)[:3],
    "cube12_unsorted_fixed": lambda: (lambda v, t, f, l: (v, t, f[::-1].copy()))(*cases.cube_case(12)),
    "cube6_nofixed": lambda: (*cases.cube_case(6)[:2], np.zeros(0, np.int32)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_structure_bit_exact(port_oracle, name):
    v, t, fixed = CASES[name]()
    sim, ora = _mk(port_oracle, v, t, fixed)
    assert sim.r == ora.r and sim.nnz_K == ora.nnz_K and sim.nnz_M == ora.nnz_M
    assert sim.rows_sys == ora.rows_sys and sim.nnz_sys == ora.nnz_sys
    ia, ja, _ = sim.K_csr()
    oia, oja, _ = ora.K_csr(values=False)
    assert np.array_equal(ia, oia) and np.array_equal(ja, oja)
    for a, b in zip(sim.element_maps(), ora.element_maps()):
        assert np.array_equal(a, b)
    for a, b in zip(sim.M_csr(), ora.M_csr()):
        assert np.array_equal(a, b)  # mass values bit-exact too
    assert np.array_equal(sim.submatrix_map(), ora.submatrix_map())
    sia, sja, _ = sim.sys_csr(values=False)
    osia, osja, _ = ora.sys_csr(values=False)
    assert np.array_equal(sia, osia) and np.array_equal(sja, osja)
    for a, b in zip(sim.super_maps(), ora.super_maps()):
        assert np.array_equal(a, b)
    mi, k0 = sim.element_data()
    omi, ok0 = ora.element_data()
    assert np.array_equal(mi.reshape(-1, 4, 4)[:, :, :3], omi.reshape(-1, 4, 4)[:, :, :3])
    assert np.array_equal(k0, ok0)


@pytest.mark.parametrize("name", list(CASES))
def test_force_and_stiffness_bit_exact(port_oracle, name):
    v, t, fixed = CASES[name]()
    sim, ora = _mk(port_oracle, v, t, fixed)
    for seed, scale in ((0, 0.0), (1, 1.0), (2, 4.0)):
        u = cases.perturbation(v, scale, seed)
        f, K = sim.force_and_matrix(u)
        of, oK = ora.force_and_matrix(u)
        assert cases.rel_err(K, oK) <= 1e-12  # the stated bar
        assert cases.rel_err(f, of) <= 1e-12
        assert np.array_equal(K, oK), f"K not bit-exact: rel {cases.rel_err(K, oK):.3e}"
        assert np.array_equal(f, of), f"f not bit-exact: rel {cases.rel_err(f, of):.3e}"


@pytest.mark.parametrize("name", ["two_tetra", "cube7", "slab_5x3x9", "cube12_unsorted_fixed"])
def test_timestep_parity(port_oracle, name):
    v, t, fixed = CASES[name]()
    sim, ora = _mk(port_oracle, v, t, fixed)
    r = sim.r
    load_vertex = int(np.argmax(v[:, 1] * 1000 + v[:, 0]))
    f = cases.point_load(r, load_vertex)
    u0 = cases.perturbation(v, 0.5, 3)
    fixed_dofs = sim.constrained_dofs()
    u0[fixed_dofs] = 0.0
    v0 = 0.3 * cases.perturbation(v, 0.5, 4)
    v0[fixed_dofs] = 0.0
    for s in (sim, ora):
        s.set_state(u0, v0)
        s.set_external_forces(f)
    for step in range(3):
        assert sim.do_timestep() == 0 and ora.do_timestep() == 0
        # everything that defines the linear system is bit-exact
        assert np.array_equal(sim.K_values(), ora.K_values()), "Keff"
        assert np.array_equal(sim.internal_forces(), ora.internal_forces()), "fint"
        if step == 0:
            assert np.array_equal(sim.rhs(), ora.rhs()), "rhs"
        _, oit = ora.solve(eps=1e-6)
        assert abs(sim.last_cg_iterations - oit) <= max(3, oit // 50), (sim.last_cg_iterations, oit)
        q, qv, qa = sim.get_state()
        oq, oqv, oqa = ora.get_state()
        # both stopped at eps = 1e-6 (possibly an iteration apart): loose check here, tight check below
        assert cases.rel_err(qv, oqv) <= 1e-4 and cases.rel_err(q, oq) <= 1e-4
        assert np.all(q[fixed_dofs] == 0) and np.all(qv[fixed_dofs] == 0) and np.all(qa == 0)
        # keep the two trajectories on identical inputs for the next step's bit-exact checks
        sim.set_state(oq, oqv, oqa)


@pytest.mark.parametrize("name", ["cube7", "slab_5x3x9"])
def test_converged_displacement_1e8(port_oracle, name):
    """SURVEY §7 protocol (b): both solvers driven to eps = 1e-12 on the same system => <= 1e-8 relative."""
    v, t, fixed = CASES[name]()
    sim, ora = _mk(port_oracle, v, t, fixed)
    f = cases.point_load(sim.r, int(np.argmax(v[:, 1] * 1000 + v[:, 0])))
    for s in (sim, ora):
        s.set_external_forces(f)
        s.do_timestep()
    x, it = sim.solve(eps=1e-12, max_iter=20000)
    ox, oit = ora.solve(eps=1e-12, max_iter=20000)
    assert it > 0 and oit > 0
    assert cases.rel_err(x, ox) <= 1e-8, cases.rel_err(x, ox)
    # displacement after the state update q = q0 + h (v0 + dv) inherits the bound
    h = sim.params.timestep
    assert cases.rel_err(h * x, h * ox) <= 1e-8
    # SpMV on the constrained system
    y, oy = sim.sys_spmv(ox), ora.sys_spmv(ox)
    assert cases.rel_err(y, oy) <= 1e-13


def _hub_mesh(n, seed=5):
    """n tetrahedra that share ONE vertex (the hub) and nothing else: a vertex with n incident elements."""
    rng = np.random.default_rng(seed)
    verts = [np.zeros(3)]
    tets = []
    for k in range(n):
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        a = np.cross(d, rng.standard_normal(3))
        a /= np.linalg.norm(a)
        b = np.cross(d, a)
        base = len(verts)
        verts += [d + 0.3 * a, d + 0.3 * b, 1.4 * d - 0.2 * a]
        tets.append([0, base, base + 1, base + 2])
    return np.array(verts), np.array(tets, np.int32), np.array([1, 2, 3], np.int32)


@pytest.mark.parametrize("n", [20, 150, 230, 300])  # 128 / 192 / 256 incidences per CTA, then the two-phase path
def test_assembly_paths_by_vertex_valence_bit_exact(port_oracle, n):
    v, t, fixed = _hub_mesh(n)
    sim, ora = _mk(port_oracle, v, t, fixed)
    u = 0.05 * np.random.default_rng(n).standard_normal(v.size)
    f, K = sim.force_and_matrix(u)
    of, oK = ora.force_and_matrix(u)
    assert np.array_equal(K, oK) and np.array_equal(f, of)
    for s in (sim, ora):
        s.set_state(u, np.zeros_like(u))
        s.set_external_forces(cases.point_load(sim.r, 5))
        assert s.do_timestep() == 0
    assert np.array_equal(sim.K_values(), ora.K_values()) and np.array_equal(sim.rhs(), ora.rhs())
    assert np.array_equal(sim.internal_forces(), ora.internal_forces())


@pytest.mark.parametrize("env", [{"FEMBRAIN_B200_ASSEMBLY": "twophase"}, {"FEMBRAIN_B200_GA_CAP": "192"}, {"FEMBRAIN_B200_GA_CAP": "256"}])
def test_assembly_variants_identical_on_a_cube(port_oracle, env, monkeypatch):
    """Every assembly configuration (two-phase scratch path; row-gather at 128/192/256 incidences per CTA) produces the same bits."""
    for k, val in env.items():
        monkeypatch.setenv(k, val)
    v, t, fixed = cases.cube_case(9)[:3]
    sim, ora = _mk(port_oracle, v, t, fixed)
    u = cases.perturbation(v, 2.0, 9)
    f, K = sim.force_and_matrix(u)
    of, oK = ora.force_and_matrix(u)
    assert np.array_equal(K, oK) and np.array_equal(f, of)
    for s in (sim, ora):
        s.set_state(u, np.zeros_like(u))
        assert s.do_timestep() == 0
    assert np.array_equal(sim.K_values(), ora.K_values()) and np.array_equal(sim.rhs(), ora.rhs())


def polar_edge_states(v, t):
    """Displacements that drive chosen tets through the two special branches of the rotation extraction:
    det(F) < 0 -> R = -R (corotationalLinearFEM.cpp:264-268) and det == 0 -> the Newton loop breaks with the warning
    (polarDecomposition.cpp:63-67).  'inverted': vertex 3 of tet 0 mirrored through the plane of its other three vertices;
    'flat': the same vertex moved INTO that plane, so the deformed tet has exactly zero volume (x-aligned coordinates make
    the determinant exactly 0.0 in floating point)."""
    out = {}
    a, b, c, d = (v[t[0, k]] for k in range(4))
    n = np.cross(b - a, c - a)
    n /= np.linalg.norm(n)
    h = float(np.dot(d - a, n))
    u = np.zeros_like(v)
    u[t[0, 3]] = -2.0 * h * n
    out["inverted"] = u.reshape(-1).copy()
    u[t[0, 3]] = -1.0 * h * n
    out["flat"] = u.reshape(-1).copy()
    # every vertex of the mesh squashed onto the plane y = 0: ALL tets flat (what the floor snap does to a layer of the mesh)
    u = np.zeros_like(v)
    u[:, 1] = -v[:, 1]
    out["all_flat"] = u.reshape(-1).copy()
    # the whole mesh mirrored (x -> -x): every tet inverted
    u = np.zeros_like(v)
    u[:, 0] = -2.0 * v[:, 0]
    out["all_inverted"] = u.reshape(-1).copy()
    return out


@pytest.mark.parametrize("name", ["one_tetra", "two_tetra", "cube3"])
def test_polar_edge_branches_bit_exact(port_oracle, name):
    """Aimed at the det < 0 and det == 0 branches (a9): K and f stay bit-identical to the oracle, NaNs included where the
    reference itself produces them (a zero determinant divides by zero in the next Newton step of a DIFFERENT element only
    if that element is degenerate; here the loop breaks and R is whatever the iteration held)."""
    v, t, fixed = CASES[name]()
    sim, ora = _mk(port_oracle, v, t, fixed)
    for label, u in polar_edge_states(v, t).items():
        f, K = sim.force_and_matrix(u)
        of, oK = ora.force_and_matrix(u)
        assert np.array_equal(K, oK, equal_nan=True), f"{name}/{label}: K"
        assert np.array_equal(f, of, equal_nan=True), f"{name}/{label}: f"
        if label == "inverted":
            assert np.isfinite(oK).all() and not np.array_equal(oK, ora.force_and_matrix(np.zeros_like(u))[1])
